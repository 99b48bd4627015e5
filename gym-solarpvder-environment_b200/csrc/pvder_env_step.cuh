// Per-environment device logic of the PVDER-v0 step: half-cycle Rosenbrock integrator, anti-windup
// mode sampling, event draw, outputs (obs / reward).  Shared by the CUDA kernels
// (pvder_kernels.cu) and -- compiled as plain C++ -- by the CPU-side test harness
// tests/host_emul (test infrastructure only; the product path is the CUDA build).
//
//   advance_env    <- PVDER.step            reference gym_PVDER/envs/PVDER_env.py:138-196
//   compute_outputs<- PVDER.state :531-542, PVDER.reward_calc :231-301
//   draw_event     <- generate_simulation_events :400-411 (pvder create_random_events, A.8)
#pragma once
#include <cstring>
#include "pvder_common.cuh"

namespace pvder {

// ---- Half-cycle integrator: a Rosenbrock method in the transformed form (Hairer & Wanner, Solving ODEs II, IV.7)
//        (I/(h g) - J) K_i = f(Y_i) + sum_j (c_ij/h) K_j ,  Y_i = y + sum_j a_ij K_j ,  y+ = y + sum_i m_i K_i
// one LU per step, no Newton loop.  Two schemes are compiled from this file (PVDER_SCHEME):
//   4 (default)  ROS4-L: the L-stable 4-stage order-4 method of Hairer & Wanner's ROS4 (gamma = 0.57282; stage 4
//                re-uses f(Y_3): 3 right-hand sides + 4 solves per step).  On this model it is as accurate as Rodas4
//                at the half-cycle step (tools/integrator_study.py: every state within 15 % of Rodas4's error against the
//                tight oracle, incl. the PLL states) for 60 % of the arithmetic.
//   6            Rodas4 (IV.10): L-stable, stiffly accurate, 6 stages, y+ = Y_6 + K_6.  The round-1 scheme, kept as the
//                cross-check build.
// The order conditions of both tableaux were checked numerically for these digits (tools/integrator_study.py).
#ifndef PVDER_SCHEME
#define PVDER_SCHEME 4
#endif
#ifndef PVDER_FOLD
#define PVDER_FOLD 1   // Rodas4: fold K1..K4 into the stage-5/6 sums early: same FMAs, 2 fewer live vectors (B200: 3.11 -> 3.04 ms)
#endif
#ifndef PVDER_STAGE4_DELTA
#define PVDER_STAGE4_DELTA 1   // ROS4-L: the fourth solve yields K_4 - K_3 (b_3 is not re-added: 11 FP64 instructions fewer per step)
#endif
#ifndef PVDER_AUX_SHORT
#define PVDER_AUX_SHORT 0   // one-thread kernels: incremental side-inputs with the shorter dependency chains (aux_advance_sv).
#endif                      // B200: 1.410 -> 1.424 ms single-phase (off), 8.18 -> 8.11 ms three-lane (always on there)
#ifndef PVDER_FREE_PATH
#define PVDER_FREE_PATH 0   // 1: separate clamp-free instantiation of the stepper core, chosen per warp (experiment)
#endif
#if PVDER_SCHEME == 4
constexpr double RG = 0.57282;
#else
constexpr double RG = 0.25;
#endif
template <int TAG>
struct RodasCoefT {   // a_ij and c_ij/h, read by DFMA straight from the constant bank
  double a21, a31, a32, a41, a42, a43, a51, a52, a53, a54;
  double c21, c31, c32, c41, c42, c43, c51, c52, c53, c54, c61, c62, c63, c64, c65;
  double ghinv;     // 1/(h*gamma)
  double luc[16];   // launch-constant reciprocal pivots of the symbolic LU (M::lu_consts)
  // Unit-pivot rows (M::unit_row: the pure integrators x, xDC, xQ, xPLL): their equation is scaled by h*gamma,
  // i.e. their gains and their sums of c_ij/h K_j are -- then the pivot is exactly 1 and costs no multiply.
  double cs21, cs31, cs32, cs41, cs42, cs43, cs51, cs52, cs53, cs54, cs61, cs62, cs63, cs64, cs65;   // c_ij * gamma
  double kx, kdc, kq, kpll;   // Ki_GCC, Ki_DC, Ki_Q, Ki_PLL times h*gamma
  double du_free, du_frz;     // reciprocal pivot of a free / clamped u row: 1/(1/(h g) + wp), h g
  // ROS4-L only: weights of y+ = y + sum m_i K_i, and d4j = (c4j - c3j)/h (stage 4 re-uses stage 3's right-hand side:
  // b_4 = b_3 + d41 K1 + d42 K2 + c43/h K3); ds4j = the same times h*gamma for the unit-pivot rows
  double m1, m2, m3, m4;
  double d41, d42, ds41, ds42;
  double m34;       // m3 + m4 (PVDER_STAGE4_DELTA: the fourth solve yields K_4 - K_3)
};
using RodasCoef = RodasCoefT<0>;
// The kernels' coefficient table (a __grid_constant__ launch parameter): the set for the half-cycle step h that the hot
// loop reads straight from the constant bank, and PVDER_FINE_LEVELS sets for steps of h/2, h/4, h/8 used by the
// out-of-line fine-step path (ros_slow): the sub-step that follows a change of the inputs and the PLL pull-in right
// after reset are taken as 2^level shorter steps (pvder_env_config.refine_input_level / startup_level).
struct RodasTab : RodasCoef {
  RodasCoef fine[PVDER_FINE_LEVELS];
};

template <class M, class T>
PVDER_HD void fill_rodas(T& t, const Params& par, double hinv) {
  static_assert(M::N_LUC <= 16, "luc table too small");
#if PVDER_SCHEME == 4
  t.a21 = 0.2000000000000000e+01;
  t.a31 = 0.1867943637803922e+01; t.a32 = 0.2344449711399156e+00;
  const double c41 = -0.2137148994382534e+01, c42 = -0.3214669691237626e+00, c31 = 0.2580708087951457e+01, c32 = 0.6515950076447975e+00;
  t.c21 = -0.7137615036412310e+01 * hinv;
  t.c31 = c31 * hinv; t.c32 = c32 * hinv;
  t.c41 = c41 * hinv; t.c42 = c42 * hinv; t.c43 = -0.6949742501781779e+00 * hinv;
  t.d41 = (c41 - c31) * hinv; t.d42 = (c42 - c32) * hinv;
  t.m1 = 0.2255570073418735e+01; t.m2 = 0.2870493262186792e+00; t.m3 = 0.4353179431840180e+00; t.m4 = 0.1093502252409163e+01;
  t.m34 = t.m3 + t.m4;
#else
  t.a21 = 0.1544000000000000e+01;
  t.a31 = 0.9466785280815826e+00; t.a32 = 0.2557011698983284e+00;
  t.a41 = 0.3314825187068521e+01; t.a42 = 0.2896124015972201e+01; t.a43 = 0.9986419139977817e+00;
  t.a51 = 0.1221224509226641e+01; t.a52 = 0.6019134481288629e+01; t.a53 = 0.1253708332932087e+02;
  t.a54 = -0.6878860361058950e+00;
  t.c21 = -0.5668800000000000e+01 * hinv;
  t.c31 = -0.2430093356833875e+01 * hinv; t.c32 = -0.2063599157091915e+00 * hinv;
  t.c41 = -0.1073529058151375e+00 * hinv; t.c42 = -0.9594562251023355e+01 * hinv; t.c43 = -0.2047028614809616e+02 * hinv;
  t.c51 = 0.7496443313967647e+01 * hinv; t.c52 = -0.1024680431464352e+02 * hinv; t.c53 = -0.3399990352819905e+02 * hinv;
  t.c54 = 0.1170890893206160e+02 * hinv;
  t.c61 = 0.8083246795921522e+01 * hinv; t.c62 = -0.7981132988064893e+01 * hinv; t.c63 = -0.3152159432874371e+02 * hinv;
  t.c64 = 0.1631930543123136e+02 * hinv; t.c65 = -0.6058818238834054e+01 * hinv;
#endif
  t.ghinv = hinv * (1.0 / RG);
  const double hg = 1.0 / t.ghinv;
  t.cs21 = t.c21 * hg;
  t.cs31 = t.c31 * hg; t.cs32 = t.c32 * hg;
  t.cs41 = t.c41 * hg; t.cs42 = t.c42 * hg; t.cs43 = t.c43 * hg;
  t.cs51 = t.c51 * hg; t.cs52 = t.c52 * hg; t.cs53 = t.c53 * hg; t.cs54 = t.c54 * hg;
  t.cs61 = t.c61 * hg; t.cs62 = t.c62 * hg; t.cs63 = t.c63 * hg; t.cs64 = t.c64 * hg; t.cs65 = t.c65 * hg;
  t.ds41 = t.d41 * hg; t.ds42 = t.d42 * hg;
  t.kx = par.Ki_GCC * hg; t.kdc = par.Ki_DC * hg; t.kq = par.Ki_Q * hg; t.kpll = par.Ki_PLL * hg;
  t.du_free = 1.0 / (t.ghinv + par.wp);
  t.du_frz = hg;
  for (int i = 0; i < 16; ++i) t.luc[i] = 0.0;
  M::lu_consts(par, t.ghinv, t.luc);
}

template <class M>
PVDER_HD RodasTab make_rodas_tab(const Params& par, double hinv) {
  RodasTab t = {};
  fill_rodas<M>(t, par, hinv);
  for (int l = 0; l < PVDER_FINE_LEVELS; ++l) fill_rodas<M>(t.fine[l], par, (double)(2 << l) * hinv);
  return t;
}

// Aux record at PLL angle dl and DC voltage V with full-accuracy library functions (once per env
// step, and whenever a stage leaves the incremental range).
PVDER_DEV void aux_exact_sv(const Params& par, const Inputs& in, double dl, double V, Aux& a) {
  sincos(dl, &a.sn, &a.cs);
  a.E = exp(par.kappa * V);
  pov_from_exp(par, in, a.E, a.PoV, a.dPoV);
}

template <class M>
PVDER_DEV void aux_exact(const Params& par, const Inputs& in, const double (&y)[M::NS], Aux& a) {
  aux_exact_sv(par, in, y[M::IDX_DL], y[M::IDX_VDC], a);
}

// Aux record at (dl, V) close to the base point (angle dl0, DC voltage V0, record b):
//   sin/cos(dl0 + d) by rotating (sn0, cs0) through d,  exp(kappa (V0 + dv)) = E0 * exp(kappa dv)
//   -- short Taylor polynomials instead of 2 library calls per stage.  Degrees are chosen against the integrator,
//   not against the ulp: within the range |d|, |kappa dv| < 2^-4 the truncation is below 1e-10 relative (sin:
//   d^7/5040, cos: d^6/720, exp: x^6/720), three orders below the 1e-7 accuracy of a half-cycle step; and because
//   every argument is O(h), the error terms are O(h^6) and beyond -- past the method's own order, so they do not
//   change its convergence.  Outside the range (PLL pull-in right after reset) the step is redone by the EXACT
//   instantiation, kept out of line so the hot loop stays small.
// FULL = false (inner stages): sin, cos, E and the array current Ppv/Vdc the right-hand side reads.  FULL = true (the
// state the step ends in, base point of the next step's Jacobian): sin, cos, E; the PV part is evaluated from E at
// the start of the next step, with the inputs in force then (ros_core).
template <bool EXACT, bool FULL = true, bool SHORT = (PVDER_AUX_SHORT != 0)>
PVDER_DEV void aux_advance_sv(const Params& par, const Inputs& in, const Aux& b, double dl0, double V0, double dl,
                              double V, Aux& a, bool& out_of_range) {
  if (EXACT) {
    aux_exact_sv(par, in, dl, V, a);
    return;
  }
  const double d = dl - dl0;
  const double dv = V - V0;
  const double x = par.kappa * dv;
  // outside the polynomial range the caller discards this step and redoes it with EXACT = true
  out_of_range |= !(fabs(d) < 0.0625 && fabs(x) < 0.0625);
  if (SHORT) {
    // the same polynomials with shorter dependency chains (the side-inputs sit on the critical path of every stage:
    // sin/cos 6 -> 5 levels, exp 7 -> 5 by Estrin's scheme; one multiply more)
    const double d2 = d * d;
    const double d3 = d2 * d;
    const double ps = fma(d2, 1.0 / 120.0, -1.0 / 6.0);
    const double sd = fma(ps, d3, d);                          // sin d  = d - d^3/6 + d^5/120
    const double pc = fma(d2, 1.0 / 24.0, -0.5);
    const double cdm1 = pc * d2;                               // cos d - 1 = -d^2/2 + d^4/24
    a.sn = fma(b.cs, sd, fma(b.sn, cdm1, b.sn));
    a.cs = fma(-b.sn, sd, fma(b.cs, cdm1, b.cs));
    const double x2 = x * x;
    const double pu = fma(x, 0.5, 1.0);
    const double pv = fma(x, 1.0 / 24.0, 1.0 / 6.0);
    const double pw = fma(x2, 1.0 / 120.0, pv);
    const double pe = fma(x2, pw, pu);                         // 1 + x/2 + x^2/6 + x^3/24 + x^4/120
    a.E = fma(b.E * x, pe, b.E);                               // E0 * exp(x), exp to x^5/120
  } else {
    const double d2 = d * d;
    const double ps = fma(d2, 1.0 / 120.0, -1.0 / 6.0);
    const double sd = fma(ps * d2, d, d);                      // sin d  = d - d^3/6 + d^5/120
    const double pc = fma(d2, 1.0 / 24.0, -0.5);
    const double cdm1 = pc * d2;                               // cos d - 1 = -d^2/2 + d^4/24
    a.sn = fma(b.sn, cdm1, fma(b.cs, sd, b.sn));
    a.cs = fma(b.cs, cdm1, fma(-b.sn, sd, b.cs));
    double pe = fma(x, 1.0 / 120.0, 1.0 / 24.0);
    pe = fma(pe, x, 1.0 / 6.0);
    pe = fma(pe, x, 0.5);
    pe = fma(pe, x, 1.0);
    a.E = fma(b.E * pe, x, b.E);                               // E0 * exp(x), exp to x^5/120
  }
  if (FULL) a.PoV = a.dPoV = 0.0;
  else a.PoV = ppv_over_v_from_exp(par, in, a.E);
}

template <class M, bool EXACT, bool FULL = true>
PVDER_DEV void aux_advance(const Params& par, const Inputs& in, const Aux& b, double dl0, double V0,
                           const double (&Y)[M::NS], Aux& a, bool& out_of_range) {
  aux_advance_sv<EXACT, FULL>(par, in, b, dl0, V0, Y[M::IDX_DL], Y[M::IDX_VDC], a, out_of_range);
}

// Effective gains of the freezable rows (bit order of freeze_bits): the parameter, or 0 while the row is
// clamped; the unit-pivot rows (x, xDC, xQ) carry theirs pre-scaled by h*gamma, and gn[NFRZ] is the scaled
// Ki_PLL.  Computed once per sub-step; the generated RHS/Jacobian take them as inputs.
template <class M, class TAB>
PVDER_DEV void make_gains(const Params& par, const TAB& tab, unsigned frz, double (&gn)[M::NGAIN]) {
#pragma unroll
  for (int k = 0; k < M::PHASES; ++k) {
    gn[4 * k] = (frz & (1u << (4 * k))) ? 0.0 : tab.kx;
    gn[4 * k + 1] = (frz & (1u << (4 * k + 1))) ? 0.0 : tab.kx;
    gn[4 * k + 2] = (frz & (1u << (4 * k + 2))) ? 0.0 : par.wp;
    gn[4 * k + 3] = (frz & (1u << (4 * k + 3))) ? 0.0 : par.wp;
  }
  gn[4 * M::PHASES] = (frz & (1u << (4 * M::PHASES))) ? 0.0 : tab.kdc;
  gn[4 * M::PHASES + 1] = (frz & (1u << (4 * M::PHASES + 1))) ? 0.0 : tab.kq;
  gn[M::NFRZ] = tab.kpll;
  // reciprocal pivots of the u rows: 1/(1/(h g) + wp), or h g while the row is clamped
#pragma unroll
  for (int k = 0; k < M::PHASES; ++k) {
    gn[M::NFRZ + 1 + 2 * k] = (frz & (1u << (4 * k + 2))) ? tab.du_frz : tab.du_free;
    gn[M::NFRZ + 2 + 2 * k] = (frz & (1u << (4 * k + 3))) ? tab.du_frz : tab.du_free;
  }
}

#ifndef PVDER_LAZY_GAINS
#define PVDER_LAZY_GAINS 0   // 1: re-derive the effective gains from the clamp bits at every stage (selects on the idle ALU pipe)
#endif                       //    instead of holding 9 doubles across the whole step: no spills left with ROS4-L (1.481 -> 1.468 ms)
PVDER_DEV unsigned opaque_bits(unsigned v) {
#ifdef __CUDACC__
  asm volatile("" : "+r"(v));     // the compiler must treat every call's result as a new value: no CSE, no hoisting
#endif
  return v;
}

#ifndef PVDER_CARRY_LIMIT_FLAGS
#define PVDER_CARRY_LIMIT_FLAGS 0   // 1 (study switch, measured SLOWER): the limit test of the anti-windup clamp for the NEXT sub-step is
#endif                              // evaluated at the end of the step (same basic block as the final side-input advance) instead of at
                                    // the loop head, where its serial chain stands alone (9 % of the stall samples on 3.6 % of the
                                    // instructions).  B200: 1.398 -> 1.441 ms per 1 Mi-env step, 1.442 -> 1.491 ms over a full episode
                                    // (profiles/r2e_carried_limit_flags_ab.txt): at the loop head the test shares vR, vI, m, Q, i_ref
                                    // with stage 1; moved away it computes them twice (+14 FP64 instructions), which costs more
                                    // than the overlap buys.
// Limit flags of the anti-windup clamp (A.3) at state y: |m| > 10 m_limit in some phase, |i_ref| > iref_limit.
template <class M>
PVDER_DEV void limit_flags(const double (&y)[M::NS], const Params& par, const Inputs& in, bool& m_over, bool& i_over);

// One half-cycle step of the scheme PVDER_SCHEME selects.  `base` is the Aux record at y on entry and at the new y on exit.
// Returns false (y, base untouched) when EXACT == false and a stage left the incremental range.
// FREE: no clamp is active (frz == 0 in every lane that takes this instantiation): the effective gains are
// the parameters themselves, read from the constant bank instead of occupying registers.
// next_m / next_i (hot loop only, PVDER_CARRY_LIMIT_FLAGS): receive the clamp's limit flags at the NEW state when the step is
// accepted (untouched otherwise) -- evaluated here so that their serial chain shares a scheduling region with the final
// side-input advance instead of standing alone at the head of the next sub-step.
template <class M, bool EXACT, bool FREE = false, class TAB>
PVDER_DEV bool ros_core(double (&y)[M::NS], const Params& par, const Inputs& in, const TAB& tab,
                           unsigned frz, Aux& base, bool* next_m = nullptr, bool* next_i = nullptr) {
  double gn[M::NGAIN];
  make_gains<M>(par, tab, FREE ? 0u : frz, gn);
#if PVDER_LAZY_GAINS
#define PVDER_WITH_GAINS(stmt) { double gs_[M::NGAIN]; make_gains<M>(par, tab, FREE ? 0u : opaque_bits(frz), gs_); stmt; }
#define PVDER_GN gs_
#else
#define PVDER_WITH_GAINS(stmt) { stmt; }
#define PVDER_GN gn
#endif
  constexpr int NS = M::NS;
  bool oor = false;
  const double dl0 = y[M::IDX_DL], V0 = y[M::IDX_VDC];
  pov_from_exp(par, in, base.E, base.PoV, base.dPoV);      // inputs (insolation) may have changed
  typename M::LU lu;
  M::factor(y, par, in, base, gn, tab.ghinv, tab.luc, lu);
  double K1[NS], K2[NS], K3[NS], K4[NS], Y[NS];
  Aux ax;
  // stage 1
  PVDER_WITH_GAINS(M::rhs(y, par, in, base, PVDER_GN, K1))
  M::solve(lu, tab.luc, K1);
#if PVDER_SCHEME == 4
  // ROS4-L.  The sum of c_ij/h K_j is pre-loaded into K_i (newest term last, so that everything but one FMA is off
  // the critical path) and the right-hand side is accumulated onto it (rhs_acc: no separate adds).
#define PVDER_C(nn) (M::unit_row(i) ? tab.cs##nn : tab.c##nn)
#define PVDER_D(nn) (M::unit_row(i) ? tab.ds##nn : tab.d##nn)
  // stage 2
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    Y[i] = fma(tab.a21, K1[i], y[i]);
    K2[i] = PVDER_C(21) * K1[i];
  }
  aux_advance<M, EXACT, false>(par, in, base, dl0, V0, Y, ax, oor);
  PVDER_WITH_GAINS(M::rhs_acc(Y, par, in, ax, PVDER_GN, K2))
  M::solve(lu, tab.luc, K2);
  // stage 3; K1 and K2 are folded into everything that still needs them (Y3, the pre-loaded sums of stages 3 and 4,
  // the new state) as soon as K2 exists
  double yn[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    Y[i] = fma(tab.a32, K2[i], fma(tab.a31, K1[i], y[i]));
    K3[i] = fma(PVDER_C(32), K2[i], PVDER_C(31) * K1[i]);
    K4[i] = fma(PVDER_D(42), K2[i], PVDER_D(41) * K1[i]);
    yn[i] = fma(tab.m2, K2[i], fma(tab.m1, K1[i], y[i]));
  }
  aux_advance<M, EXACT, false>(par, in, base, dl0, V0, Y, ax, oor);
  PVDER_WITH_GAINS(M::rhs_acc(Y, par, in, ax, PVDER_GN, K3))
  // stage 4 re-uses f(Y3): b_4 = b_3 + sum_j (c_4j - c_3j)/h K_j + c_43/h K_3
#if PVDER_STAGE4_DELTA
  // W K_3 = b_3, so W (K_4 - K_3) = sum_j (c_4j - c_3j)/h K_j + c_43/h K_3: the fourth solve yields K_4 - K_3 and b_3 is
  // never added (11 DADD fewer, and b_3 need not survive the third solve);  y+ = yn + (m_3 + m_4) K_3 + m_4 (K_4 - K_3)
  M::solve(lu, tab.luc, K3);
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    K4[i] = fma(PVDER_C(43), K3[i], K4[i]);
    yn[i] = fma(tab.m34, K3[i], yn[i]);
  }
#else
#pragma unroll
  for (int i = 0; i < NS; ++i) K4[i] += K3[i];
  M::solve(lu, tab.luc, K3);
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    K4[i] = fma(PVDER_C(43), K3[i], K4[i]);
    yn[i] = fma(tab.m3, K3[i], yn[i]);
  }
#endif
#undef PVDER_C
#undef PVDER_D
  M::solve(lu, tab.luc, K4);
#pragma unroll
  for (int i = 0; i < NS; ++i) Y[i] = fma(tab.m4, K4[i], yn[i]);
  aux_advance<M, EXACT>(par, in, base, dl0, V0, Y, ax, oor);
#else   // Rodas4
  double K5[NS];
  // Stages 2-4: the sum of c_ij/h K_j is pre-loaded into K_i (newest term last, so that everything but one FMA
  // is off the critical path) and the right-hand side is accumulated onto it (rhs_acc: no separate adds).
#define PVDER_C(nn) (M::unit_row(i) ? tab.cs##nn : tab.c##nn)
  // stage 2
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    Y[i] = fma(tab.a21, K1[i], y[i]);
    K2[i] = PVDER_C(21) * K1[i];
  }
  aux_advance<M, EXACT, false>(par, in, base, dl0, V0, Y, ax, oor);
  PVDER_WITH_GAINS(M::rhs_acc(Y, par, in, ax, PVDER_GN, K2))
  M::solve(lu, tab.luc, K2);
  // stage 3
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    Y[i] = fma(tab.a32, K2[i], fma(tab.a31, K1[i], y[i]));
    K3[i] = fma(PVDER_C(32), K2[i], PVDER_C(31) * K1[i]);
  }
  aux_advance<M, EXACT, false>(par, in, base, dl0, V0, Y, ax, oor);
  PVDER_WITH_GAINS(M::rhs_acc(Y, par, in, ax, PVDER_GN, K3))
  M::solve(lu, tab.luc, K3);
  // stage 4
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    Y[i] = fma(tab.a43, K3[i], fma(tab.a42, K2[i], fma(tab.a41, K1[i], y[i])));
    K4[i] = fma(PVDER_C(43), K3[i], fma(PVDER_C(42), K2[i], PVDER_C(41) * K1[i]));
  }
  aux_advance<M, EXACT, false>(par, in, base, dl0, V0, Y, ax, oor);
  PVDER_WITH_GAINS(M::rhs_acc(Y, par, in, ax, PVDER_GN, K4))
#undef PVDER_C
  M::solve(lu, tab.luc, K4);
#if PVDER_FOLD
  // K1..K4 are folded into the stage-5/6 sums as soon as K4 exists (same FMAs, done early): three
  // vectors (Y5, C5, C6) instead of five stay live through the last two stages.
  double C6[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    Y[i] = fma(tab.a54, K4[i], fma(tab.a53, K3[i], fma(tab.a52, K2[i], fma(tab.a51, K1[i], y[i]))));
    C6[i] = fma((M::unit_row(i) ? tab.cs64 : tab.c64), K4[i], fma((M::unit_row(i) ? tab.cs63 : tab.c63), K3[i], fma((M::unit_row(i) ? tab.cs62 : tab.c62), K2[i], (M::unit_row(i) ? tab.cs61 : tab.c61) * K1[i])));
    K5[i] = fma((M::unit_row(i) ? tab.cs54 : tab.c54), K4[i], fma((M::unit_row(i) ? tab.cs53 : tab.c53), K3[i], fma((M::unit_row(i) ? tab.cs52 : tab.c52), K2[i], (M::unit_row(i) ? tab.cs51 : tab.c51) * K1[i])));
  }
  // stage 5: the right-hand side is accumulated onto the pre-loaded sum (rhs_acc folds the addend into each
  // row's last multiply: no separate adds)
  aux_advance<M, EXACT, false>(par, in, base, dl0, V0, Y, ax, oor);
  PVDER_WITH_GAINS(M::rhs_acc(Y, par, in, ax, PVDER_GN, K5))
  M::solve(lu, tab.luc, K5);
  // stage 6 (Y6 = Y5 + K5; y+ = Y6 + K6: stiffly accurate)
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    Y[i] += K5[i];
    C6[i] = fma((M::unit_row(i) ? tab.cs65 : tab.c65), K5[i], C6[i]);
  }
  aux_advance<M, EXACT, false>(par, in, base, dl0, V0, Y, ax, oor);
  PVDER_WITH_GAINS(M::rhs_acc(Y, par, in, ax, PVDER_GN, C6))
  M::solve(lu, tab.luc, C6);
#pragma unroll
  for (int i = 0; i < NS; ++i) Y[i] += C6[i];
  aux_advance<M, EXACT>(par, in, base, dl0, V0, Y, ax, oor);
#else
  // stage 5
#pragma unroll
  for (int i = 0; i < NS; ++i)
    Y[i] = fma(tab.a54, K4[i], fma(tab.a53, K3[i], fma(tab.a52, K2[i], fma(tab.a51, K1[i], y[i]))));
  aux_advance<M, EXACT, false>(par, in, base, dl0, V0, Y, ax, oor);
  M::rhs(Y, par, in, ax, gn, K5);
#pragma unroll
  for (int i = 0; i < NS; ++i)
    K5[i] = fma((M::unit_row(i) ? tab.cs54 : tab.c54), K4[i], fma((M::unit_row(i) ? tab.cs53 : tab.c53), K3[i], fma((M::unit_row(i) ? tab.cs52 : tab.c52), K2[i], fma((M::unit_row(i) ? tab.cs51 : tab.c51), K1[i], K5[i]))));
  M::solve(lu, tab.luc, K5);
  // stage 6 (Y6 = Y5 + K5; y+ = Y6 + K6: stiffly accurate)
#pragma unroll
  for (int i = 0; i < NS; ++i) Y[i] += K5[i];
  aux_advance<M, EXACT, false>(par, in, base, dl0, V0, Y, ax, oor);
  double K6[NS];
  M::rhs(Y, par, in, ax, gn, K6);
#pragma unroll
  for (int i = 0; i < NS; ++i)
    K6[i] = fma((M::unit_row(i) ? tab.cs65 : tab.c65), K5[i], fma((M::unit_row(i) ? tab.cs64 : tab.c64), K4[i], fma((M::unit_row(i) ? tab.cs63 : tab.c63), K3[i], fma((M::unit_row(i) ? tab.cs62 : tab.c62), K2[i], fma((M::unit_row(i) ? tab.cs61 : tab.c61), K1[i], K6[i])))));
  M::solve(lu, tab.luc, K6);
#pragma unroll
  for (int i = 0; i < NS; ++i) Y[i] += K6[i];
  aux_advance<M, EXACT>(par, in, base, dl0, V0, Y, ax, oor);
#endif
#endif
  bool nm = false, ni = false;
  if (next_m) limit_flags<M>(Y, par, in, nm, ni);
  if (oor) return false;
#pragma unroll
  for (int i = 0; i < NS; ++i) y[i] = Y[i];
  base = ax;
  if (next_m) { *next_m = nm; *next_i = ni; }
  return true;
}
#undef PVDER_WITH_GAINS
#undef PVDER_GN

// Out-of-line slow path.  Everything travels by value, so that nothing in the caller's hot loop has its address taken
// (that would pin the state and the aux record to local memory: 18 doubles stored and re-loaded per sub-step).
//   level == 0: the half-cycle step redone with library transcendentals at every stage (a stage left the range of the
//               incremental side-inputs);
//   level  > 0: the sub-step taken as 2^level steps of h / 2^level with the coefficient set tab->fine[level - 1] (read
//               through the pointer: this path is cold, the hot loop keeps its constant-bank operands).  The clamp mode is
//               re-sampled before every fine step, like before every other integrator step.
template <class M>
struct StepState {   // <= 128 bytes for the single-phase models: travels in registers both ways
  double y[M::NS];
  double sn, cs, E;    // the aux record without its PV part (ros_core re-derives that from E and the inputs)
  int exact;           // steps that needed library transcendentals
  int flags;           // bit 0: an anti-windup clamp was active in one of the fine steps, bit 1: duty-cycle limit exceeded
};
template <class M>
PVDER_DEV unsigned freeze_bits(const double (&y)[M::NS], const Params& par, const Inputs& in, bool& m_over_out);
#ifndef PVDER_SLOW_INLINE
#define PVDER_SLOW_INLINE 1   // 1: the slow path is inlined into the (cold) segment loop, so its fine-step tables are constant-bank
#endif                        //    operands too (B200: 1.569 -> 1.510 ms per 1 Mi-env step against the noinline function, whose
                              //    generic loads from the parameter space made a fine step cost several hot steps)
#if PVDER_SLOW_INLINE
#define PVDER_SLOW_LINKAGE PVDER_DEV
#else
#define PVDER_SLOW_LINKAGE PVDER_NOINLINE
#endif
template <class M>
PVDER_SLOW_LINKAGE StepState<M> ros_slow(StepState<M> s, const Params* par, Inputs in, const RodasTab* tab, unsigned frz,
                                         int level) {
  s.exact = 0;
  s.flags = 0;
  Aux base;
  base.sn = s.sn; base.cs = s.cs; base.E = s.E;
  base.PoV = base.dPoV = 0.0;
  if (level <= 0) {
    ros_core<M, true>(s.y, *par, in, static_cast<const RodasCoef&>(*tab), frz, base);
    s.exact = 1;
    s.sn = base.sn; s.cs = base.cs; s.E = base.E;
    return s;
  }
  const RodasCoef& ft = tab->fine[level - 1];
  const int nf = 1 << level;
#pragma unroll 1
  for (int j = 0; j < nf; ++j) {
    if (j) {
      bool m_over;
      frz = freeze_bits<M>(s.y, *par, in, m_over);
      s.flags |= (frz != 0u ? 1 : 0) | (m_over ? 2 : 0);
    }
    if (!ros_core<M, false>(s.y, *par, in, ft, frz, base)) {
      ros_core<M, true>(s.y, *par, in, ft, frz, base);
      s.exact += 1;
    }
  }
  s.sn = base.sn; s.cs = base.cs; s.E = base.E;
  return s;
}

// One half-cycle sub-step of the hot loop: the inlined step.  Returns false (y, base untouched) when a stage left the
// incremental range: the caller leaves the hot loop and redoes the sub-step out of line (ros_slow_substep, level 0) --
// the hot loop itself contains no call, so nothing in it is bound by the calling convention's register classes.
template <class M, class TAB>
PVDER_DEV bool ros_step(double (&y)[M::NS], const Params& par, const Inputs& in, const TAB& tab,
                        unsigned frz, Aux& base, bool* next_m = nullptr, bool* next_i = nullptr) {
  // One instantiation serves clamped and unclamped envs (the clamp enters through the per-row
  // effective gains): with a random policy two thirds of the warps hold a clamped lane late in an
  // episode, so a separate divergent code path for them cost 2x there.
#if PVDER_FREE_PATH
  // warp-uniform choice between the clamp-free instantiation and the general one (never both in a warp)
#ifdef __CUDACC__
  const bool any_frz = __any_sync(__activemask(), frz != 0u) != 0;
#else
  const bool any_frz = frz != 0u;
#endif
  return any_frz ? ros_core<M, false, false>(y, par, in, tab, frz, base, next_m, next_i)
                 : ros_core<M, false, true>(y, par, in, tab, frz, base, next_m, next_i);
#else
  return ros_core<M, false>(y, par, in, tab, frz, base, next_m, next_i);
#endif
}

// One half-cycle sub-step through the out-of-line path: level 0 = the step redone with library transcendentals, level > 0
// = refined into 2^level fine steps.  Returns the steps that used library transcendentals; flags: see StepState.
template <class M, class TAB>
PVDER_DEV int ros_slow_substep(double (&y)[M::NS], const Params& par, const Inputs& in, const TAB& tab,
                               unsigned frz, Aux& base, int level, int& flags) {
  StepState<M> s;
#pragma unroll
  for (int i = 0; i < M::NS; ++i) s.y[i] = y[i];
  s.sn = base.sn; s.cs = base.cs; s.E = base.E;
  s = ros_slow<M>(s, &par, in, &tab, frz, level);
#pragma unroll
  for (int i = 0; i < M::NS; ++i) y[i] = s.y[i];
  base.sn = s.sn; base.cs = s.cs; base.E = s.E;
  flags |= s.flags;
  return s.exact;
}

// pvder's clamping test np.sign(a) == np.sign(b)
PVDER_DEV bool same_sign(double a, double b) {
  const int sa = (a > 0.0) - (a < 0.0);
  const int sb = (b > 0.0) - (b < 0.0);
  return sa == sb;
}

// The same test on the bit patterns (integer pipe only, no FP64 compare): equal iff both are zero (+-0), or neither is and
// the sign bits agree.  Identical to same_sign for every non-NaN argument (a NaN -- an env that is about to be reported
// NONFINITE -- counts as a signed non-zero here, as a zero there).
PVDER_DEV bool same_sign_bits(double a, double b) {
#ifdef __CUDACC__
  const int ha = __double2hiint(a), hb = __double2hiint(b);
  const unsigned la = (unsigned)__double2loint(a), lb = (unsigned)__double2loint(b);
#else
  unsigned long long ua, ub;
  std::memcpy(&ua, &a, 8);
  std::memcpy(&ub, &b, 8);
  const int ha = (int)(ua >> 32), hb = (int)(ub >> 32);
  const unsigned la = (unsigned)ua, lb = (unsigned)ub;
#endif
  const bool za = ((((unsigned)ha) << 1) | la) == 0u, zb = ((((unsigned)hb) << 1) | lb) == 0u;
  return (za == zb) && (za || ((ha ^ hb) >= 0));
}

#ifndef PVDER_FREEZE_BRANCHFREE
#define PVDER_FREEZE_BRANCHFREE 0   // 1: freeze_bits without a branch (study switch, measured SLOWER: see there)
#endif

PVDER_DEV void phase_rot(int P, int k, double& rr, double& ri) {
  if (P == 1 || k == 0) { rr = 1.0; ri = 0.0; }
  else if (k == 1) { rr = -0.5; ri = -0.86602540378443864676; }
  else { rr = -0.5; ri = 0.86602540378443864676; }
}

// Anti-windup mode, sampled once per half-cycle sub-step.  Bit order = M::NFRZ
// rows: per phase xR,xI,uR,uI ; then xDC, xQ.
// MODE 0: flags and rows (freeze_bits); 1: flags only (limit_flags: returns 0); 2: rows for GIVEN flags (freeze_rows).
template <class M, int MODE>
PVDER_DEV unsigned freeze_impl(const double (&y)[M::NS], const Params& par, const Inputs& in, bool& m_over_io, bool& i_over_io) {
  constexpr int P = M::PHASES;
  constexpr int B = 6 * P;
  // Same expression trees as the generated right-hand side (vR, vI, qs, Qp, iref): when the stepper is inlined
  // after this function the compiler shares them with stage 1 instead of computing them twice.
  double Qs = 0.0;
  bool m_over = false;
#pragma unroll
  for (int k = 0; k < P; ++k) {
    double rr, ri;
    phase_rot(P, k, rr, ri);
    const double iR = y[6 * k], iI = y[6 * k + 1];
    const double mR = fma(par.Kp_GCC, y[6 * k + 4], y[6 * k + 2]);
    const double mI = fma(par.Kp_GCC, y[6 * k + 5], y[6 * k + 3]);
    if (MODE != 2) m_over |= (mR * mR + mI * mI) > par.m_limit10 * par.m_limit10;
    const double vgk = M::BALANCED3 ? in.vg : vg_of_phase(in, P, k);
    double vR, vI;
    if (P == 1 || k == 0) {
      vR = fma(par.Rt, iR, fma(-par.Xt, iI, vgk));
      vI = fma(par.Xt, iR, par.Rt * iI);
    } else {
      vR = fma(par.Rt, iR, fma(-par.Xt, iI, rr * vgk));
      vI = fma(par.Xt, iR, fma(par.Rt, iI, ri * vgk));
    }
    const double qs = fma(vI, iR, -(vR * iI));
    Qs = (k == 0) ? qs : Qs + qs;
  }
  if (MODE == 2) m_over = m_over_io;
  else m_over_io = m_over;
  const double Vdc = y[B], xDC = y[B + 1], xQ = y[B + 2];
  const double dV = in.Vdcref - Vdc;
  const double dQ = fma(-(0.5 * M::PMULT), Qs, in.Qref);      // Qref - Q
  const double irefR = fma(par.Kp_DC, dV, xDC);
  const double irefI = fma(-par.Kp_Q, dQ, xQ);
  const bool i_over = (MODE == 2) ? i_over_io : ((irefR * irefR + irefI * irefI) > par.iref_limit * par.iref_limit);
  if (MODE != 2) i_over_io = i_over;
  if (MODE == 1) return 0u;
#if PVDER_FREEZE_BRANCHFREE
  // Branch-free form (study switch, OFF).  The early return below makes the limit test -- a serial chain of ~15 dependent
  // FP64 operations -- a basic block of its own at the head of the hot loop, with nothing to overlap it: 9 % of the loop's
  // stall samples sit on 3.6 % of its instructions (ncu source page of the r2d capture).  This form evaluates every row test
  // (sign tests on the integer pipe, +8 executed FP64 instructions per sub-step) and lets the two limit flags select, so
  // the whole sub-step is one scheduling region.  MEASURED on B200 (profiles/r2e_branchfree_clamp_ab.txt): 1.397 -> 1.447 ms
  // per 1 Mi-env step, 1.440 -> 1.477 ms over a full episode -- ptxas does not use the freedom, the extra work costs more.
  {
    unsigned mb = 0u, ib = 0u;
#pragma unroll
    for (int k = 0; k < P; ++k) {
      double rr, ri;
      phase_rot(P, k, rr, ri);
      const double uR = y[6 * k + 4], uI = y[6 * k + 5];
      const double duR = par.wp * (-uR + (rr * irefR - ri * irefI) - y[6 * k]);
      const double duI = par.wp * (-uI + (ri * irefR + rr * irefI) - y[6 * k + 1]);
      mb |= same_sign_bits(par.Ki_GCC * uR, y[6 * k + 2]) ? (1u << (4 * k)) : 0u;
      mb |= same_sign_bits(par.Ki_GCC * uI, y[6 * k + 3]) ? (1u << (4 * k + 1)) : 0u;
      mb |= same_sign_bits(duR, uR) ? (1u << (4 * k + 2)) : 0u;
      mb |= same_sign_bits(duI, uI) ? (1u << (4 * k + 3)) : 0u;
    }
    ib |= same_sign_bits(par.Ki_DC * dV, xDC) ? (1u << (4 * P)) : 0u;
    ib |= same_sign_bits(-par.Ki_Q * dQ, xQ) ? (1u << (4 * P + 1)) : 0u;
    return (m_over ? mb : 0u) | (i_over ? ib : 0u);
  }
#endif
  if (!(m_over || i_over)) return 0u;
  unsigned bits = 0u;
  if (m_over) {
#pragma unroll
    for (int k = 0; k < P; ++k) {
      double rr, ri;
      phase_rot(P, k, rr, ri);
      const double uR = y[6 * k + 4], uI = y[6 * k + 5];
      const double duR = par.wp * (-uR + (rr * irefR - ri * irefI) - y[6 * k]);
      const double duI = par.wp * (-uI + (ri * irefR + rr * irefI) - y[6 * k + 1]);
      if (same_sign(par.Ki_GCC * uR, y[6 * k + 2])) bits |= 1u << (4 * k);
      if (same_sign(par.Ki_GCC * uI, y[6 * k + 3])) bits |= 1u << (4 * k + 1);
      if (same_sign(duR, uR)) bits |= 1u << (4 * k + 2);
      if (same_sign(duI, uI)) bits |= 1u << (4 * k + 3);
    }
  }
  if (i_over) {
    if (same_sign(par.Ki_DC * dV, xDC)) bits |= 1u << (4 * P);
    if (same_sign(-par.Ki_Q * dQ, xQ)) bits |= 1u << (4 * P + 1);
  }
  return bits;
}

template <class M>
PVDER_DEV unsigned freeze_bits(const double (&y)[M::NS], const Params& par, const Inputs& in, bool& m_over_out) {
  bool i_over = false;
  return freeze_impl<M, 0>(y, par, in, m_over_out, i_over);
}
template <class M>
PVDER_DEV void limit_flags(const double (&y)[M::NS], const Params& par, const Inputs& in, bool& m_over, bool& i_over) {
  freeze_impl<M, 1>(y, par, in, m_over, i_over);
}
// the clamped rows for limit flags evaluated earlier at the same state (hot loop, PVDER_CARRY_LIMIT_FLAGS)
template <class M>
PVDER_DEV unsigned freeze_rows(const double (&y)[M::NS], const Params& par, const Inputs& in, bool m_over, bool i_over) {
  return freeze_impl<M, 2>(y, par, in, m_over, i_over);
}

// Event j of (env, episode): which quantity changes and its new value (A.8).  Philox counter
// (env, episode, j, STREAM_EVENTS), key = seed.  Arithmetic uses explicit round-to-nearest ops so
// that the numpy twin reproduces every bit.
PVDER_DEV void draw_event(const pvder_env_config& cfg, uint32_t env, uint32_t episode,
                                           uint32_t j, double& Vgrid, double& Sinsol) {
  uint32_t r[4];
  philox4x32_10(env, episode, j, STREAM_EVENTS, (uint32_t)cfg.seed, (uint32_t)(cfg.seed >> 32), r);
  bool voltage;
  if (cfg.ev_voltage_enable && cfg.ev_insol_enable) voltage = (r[0] >> 31) != 0u;   // random.choice
  else voltage = cfg.ev_voltage_enable != 0;
  const double u = u53(r[1], r[2]);
  if (voltage) Vgrid = __dadd_rn(cfg.ev_v_min, __dmul_rn(__dadd_rn(cfg.ev_v_max, -cfg.ev_v_min), u));
  else Sinsol = __dadd_rn(cfg.ev_s_min, __dmul_rn(__dadd_rn(cfg.ev_s_max, -cfg.ev_s_min), u));
}

PVDER_DEV void apply_event(const pvder_env_config& cfg, const double* vtab, const double* stab,
                                            int64_t ld, int64_t e, uint32_t env_glob, uint32_t episode, int j,
                                            double& Vgrid, double& Sinsol) {
  if (cfg.event_mode == PVDER_EVENTS_PHILOX) {
    if (cfg.ev_voltage_enable || cfg.ev_insol_enable) draw_event(cfg, env_glob, episode, (uint32_t)j, Vgrid, Sinsol);
  } else if (cfg.event_mode == PVDER_EVENTS_TABLE) {
    Vgrid = vtab[(int64_t)j * ld + e];
    Sinsol = stab[(int64_t)j * ld + e];
  }
}

// Algebraic outputs at the current state with the event values in force (PVDER_env.py:531-542,
// :231-301).  Every operation that feeds the discrete reward is an explicit _rn op in a fixed
// order (contract shared with oracle/twin.py::outputs_twin) -> integer outputs are bit-exact.
struct Outputs {
  double obs[PVDER_OBS_DIM];
  double reward;
  int reward_i;
};

// Balanced three-phase set carried by phase a -> full 23-state vector (phases b, c are phase a
// rotated by -/+120 degrees).  Individually rounded ops: the stored state and the outputs computed
// from it are reproducible bit for bit.
PVDER_DEV void expand_balanced(const double (&y)[11], double (&z)[23]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double rr, ri;
    phase_rot(3, k, rr, ri);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double a = y[2 * j], b = y[2 * j + 1];
      if (k == 0) {
        z[2 * j] = a;
        z[2 * j + 1] = b;
      } else {
        z[6 * k + 2 * j] = __dadd_rn(__dmul_rn(a, rr), -__dmul_rn(b, ri));
        z[6 * k + 2 * j + 1] = __dadd_rn(__dmul_rn(a, ri), __dmul_rn(b, rr));
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 5; ++j) z[18 + j] = y[6 + j];
}

// True when phases b, c of a stored 23-state vector are phase a rotated by -/+120 degrees (to
// rounding): the state can then be integrated on phase a alone (Model3phBal).
PVDER_DEV bool is_balanced(const double (&z)[23]) {
  bool ok = true;
#pragma unroll
  for (int k = 1; k < 3; ++k) {
    double rr, ri;
    phase_rot(3, k, rr, ri);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double a = z[2 * j], b = z[2 * j + 1];
      const double tol = 1e-12 * (fabs(a) + fabs(b)) + 1e-300;
      ok &= fabs(z[6 * k + 2 * j] - (a * rr - b * ri)) <= tol;
      ok &= fabs(z[6 * k + 2 * j + 1] - (a * ri + b * rr)) <= tol;
    }
  }
  return ok;
}

// Tail of the output computation from the phase sums (shared by the one-thread and the
// three-lane kernels, so both produce the same bits from the same sums).
template <int P>
PVDER_DEV void finish_outputs(const pvder_env_config& cfg, const Inputs& in, double iaR, double iaI, double vaR,
                              double vaI, double Ppcc, double Qpcc, double v2, double Vdc, double Qref,
                              double Vdcref, int k, Outputs& o) {
  const Params& par = cfg.par;
  const double SQRT2 = 1.4142135623730951;
  const double Vrms = (P == 1) ? __ddiv_rn(__dsqrt_rn(v2), SQRT2) : __ddiv_rn(__dsqrt_rn(__ddiv_rn(v2, 3.0)), SQRT2);
  double Ppv, dPpv;
  ppv_eval(par, in, Vdc, Ppv, dPpv);
  o.obs[0] = iaR; o.obs[1] = iaI; o.obs[2] = vaR; o.obs[3] = vaI; o.obs[4] = Ppcc; o.obs[5] = Qpcc;
  o.obs[6] = Vdc; o.obs[7] = Ppv; o.obs[8] = Vdcref; o.obs[9] = Qref;
  o.obs[10] = __ddiv_rn(__ddiv_rn((double)k, cfg.substeps_per_sec), cfg.max_sim_time);
  // PVDER_env.py:249-299: the reward is the sum over the goal's reward list (`my_spec`: the goal's required term, plus
  // optional ones) of one term per entry, Python's sum() starting from 0.
  double rsum = 0.0;
  int isum = 0;
#pragma unroll
  for (int t = 0; t < PVDER_MAX_REWARD_TERMS; ++t) {
    const int id = cfg.reward_terms[t];
    if (id < 0) break;
    double x, target, lo = 0.01, hi = 0.05;
    if (id == PVDER_TERM_VOLTAGE) { x = Vrms; target = par.Vrms_ref; }                         // :276-287
    else if (id == PVDER_TERM_Q) {                                                             // :263-275, target :238-241
      x = Qpcc;
      target = (cfg.goal == PVDER_GOAL_VOLTAGE) ? Qref : par.q_target;
      if (cfg.discrete_reward && target == 0.0) target = 1e-6;
    } else if (id == PVDER_TERM_POWER) { x = Ppcc; target = par.p_target; hi = 0.03; }         // :290-299
    else { x = Vdc; target = Vdcref; lo = 0.02; }                                              // :251-262 (Vdc_error)
    if (cfg.discrete_reward) {
      const double err = __ddiv_rn(fabs(__dadd_rn(x, -target)), fabs(target));
      isum += (err <= lo) ? 1 : ((err >= hi) ? -5 : -1);
    } else {
      const double d = __dadd_rn(x, -target);
      rsum = __dadd_rn(rsum, -__dmul_rn(d, d));
    }
  }
  o.reward_i = cfg.discrete_reward ? isum : 0;
  o.reward = cfg.discrete_reward ? (double)isum : rsum;
}

template <int P, bool UNBAL = true>
PVDER_DEV void compute_outputs_p(const pvder_env_config& cfg, const double (&y)[6 * P + 5], double Qref,
                                 double Vdcref, double Vgrid, double Sinsol, int k, Outputs& o) {
  constexpr int B = 6 * P;
  const Params& par = cfg.par;
  const Inputs in = make_inputs(cfg, Vgrid, Qref, Vdcref, Sinsol);
  double Ppcc = 0.0, Qpcc = 0.0, v2 = 0.0, vaR = 0.0, vaI = 0.0;
#pragma unroll
  for (int ph = 0; ph < P; ++ph) {
    double rr, ri;
    phase_rot(P, ph, rr, ri);
    const double jR = y[6 * ph], jI = y[6 * ph + 1];
    const double vg = UNBAL ? vg_of_phase(in, P, ph) : in.vg;   // the balanced reduction carries a balanced grid
    const double vkR = __dadd_rn(__dmul_rn(vg, rr), __dadd_rn(__dmul_rn(par.Rt, jR), -__dmul_rn(par.Xt, jI)));
    const double vkI = __dadd_rn(__dmul_rn(vg, ri), __dadd_rn(__dmul_rn(par.Xt, jR), __dmul_rn(par.Rt, jI)));
    Ppcc = __dadd_rn(Ppcc, __dmul_rn(0.5, __dadd_rn(__dmul_rn(vkR, jR), __dmul_rn(vkI, jI))));
    Qpcc = __dadd_rn(Qpcc, __dmul_rn(0.5, __dadd_rn(__dmul_rn(vkI, jR), -__dmul_rn(vkR, jI))));
    v2 = __dadd_rn(v2, __dadd_rn(__dmul_rn(vkR, vkR), __dmul_rn(vkI, vkI)));
    if (ph == 0) { vaR = vkR; vaI = vkI; }
  }
  finish_outputs<P>(cfg, in, y[0], y[1], vaR, vaI, Ppcc, Qpcc, v2, y[B], Qref, Vdcref, k, o);
}

template <class M>
PVDER_DEV void compute_outputs(const pvder_env_config& cfg, const double (&y)[M::NS], double Qref, double Vdcref,
                               double Vgrid, double Sinsol, int k, Outputs& o) {
  if constexpr (M::BALANCED3) {
    double z[23];
    expand_balanced(y, z);
    compute_outputs_p<3, false>(cfg, z, Qref, Vdcref, Vgrid, Sinsol, k, o);
  } else {
    compute_outputs_p<M::PHASES>(cfg, y, Qref, Vdcref, Vgrid, Sinsol, k, o);
  }
}

// Stored state (sd rows 0..NS_STORE-1) <-> registers.  The balanced model reads phase a and the
// shared rows only and writes all 23 rows back.
template <class M>
PVDER_DEV void load_state(const double* sd, int64_t ld, int64_t e, double (&y)[M::NS]) {
  if constexpr (M::BALANCED3) {
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] = sd[(int64_t)i * ld + e];
#pragma unroll
    for (int i = 0; i < 5; ++i) y[6 + i] = sd[(int64_t)(18 + i) * ld + e];
  } else {
#pragma unroll
    for (int i = 0; i < M::NS; ++i) y[i] = sd[(int64_t)i * ld + e];
  }
}

template <class M>
PVDER_DEV void store_state(double* sd, int64_t ld, int64_t e, const double (&y)[M::NS]) {
  if constexpr (M::BALANCED3) {
    double z[23];
    expand_balanced(y, z);
#pragma unroll
    for (int i = 0; i < 23; ++i) sd[(int64_t)i * ld + e] = z[i];
  } else {
#pragma unroll
    for (int i = 0; i < M::NS; ++i) sd[(int64_t)i * ld + e] = y[i];
  }
}


// Trajectory recording (what the reference's SimulationResults plots, PVDER_env.py:358-364): the
// stored-layout state after half-cycle sub-step s of this env step, then the event values that were in
// force during it.  traj is this env's column of a [n_sub_per_step][NS_STORE + 2][traj_ld] array.
template <class M>
PVDER_DEV void record_substep(double* traj, int64_t traj_ld, int s, const double (&y)[M::NS], double Vgrid, double Sinsol) {
  constexpr int ROWS = M::NS_STORE + 2;
  double* row = traj + (int64_t)s * ROWS * traj_ld;
  if constexpr (M::BALANCED3) {
    double z[23];
    expand_balanced(y, z);
#pragma unroll
    for (int i = 0; i < 23; ++i) row[(int64_t)i * traj_ld] = z[i];
  } else {
#pragma unroll
    for (int i = 0; i < M::NS; ++i) row[(int64_t)i * traj_ld] = y[i];
  }
  row[(int64_t)M::NS_STORE * traj_ld] = Vgrid;
  row[(int64_t)(M::NS_STORE + 1) * traj_ld] = Sinsol;
}

// Registers of one environment.
template <class M>
struct EnvRegs {
  double y[M::NS];
  double Qref, Vdcref, Vgrid, Sinsol, ret, last_reward;
  int k, steps, episode, status, done, windup, exact;
};

template <class M>
PVDER_DEV void init_env(const pvder_env_config& cfg, double (&y)[M::NS], double& Qref, double& Vdcref,
                        double& Vgrid, double& Sinsol) {
  if constexpr (M::BALANCED3) {
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] = cfg.y0[i];
#pragma unroll
    for (int i = 0; i < 5; ++i) y[6 + i] = cfg.y0[18 + i];
  } else {
#pragma unroll
    for (int i = 0; i < M::NS; ++i) y[i] = cfg.y0[i];
  }
  Qref = cfg.Q_ref0;
  Vdcref = cfg.Vdc_ref0;
  Vgrid = 1.0;
  Sinsol = 100.0;
}

// One env step for one env (everything between the state load and the state store).
// Returns true when the env advanced (state must be written back).  hist_inc: action whose
// histogram counter must be incremented (-1: none); hist_clear: auto-reset happened.
template <class M, bool RECORD = true>
PVDER_DEV bool advance_env(const pvder_env_config& cfg, const RodasTab& tab, EnvRegs<M>& r, int act, bool active, const double* vtab,
                           const double* stab, int64_t ld, int64_t e, uint32_t env_glob, Outputs& o, int& done_out,
                           int& hist_inc, bool& hist_clear, double* traj = nullptr, int64_t traj_ld = 0) {
  constexpr int NS = M::NS;
  const Params& par = cfg.par;
  hist_inc = -1;
  hist_clear = false;
  bool run = active && !r.done;                     // PVDER_env.py:145-154: step after done is a no-op
  if (run && (unsigned)act >= (unsigned)PVDER_N_ACTIONS) {   // PVDER_env.py:201
    r.status = PVDER_STATUS_BAD_ACTION;
    run = false;
  } else if (run && r.status == PVDER_STATUS_BAD_ACTION) {
    r.status = PVDER_STATUS_OK;      // the rejected action changed nothing: the next valid step clears the flag
  }
  if (run) {
    hist_inc = act;                                          // env_utilities.py:25-30
    r.steps += 1;                                            // PVDER_env.py:156
    const double dQ = (act == 1) ? cfg.delQ_pu : ((act == 2) ? -cfg.delQ_pu : 0.0);
    const double dV = (act == 3) ? cfg.delVdc_pu : ((act == 4) ? -cfg.delVdc_pu : 0.0);
    r.Qref = __dadd_rn(r.Qref, dQ);                          // PVDER_env.py:225
    r.Vdcref = __dadd_rn(r.Vdcref, dV);                      // PVDER_env.py:229
    Aux base;
    Inputs in = make_inputs(cfg, r.Vgrid, r.Qref, r.Vdcref, r.Sinsol);   // changes only when an event fires
    aux_exact<M>(par, in, r.y, base);         // library sincos/exp/div once per env step, then incremental
    // The sub-steps of this env step run in SEGMENTS: a refined sub-step (fine-step level > 0: the inputs changed at
    // its start -- an event instant, or with refine_on_action an action that moved a reference --, or the PLL pull-in of
    // the first startup_substeps after reset, or base_level > 0) goes through the out-of-line path; everything up to the
    // next event instant / the end of the env step is one run of the hot loop, which carries two integers and contains
    // no call (a sub-step whose stages leave the incremental range ends the run and is redone out of line) -- the
    // stepper needs every other register.  In the hot loop the clamp mode is sampled
    // right before every step, in the same basic block as the step itself: freeze_bits uses the expression trees of
    // the generated right-hand side, so stage 1 reuses its vR, vI, m, Q, iref instead of recomputing them.
    const int k0 = r.k, k_end = r.k + cfg.n_sub_per_step;
    constexpr int NO_EVENT = 1 << 30;
    int ev_left;   // sub-steps until the next event instant (it fires at the END of a sub-step)
    {
      const int j_next = (r.k < cfg.ev_start_k) ? 0 : (r.k - cfg.ev_start_k) / cfg.ev_step_k + 1;
      ev_left = (j_next < cfg.ev_count) ? cfg.ev_start_k + j_next * cfg.ev_step_k - r.k : NO_EVENT;
    }
    const bool ev_here = r.k >= cfg.ev_start_k && (r.k - cfg.ev_start_k) % cfg.ev_step_k == 0 &&
                         (r.k - cfg.ev_start_k) / cfg.ev_step_k < cfg.ev_count;
    int lvl_in = ((cfg.refine_on_action && act != 0) || ev_here) ? cfg.refine_input_level : 0;
    bool redo = false;   // the hot loop stopped at a sub-step that needs library transcendentals
    auto event_due = [&]() {
      if (ev_left != 0) return;
      const int j = (r.k - cfg.ev_start_k) / cfg.ev_step_k;
      apply_event(cfg, vtab, stab, ld, e, env_glob, (uint32_t)r.episode, j, r.Vgrid, r.Sinsol);
      in = make_inputs(cfg, r.Vgrid, r.Qref, r.Vdcref, r.Sinsol);
      ev_left = (j + 1 < cfg.ev_count) ? cfg.ev_step_k : NO_EVENT;
      lvl_in = cfg.refine_input_level;
    };
    do {
      // [A] one sub-step through the slow path, if this one needs it.  No `else`: lanes that diverge here (an action
      // refined in some envs of the warp only) rejoin the others for the hot run below instead of serialising it.
      {
        const int lvl_st = (r.k < cfg.startup_substeps) ? cfg.startup_level : cfg.base_level;
        const int lvl = lvl_in > lvl_st ? lvl_in : lvl_st;
        lvl_in = 0;
        if (lvl != 0 || redo) {
          redo = false;
          bool m_over;
          const unsigned frz = freeze_bits<M>(r.y, par, in, m_over);
          int flags = (frz != 0u ? 1 : 0) | (m_over ? 2 : 0);
          r.exact += ros_slow_substep<M>(r.y, par, in, tab, frz, base, lvl, flags);
          r.windup += flags & 1;
          if (M::BALANCED3 && (flags & 2)) r.status = PVDER_STATUS_UNBALANCED;
          if (RECORD) {
            if (traj) record_substep<M>(traj, traj_ld, r.k - k0, r.y, r.Vgrid, r.Sinsol);
          }
          r.k += 1;
          ev_left -= 1;
          event_due();
        }
      }
      // [B] the run of plain sub-steps that follows: up to the next event instant, the end of the start-up phase or
      // the end of the env step (empty when the next sub-step is refined again)
      int seg = k_end - r.k;
      if (ev_left < seg) seg = ev_left;
      if (r.k < cfg.startup_substeps) {
        if (cfg.startup_level != 0) seg = 0;
        else if (cfg.startup_substeps - r.k < seg) seg = cfg.startup_substeps - r.k;   // base_level starts there
      } else if (cfg.base_level != 0) seg = 0;
      if (lvl_in != 0) seg = 0;
      if (seg > 0) {
        // hot loop: a countdown and the clamped-sub-step counter are all the integers it carries
        int left = seg, wind = 0;
#if PVDER_CARRY_LIMIT_FLAGS
        // the limit flags of the clamp travel from one sub-step to the next: ros_core evaluates them for the state it ends
        // in (the inputs are constant within a run), the loop head only branches on them
        bool m_next, i_next;
        limit_flags<M>(r.y, par, in, m_next, i_next);
#endif
#if defined(PVDER_HOT_UNROLL) && defined(__CUDACC__)
        constexpr int kHotUnroll = PVDER_HOT_UNROLL;
#pragma unroll kHotUnroll
#endif
        do {
#if PVDER_CARRY_LIMIT_FLAGS
          const bool m_over = m_next;
          const unsigned frz = (m_next || i_next) ? freeze_rows<M>(r.y, par, in, m_next, i_next) : 0u;
          if (!ros_step<M>(r.y, par, in, tab, frz, base, &m_next, &i_next)) break;
#else
          bool m_over;
          const unsigned frz = freeze_bits<M>(r.y, par, in, m_over);
          if (!ros_step<M>(r.y, par, in, tab, frz, base)) break;
#endif
          wind += frz != 0u ? 1 : 0;
          // the duty-cycle clamp acts on Re/Im parts per phase and would break the symmetry the
          // balanced representation relies on: report instead of integrating something else
          if (M::BALANCED3 && m_over) r.status = PVDER_STATUS_UNBALANCED;
          if (RECORD) {
            if (traj) record_substep<M>(traj, traj_ld, r.k + (seg - left) - k0, r.y, r.Vgrid, r.Sinsol);
          }
        } while (--left != 0);
        redo = left != 0;
        r.k += seg - left;
        ev_left -= seg - left;
        r.windup += wind;
        event_due();
      }
    } while (r.k != k_end);
    bool finite = true;
#pragma unroll
    for (int i = 0; i < NS; ++i) finite &= (bool)isfinite(r.y[i]);
    if (!finite) r.status = PVDER_STATUS_NONFINITE;
  }

  compute_outputs<M>(cfg, r.y, r.Qref, r.Vdcref, r.Vgrid, r.Sinsol, r.k, o);
  done_out = r.done;
  if (run) {
    if (r.status == PVDER_STATUS_NONFINITE || r.status == PVDER_STATUS_UNBALANCED) {   // PVDER_env.py:170-172: -100, episode ends
      o.reward = -100.0;
      o.reward_i = -100;
      done_out = 1;
    }
    if (r.k >= cfg.done_substep) done_out = 1;  // PVDER_env.py:183
    r.last_reward = o.reward;
    r.ret += o.reward;                          // env_utilities.py:32-38
    r.done = done_out;
  } else {
    o.reward = r.last_reward;                   // cached tuple, PVDER_env.py:196
    o.reward_i = (int)r.last_reward;
  }
  if (run && done_out && cfg.auto_reset) {
    // vector-env convention: final reward/done are reported, obs is the first of the new episode
    const double rew = o.reward;
    const int rew_i = o.reward_i;
    init_env<M>(cfg, r.y, r.Qref, r.Vdcref, r.Vgrid, r.Sinsol);
    r.episode += 1;
    r.k = 0; r.steps = 0; r.done = 0; r.ret = 0.0; r.status = PVDER_STATUS_OK; r.windup = 0; r.exact = 0;
    if (cfg.ev_start_k == 0 && cfg.ev_count > 0)
      apply_event(cfg, vtab, stab, ld, e, env_glob, (uint32_t)r.episode, 0, r.Vgrid, r.Sinsol);
    compute_outputs<M>(cfg, r.y, r.Qref, r.Vdcref, r.Vgrid, r.Sinsol, r.k, o);
    o.reward = rew;
    o.reward_i = rew_i;
    hist_inc = -1;
    hist_clear = true;
  }
  return run;
}


}  // namespace pvder
