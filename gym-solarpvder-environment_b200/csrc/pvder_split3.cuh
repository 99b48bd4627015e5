// General three-phase PV-DER model (23 states) integrated by THREE LANES PER
// ENVIRONMENT: lane p of a group owns phase p (i, x, u: 6 states) and a replicated copy of the
// shared tail (Vdc, xDC, xQ, xPLL, delta); the three phases couple only through three sums
// (reactive power, inverter power, PLL d-axis voltage) that are formed with warp shuffles.
//
// Why: one thread cannot hold the 23-state working set of a Rosenbrock step (23 states, 174 LU entries, 4-5 stage
// vectors) in 255 registers -- the one-thread kernel moves 160 GB of spill traffic per 1 Mi-env launch
// (profiles/r1d_step_kernel_3ph_general_ncu_full.csv).  Split by phase, a lane carries the same 11
// values per vector as the single-phase kernel and no dense LU at all:
//
//   W K = b,  W = I/(h g) - J  is block-arrow: per phase a 6x6 block that reduces by hand to a 2x2
//   system in (KiR, KiI) [x and u rows are eliminated in closed form], bordered by the shared tail.
//   The phase blocks respond to the border only through beta = (d wr, K_Vdc, d irefR, d irefI); the
//   border closes on u = (dQ, d vd, K_Vdc), a 3x3 system whose matrix needs 13 three-lane sums per
//   factorisation and whose right-hand side needs 3 per stage.  Everything else is local FMAs.
//
// The same source is the CUDA kernel body (V = double, one lane per phase, shuffles) and -- built as
// plain C++ for tests/host_emul -- a three-wide value type (V = V3, sums over its elements).
//
//   advance_env_split <- PVDER.step  reference gym_PVDER/envs/PVDER_env.py:138-196 (same semantics as
//                        advance_env in pvder_env_step.cuh; the integrator state is lane-split)
#pragma once
#include "pvder_env_step.cuh"

namespace pvder {

// ---------------------------------------------------------------------------------------------
// Lane abstraction
// ---------------------------------------------------------------------------------------------
#ifndef PVDER_SPLIT_SMEM_SUMS
#define PVDER_SPLIT_SMEM_SUMS 1
#endif
#define PVDER_SPLIT_SUMS_MAX 13
#ifdef __CUDACC__
// DYN = false: the whole warp is converged (hot path; the member mask is the compile-time constant
// 0xffffffff, which lets ptxas drop the destination initialisation of every SHFL).  DYN = true: a
// divergent region (out-of-line slow path) entered by the lanes in `mask`.
template <bool DYN>
struct LanesT {
  using V = double;   // per-phase value: this lane's phase
  using B = bool;
  unsigned mask;      // lanes executing the current region (DYN only)
  double* sm;         // this warp's exchange buffer [PVDER_SPLIT_SUMS_MAX][32] in shared memory (sum3n)
  int self;           // this lane's own index (lanes 30, 31 shadow 27, 28 but keep their own buffer slot)
  int base;           // first lane of the group (phase a)
  int p;              // phase of this lane
  int n1, n2;         // lanes holding the next two phases (cyclic)

  // The member mask of every shuffle/vote: the compile-time full warp on the hot path (ptxas then issues each collective as
  // a plain SHFL/VOTE with a "branch if divergent" to an out-of-line WARPSYNC.COLLECTIVE stub), the ballot mask of the
  // region in the out-of-line path.  RULE for every user: a collective must be executed by ALL lanes of its mask -- never
  // inside a per-group branch, and never as the right operand of && / || (short-circuit evaluation is such a branch: the
  // two bugs that hung this kernel on B200, or killed it with "illegal instruction" inside __shfl_sync, were a retry that
  // only some groups of a shared mask took and `any3(a) || any3(b)`).
  PVDER_DEV unsigned m() const { return DYN ? mask : 0xffffffffu; }
  // Sum over the three phases, two shuffles: own + next + next-next.  Lane a gets (a + b) + c; the other
  // lanes get the same sum in a rotated order (last-bit differences), so every DECISION derived from a
  // sum is taken group-wide (any3 / from_a) -- see gains() and the range flag of the stepper core.
  PVDER_DEV double sum3(double v) const {
    const double b = __shfl_sync(m(), v, n1), c = __shfl_sync(m(), v, n2);
    return __dadd_rn(__dadd_rn(v, b), c);
  }
  // N phase sums at once (same order as sum3: own + next + next-next).  Through shared memory when PVDER_SPLIT_SMEM_SUMS:
  // N STS.64, a warp barrier, 2 N LDS.64 into aligned register pairs and a closing barrier instead of 4 N SHFL plus
  // about as many register moves (SHFL results do not land in aligned pairs under a 255-register allocation) -- the
  // three-lane kernel issues more non-FP64 instructions per FP64 instruction than the pipe takes for free.
  // Lanes only ever read the slots of their own group, and the lanes of a group always execute together, so the barriers
  // order exactly what they have to (and they name the same member mask as every other collective: rule at m()).
  template <int N>
  PVDER_DEV void sum3n(const double (&v)[N], double (&out)[N]) const {
#if PVDER_SPLIT_SMEM_SUMS
    static_assert(N <= PVDER_SPLIT_SUMS_MAX, "exchange buffer too small");
#pragma unroll
    for (int k = 0; k < N; ++k) sm[32 * k + self] = v[k];
    __syncwarp(m());
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const double b = sm[32 * k + n1], c = sm[32 * k + n2];
      out[k] = __dadd_rn(__dadd_rn(v[k], b), c);
    }
    __syncwarp(m());
#else
#pragma unroll
    for (int k = 0; k < N; ++k) out[k] = sum3(v[k]);
#endif
  }
  // fixed order (a + b) + c in all three lanes (outputs: bit-identical to the one-thread kernel)
  PVDER_DEV double sum3_ordered(double v) const {
    const double a = __shfl_sync(m(), v, base), b = __shfl_sync(m(), v, base + 1), c = __shfl_sync(m(), v, base + 2);
    return __dadd_rn(__dadd_rn(a, b), c);
  }
  PVDER_DEV bool any3(bool b) const { return ((__ballot_sync(m(), b) >> base) & 7u) != 0u; }
  PVDER_DEV bool any_warp(bool b) const { return __ballot_sync(m(), b) != 0u; }
  PVDER_DEV double from_a(double v) const { return __shfl_sync(m(), v, base); }
  PVDER_DEV int from_a(int v) const { return __shfl_sync(m(), v, base); }
  PVDER_DEV double pc(double a, double b, double c) const { return p == 0 ? a : (p == 1 ? b : c); }
  // the lanes of this region for which pred holds (pred must be uniform within a group)
  PVDER_DEV LanesT<true> sub(bool pred) const {
    LanesT<true> l;
    l.mask = __ballot_sync(m(), pred);
    l.sm = sm; l.self = self;
    l.base = base; l.p = p; l.n1 = n1; l.n2 = n2;
    return l;
  }
};
using Lanes3 = LanesT<false>;
PVDER_DEV Lanes3 make_lanes(int lane, double* warp_buf = nullptr) {
  // lanes 3g..3g+2 = phases a, b, c of group g; lanes 30, 31 shadow lanes 27, 28
  Lanes3 l;
  l.sm = warp_buf;
  l.self = lane;
  const int g = lane < 30 ? lane / 3 : 9;
  l.p = lane < 30 ? lane - 3 * g : lane - 30;
  l.base = 3 * g;
  l.mask = 0xffffffffu;
  l.n1 = l.base + (l.p + 1) % 3;
  l.n2 = l.base + (l.p + 2) % 3;
  return l;
}
PVDER_DEV double vfma(double a, double b, double c) { return fma(a, b, c); }
PVDER_DEV double vsel(bool c, double a, double b) { return c ? a : b; }
PVDER_DEV double vsel_group(bool c, double a, double b) { return c ? a : b; }   // c: one decision for the three lanes of a group
PVDER_DEV double vrcp(double a) { return pvder_rcp(a); }
PVDER_DEV bool vgt(double a, double b) { return a > b; }
PVDER_DEV bool vor(bool a, bool b) { return a || b; }
PVDER_DEV bool vand(bool a, bool b) { return a && b; }
PVDER_DEV bool vnonfinite(double a) { return !(bool)isfinite(a); }
PVDER_DEV double vmul_rn(double a, double b) { return __dmul_rn(a, b); }
PVDER_DEV double vadd_rn(double a, double b) { return __dadd_rn(a, b); }
#else
// Host emulation (tests/host_emul only): the three lanes of a group as one three-wide value.
struct V3 {
  double v[3];
  V3() : v{0.0, 0.0, 0.0} {}
  V3(double s) : v{s, s, s} {}
  V3(double a, double b, double c) : v{a, b, c} {}
};
struct B3 {
  bool v[3];
};
#define PVDER_V3_BIN(op)                                                        \
  inline V3 operator op(const V3& a, const V3& b) {                             \
    return V3(a.v[0] op b.v[0], a.v[1] op b.v[1], a.v[2] op b.v[2]);            \
  }
PVDER_V3_BIN(+)
PVDER_V3_BIN(-)
PVDER_V3_BIN(*)
PVDER_V3_BIN(/)
#undef PVDER_V3_BIN
inline V3 operator-(const V3& a) { return V3(-a.v[0], -a.v[1], -a.v[2]); }
inline V3 vfma(const V3& a, const V3& b, const V3& c) {
  return V3(std::fma(a.v[0], b.v[0], c.v[0]), std::fma(a.v[1], b.v[1], c.v[1]), std::fma(a.v[2], b.v[2], c.v[2]));
}
inline double vfma(double a, double b, double c) { return std::fma(a, b, c); }
inline V3 vrcp(const V3& a) { return V3(1.0 / a.v[0], 1.0 / a.v[1], 1.0 / a.v[2]); }
inline V3 vsel(const B3& c, const V3& a, const V3& b) {
  return V3(c.v[0] ? a.v[0] : b.v[0], c.v[1] ? a.v[1] : b.v[1], c.v[2] ? a.v[2] : b.v[2]);
}
inline V3 vsel_group(bool c, const V3& a, const V3& b) { return c ? a : b; }
inline B3 vgt(const V3& a, const V3& b) { return B3{{a.v[0] > b.v[0], a.v[1] > b.v[1], a.v[2] > b.v[2]}}; }
inline B3 vand(const B3& a, bool b) { return B3{{a.v[0] && b, a.v[1] && b, a.v[2] && b}}; }
inline B3 vor(const B3& a, const B3& b) { return B3{{a.v[0] || b.v[0], a.v[1] || b.v[1], a.v[2] || b.v[2]}}; }
inline B3 vnonfinite(const V3& a) { return B3{{!std::isfinite(a.v[0]), !std::isfinite(a.v[1]), !std::isfinite(a.v[2])}}; }
inline V3 vmul_rn(const V3& a, const V3& b) { return a * b; }
inline V3 vadd_rn(const V3& a, const V3& b) { return a + b; }
template <bool DYN>
struct LanesT {
  using V = V3;
  using B = B3;
  double sum3(const V3& v) const { return (v.v[0] + v.v[1]) + v.v[2]; }
  double sum3_ordered(const V3& v) const { return (v.v[0] + v.v[1]) + v.v[2]; }
  template <int N>
  void sum3n(const V3 (&v)[N], double (&out)[N]) const {
    for (int k = 0; k < N; ++k) out[k] = sum3(v[k]);
  }
  bool any3(const B3& b) const { return b.v[0] || b.v[1] || b.v[2]; }
  bool any3(bool b) const { return b; }
  bool any_warp(bool b) const { return b; }
  double from_a(const V3& v) const { return v.v[0]; }
  double from_a(double v) const { return v; }
  int from_a(int v) const { return v; }
  V3 pc(double a, double b, double c) const { return V3(a, b, c); }
  LanesT<true> sub(bool) const { return LanesT<true>(); }
};
using Lanes3 = LanesT<false>;
#endif

PVDER_DEV Lanes3::B vsame_sign(const Lanes3::V& a, const Lanes3::V& b) {
  // np.sign(a) == np.sign(b)  (pvder's clamping test)
  const Lanes3::V z(0.0);
  const Lanes3::B ap = vgt(a, z), an = vgt(z, a), bp = vgt(b, z), bn = vgt(z, b);
#ifdef __CUDACC__
  return (ap == bp) && (an == bn);
#else
  return B3{{(ap.v[0] == bp.v[0]) && (an.v[0] == bn.v[0]), (ap.v[1] == bp.v[1]) && (an.v[1] == bn.v[1]),
             (ap.v[2] == bp.v[2]) && (an.v[2] == bn.v[2])}};
#endif
}

// ---------------------------------------------------------------------------------------------
// Model
// ---------------------------------------------------------------------------------------------
struct Split3 {
  using V = Lanes3::V;
  using B = Lanes3::B;
  static constexpr int NS = 23;
  static constexpr int NS_STORE = 23;
  static constexpr int PHASES = 3;
  static constexpr int N_LUC = 13;
  // Unit-pivot slots (pure integrators: x of each phase; xDC, xQ, xPLL of the shared tail): their equations are
  // scaled by h*gamma like in the one-thread models (gains and c_ij/h sums pre-scaled, pivot exactly 1).
  static constexpr PVDER_HD bool unit_p(int i) { return i == 2 || i == 3; }
  static constexpr PVDER_HD bool unit_s(int i) { return i >= 1 && i <= 3; }

  // Launch constants: reciprocal pivots and every gain-derived coefficient of the factorisation for the
  // common case that no anti-windup clamp is active (FREE instantiations read them from the constant bank
  // instead of carrying them in registers).
  enum { LC_INV_GH = 0, LC_DU = 1, LC_GX = 2, LC_TH = 3, LC_KA = 4, LC_TK = 5, LC_G4H = 6, LC_G5H = 7, LC_PID_DC = 8,
         LC_PIQ = 9, LC_PIW = 10, LC_PID = 11, LC_KIPLL_H = 12 };
  static PVDER_HD void lu_consts(const Params& par, double ghinv, double* luc) {
    const double inv_gh = 1.0 / ghinv;
    luc[LC_INV_GH] = inv_gh;
    luc[LC_DU] = 1.0 / (ghinv + par.wp);           // reciprocal pivot of a free u row (a clamped one: 1/ghinv)
    luc[LC_GX] = par.Ki_GCC * inv_gh;
    luc[LC_TH] = luc[LC_GX] + par.Kp_GCC;
    luc[LC_KA] = luc[LC_DU] * par.wp;
    luc[LC_TK] = luc[LC_TH] * luc[LC_KA];
    luc[LC_G4H] = par.Ki_DC * inv_gh;
    luc[LC_G5H] = par.Ki_Q * inv_gh;
    luc[LC_PID_DC] = luc[LC_G4H] + par.Kp_DC;
    luc[LC_PIQ] = luc[LC_G5H] + par.Kp_Q;
    const double kpll = par.Ki_PLL * inv_gh + par.Kp_PLL;
    luc[LC_PIW] = par.inv_wb * kpll;
    luc[LC_PID] = kpll * inv_gh;
    luc[LC_KIPLL_H] = par.Ki_PLL * inv_gh;
  }

  // One lane's share of a 23-vector: p = (iR, iI, xR, xI, uR, uI) of its phase, s = (Vdc, xDC, xQ,
  // xPLL, delta) replicated in the three lanes.
  struct Vec {
    V p[6];
    double s[5];
  };
  struct Consts {     // phasor rotation of this phase and the abc->dq angle offset (cos, sin)
    V rr, ri, ca, sa;
  };
  struct In {
    V vgR, vgI;       // LV-side grid phasor of this phase (unbalance ratio and rotation applied)
    double Qref, Vdcref;
  };
  struct Gains {      // effective gains of the freezable rows (0 while clamped)
    V g0, g1, g2, g3; // xR, xI (Ki_GCC * h gamma: unit-pivot rows) ; uR, uI (wp)
    V duR, duI;       // 1/(ghinv + g2), 1/(ghinv + g3)
    double g4, g5;    // xDC (Ki_DC * h gamma), xQ (Ki_Q * h gamma)
    bool any;
  };
  struct Pt {         // algebraic quantities at one state (A.2)
    V vR, vI, mR, mI, ck, sk;
    double Qp, Ps, vd, wex, wr, dV, dQ, irefR, irefI;
  };

  template <class L>
  static PVDER_DEV Consts consts(const L& ln) {
    constexpr double C3 = 0.86602540378443864676;
    Consts k;
    k.rr = ln.pc(1.0, -0.5, -0.5);
    k.ri = ln.pc(0.0, -C3, C3);
    k.ca = ln.pc(1.0, -0.5, -0.5);
    k.sa = ln.pc(0.0, C3, -C3);
    return k;
  }

  template <class L>
  static PVDER_DEV In inputs(const L& ln, const Consts& k, const Inputs& in) {
    In o;
    const V vgk = ln.pc(in.vg, in.vgb, in.vgc);
    o.vgR = vmul_rn(vgk, k.rr);
    o.vgI = vmul_rn(vgk, k.ri);
    o.Qref = in.Qref;
    o.Vdcref = in.Vdcref;
    return o;
  }

  template <class L>
  static PVDER_DEV Pt point(const L& ln, const Params& par, const Consts& k, const In& in, const Aux& ax, const Vec& Y) {
    Pt q;
    const V iR = Y.p[0], iI = Y.p[1];
    q.vR = vfma(par.Rt, iR, vfma(-par.Xt, iI, in.vgR));
    q.vI = vfma(par.Xt, iR, vfma(par.Rt, iI, in.vgI));
    q.mR = vfma(par.Kp_GCC, Y.p[4], Y.p[2]);
    q.mI = vfma(par.Kp_GCC, Y.p[5], Y.p[3]);
    const V qs = vfma(q.vI, iR, -(q.vR * iI));
    const V ps = vfma(q.mR, iR, q.mI * iI);
    q.ck = vfma(ax.cs, k.ca, ax.sn * k.sa);        // cos(delta - alpha_p)
    q.sk = vfma(ax.sn, k.ca, -(ax.cs * k.sa));     // sin(delta - alpha_p)
    const V vdk = vfma(q.vR, q.ck, q.vI * q.sk);
    const V pin[3] = {qs, ps, vdk};
    double psum[3];
    ln.sum3n(pin, psum);
    q.Qp = 0.5 * psum[0];
    q.Ps = psum[1];
    q.vd = (1.0 / 3.0) * psum[2];                  // positive-sequence d-axis voltage (A.4)
    q.wex = fma(par.Kp_PLL, q.vd, Y.s[3]);
    q.wr = fma(q.wex, par.inv_wb, par.w0 * par.inv_wb);
    q.dV = in.Vdcref - Y.s[0];
    q.dQ = in.Qref - q.Qp;
    q.irefR = fma(par.Kp_DC, q.dV, Y.s[1]);
    q.irefI = fma(-par.Kp_Q, q.dQ, Y.s[2]);
    return q;
  }

  // Autonomous right-hand side (A.3) from the point record.
  // FREE: no clamp active in this warp -> the gains are the parameters (constant bank), g is not read.
  // ACC: F arrives pre-loaded (the stage's sum of c_ij/h K_j) and the right-hand side is accumulated onto it -- the addend
  //      rides on each row's last multiply, so no separate adds sit between the right-hand side and the solve.
  template <bool FREE, bool ACC = false>
  static PVDER_DEV void rhs(const Params& par, const Consts& k, const Aux& ax, const Gains& g, const double* luc,
                            const Vec& Y, const Pt& q, Vec& F) {
    // x, xDC, xQ, xPLL rows: gains pre-scaled by h*gamma (unit-pivot rows)
    const V g0 = FREE ? V(luc[LC_GX]) : g.g0, g1 = FREE ? V(luc[LC_GX]) : g.g1;
    const V g2 = FREE ? V(par.wp) : g.g2, g3 = FREE ? V(par.wp) : g.g3;
    const double g4 = FREE ? luc[LC_G4H] : g.g4, g5 = FREE ? luc[LC_G5H] : g.g5;
    const double hV = 0.5 * Y.s[0];
    const V iR = Y.p[0], iI = Y.p[1];
    const V eR = vfma(q.mR, hV, vfma(-par.Rf, iR, -q.vR)), eI = vfma(q.mI, hV, vfma(-par.Rf, iI, -q.vI));
    const V rfR = vfma(k.rr, q.irefR, -(k.ri * q.irefI)), rfI = vfma(k.ri, q.irefR, k.rr * q.irefI);
    const V dR = (rfR - Y.p[4]) - iR, dI = (rfI - Y.p[5]) - iI;
    if (ACC) {
      F.p[0] = vfma(q.wr, iI, vfma(V(par.inv_Lf), eR, F.p[0]));
      F.p[1] = vfma(-q.wr, iR, vfma(V(par.inv_Lf), eI, F.p[1]));
      F.p[2] = vfma(g0, Y.p[4], F.p[2]);
      F.p[3] = vfma(g1, Y.p[5], F.p[3]);
      F.p[4] = vfma(g2, dR, F.p[4]);
      F.p[5] = vfma(g3, dI, F.p[5]);
      F.s[0] = fma(par.inv_C, fma(-0.25, q.Ps, ax.PoV), F.s[0]);
      F.s[1] = fma(g4, q.dV, F.s[1]);
      F.s[2] = fma(-g5, q.dQ, F.s[2]);
      F.s[3] = fma(luc[LC_KIPLL_H], q.vd, F.s[3]);
      F.s[4] = (q.wex + par.dw) + F.s[4];
    } else {
      F.p[0] = vfma(q.wr, iI, par.inv_Lf * eR);
      F.p[1] = vfma(-q.wr, iR, par.inv_Lf * eI);
      F.p[2] = g0 * Y.p[4];
      F.p[3] = g1 * Y.p[5];
      F.p[4] = g2 * dR;
      F.p[5] = g3 * dI;
      F.s[0] = par.inv_C * fma(-0.25, q.Ps, ax.PoV);      // (Ppv - Vdc Ps / 4) / (C Vdc) = (Ppv / Vdc - Ps / 4) / C
      F.s[1] = g4 * q.dV;
      F.s[2] = -(g5 * q.dQ);
      F.s[3] = luc[LC_KIPLL_H] * q.vd;
      F.s[4] = q.wex + par.dw;
    }
  }

  // Factors of W = I/(h g) - J(y) in block-arrow form (see the file header).
  struct Fac {
    V n11, n22, n12;     // inverse of the 2x2 current block [[A_R, -c], [c, A_I]] = [[n11, n12], [-n12, n22]]
    V thR, thI;          // d(m)/d(Ku): g0/ghinv + Kp_GCC
    V kaR, kaI;          // du * g2: feedback of (iref - Ki) into Ku
    V epR, epI;          // e * th * ka
    V hmR, hmI;          // 0.5 inv_Lf m: response of the current rows to K_Vdc
    V qR, qI, dR, dI, pR, pI;   // rows of dQ, d vd, d Ps with respect to (KiR, KiI)
    V gxR, gxI;          // g0/ghinv, g1/ghinv
    double e;            // inv_Lf Vdc / 2
    double RQ[3], Rv[3], RP[3];   // response of the three sums to beta_1 (d wr), beta_3, beta_4 (d iref)
    double N[9];         // inverse of the 3x3 border matrix, row-major
    double vdd;          // d vd / d delta
    double piD, piQ, g4h, g5h;
    // thR..kaI, gxR, gxI, piD, piQ, g4h, g5h depend on the gains only: FREE instantiations neither write nor read them
  };

  template <bool FREE, class L>
  static PVDER_DEV void factor(const L& ln, const Params& par, const Consts& k, const In& in, const Aux& ax,
                               const Gains& g, const Vec& y, const Pt& q, double ghinv, const double* luc, Fac& f) {
    const double inv_gh = luc[0];
    const V iR = y.p[0], iI = y.p[1];
    f.e = par.inv_Lf * (0.5 * y.s[0]);
    const double a = fma(par.inv_Lf, par.Rf + par.Rt, ghinv);
    const double c = fma(par.inv_Lf, par.Xt, q.wr);
    V tkR, tkI;
    if (FREE) {
      tkR = V(luc[LC_TK]);
      tkI = tkR;
    } else {
      f.gxR = g.g0;        // gains of the unit-pivot rows arrive scaled by h*gamma
      f.gxI = g.g1;
      f.thR = f.gxR + par.Kp_GCC;
      f.thI = f.gxI + par.Kp_GCC;
      f.kaR = g.duR * g.g2;
      f.kaI = g.duI * g.g3;
      tkR = f.thR * f.kaR;
      tkI = f.thI * f.kaI;
    }
    f.epR = f.e * tkR;
    f.epI = f.e * tkI;
    const V AR = f.epR + a, AI = f.epI + a;
    const V det = vfma(AR, AI, V(c * c));
    const V idet = vrcp(det);
    f.n11 = AI * idet;
    f.n22 = AR * idet;
    f.n12 = c * idet;
    f.hmR = (0.5 * par.inv_Lf) * q.mR;
    f.hmI = (0.5 * par.inv_Lf) * q.mI;
    f.qR = vfma(par.Xt, iR, 0.5 * in.vgI);       // d Qp / d iR = 0.5 (vgI + 2 Xt iR)
    f.qI = vfma(par.Xt, iI, -0.5 * in.vgR);
    f.dR = (1.0 / 3.0) * vfma(par.Rt, q.ck, par.Xt * q.sk);
    f.dI = (1.0 / 3.0) * vfma(par.Rt, q.sk, -(par.Xt * q.ck));
    f.pR = vfma(-iR, tkR, q.mR);
    f.pI = vfma(-iI, tkI, q.mI);
    // response of (KiR, KiI) to the four border columns, then the three sums of each
    const V c1R = vfma(f.n11, iI, -(f.n12 * iR)), c1I = vfma(-f.n12, iI, -(f.n22 * iR));          // d wr : (iI, -iR)
    const V c2R = vfma(f.n11, f.hmR, f.n12 * f.hmI), c2I = vfma(-f.n12, f.hmR, f.n22 * f.hmI);    // K_Vdc : (hmR, hmI)
    const V e3R = f.epR * k.rr, e3I = f.epI * k.ri;                                               // d irefR
    const V c3R = vfma(f.n11, e3R, f.n12 * e3I), c3I = vfma(-f.n12, e3R, f.n22 * e3I);
    const V e4R = -(f.epR * k.ri), e4I = f.epI * k.rr;                                            // d irefI
    const V c4R = vfma(f.n11, e4R, f.n12 * e4I), c4I = vfma(-f.n12, e4R, f.n22 * e4I);
    const V uR = iR * tkR, uI = iI * tkI;    // direct part of d Ps / d iref through Ku
    const V fin[13] = {vfma(f.qR, c1R, f.qI * c1I), vfma(f.qR, c2R, f.qI * c2I), vfma(f.qR, c3R, f.qI * c3I),
                       vfma(f.qR, c4R, f.qI * c4I), vfma(f.dR, c1R, f.dI * c1I), vfma(f.dR, c2R, f.dI * c2I),
                       vfma(f.dR, c3R, f.dI * c3I), vfma(f.dR, c4R, f.dI * c4I), vfma(f.pR, c1R, f.pI * c1I),
                       vfma(f.pR, c2R, f.pI * c2I), vfma(f.pR, c3R, vfma(f.pI, c3I, vfma(uR, k.rr, uI * k.ri))),
                       vfma(f.pR, c4R, vfma(f.pI, c4I, vfma(uI, k.rr, -(uR * k.ri)))), vfma(q.vI, q.ck, -(q.vR * q.sk))};
    double fs[13];
    ln.sum3n(fin, fs);
    const double RQ1 = fs[0], RQ2 = fs[1], RQ3 = fs[2], RQ4 = fs[3];
    const double Rv1 = fs[4], Rv2 = fs[5], Rv3 = fs[6], Rv4 = fs[7];
    const double RP1 = fs[8], RP2 = fs[9], RP3 = fs[10], RP4 = fs[11];
    f.vdd = (1.0 / 3.0) * fs[12];
    f.RQ[0] = RQ1; f.RQ[1] = RQ3; f.RQ[2] = RQ4;
    f.Rv[0] = Rv1; f.Rv[1] = Rv3; f.Rv[2] = Rv4;
    f.RP[0] = RP1; f.RP[1] = RP3; f.RP[2] = RP4;
    const double piw = luc[LC_PIW], pid = luc[LC_PID];
    double piD, piQ;
    if (FREE) {
      piD = luc[LC_PID_DC];
      piQ = luc[LC_PIQ];
    } else {
      f.g4h = g.g4;
      f.g5h = g.g5;
      f.piD = piD = f.g4h + par.Kp_DC;
      f.piQ = piQ = f.g5h + par.Kp_Q;
    }
    const double qC = 0.25 * par.inv_C;
    const double jVV = par.inv_C * ax.dPoV;
    // border matrix in u = (dQ, d vd, K_Vdc)
    const double m00 = fma(-RQ4, piQ, 1.0), m01 = -(RQ1 * piw), m02 = fma(piD, RQ3, -RQ2);
    const double m10 = -(Rv4 * piQ), m11 = fma(-f.vdd, pid, fma(-Rv1, piw, 1.0)), m12 = fma(piD, Rv3, -Rv2);
    const double m20 = qC * (RP4 * piQ), m21 = qC * (RP1 * piw), m22 = fma(qC, fma(-piD, RP3, RP2), ghinv - jVV);
    const double k00 = fma(m11, m22, -(m12 * m21)), k01 = fma(m02, m21, -(m01 * m22)), k02 = fma(m01, m12, -(m02 * m11));
    const double k10 = fma(m12, m20, -(m10 * m22)), k11 = fma(m00, m22, -(m02 * m20)), k12 = fma(m02, m10, -(m00 * m12));
    const double k20 = fma(m10, m21, -(m11 * m20)), k21 = fma(m01, m20, -(m00 * m21)), k22 = fma(m00, m11, -(m01 * m10));
    const double dt = fma(m00, k00, fma(m01, k10, m02 * k20));
    const double idt = pvder_rcp(dt);
    f.N[0] = k00 * idt; f.N[1] = k01 * idt; f.N[2] = k02 * idt;
    f.N[3] = k10 * idt; f.N[4] = k11 * idt; f.N[5] = k12 * idt;
    f.N[6] = k20 * idt; f.N[7] = k21 * idt; f.N[8] = k22 * idt;
  }

  // b <- W^-1 b
  template <bool FREE, class L>
  static PVDER_DEV void solve(const L& ln, const Params& par, const Consts& k, const Gains& g, const Fac& f,
                              const Vec& y, const double* luc, Vec& b) {
    const double inv_gh = luc[LC_INV_GH];
    const V duR = FREE ? V(luc[LC_DU]) : g.duR, duI = FREE ? V(luc[LC_DU]) : g.duI;
    const V thR = FREE ? V(luc[LC_TH]) : f.thR, thI = FREE ? V(luc[LC_TH]) : f.thI;
    const V kaR = FREE ? V(luc[LC_KA]) : f.kaR, kaI = FREE ? V(luc[LC_KA]) : f.kaI;
    const V gxR = FREE ? V(luc[LC_GX]) : f.gxR, gxI = FREE ? V(luc[LC_GX]) : f.gxI;
    const double piD = FREE ? luc[LC_PID_DC] : f.piD, piQ = FREE ? luc[LC_PIQ] : f.piQ;
    const double g4h = FREE ? luc[LC_G4H] : f.g4h, g5h = FREE ? luc[LC_G5H] : f.g5h;
    const double piw = luc[LC_PIW], pid = luc[LC_PID];
    const V iR = y.p[0], iI = y.p[1];
    const V hxR = b.p[2], hxI = b.p[3];             // unit-pivot rows: right-hand side arrives scaled by h*gamma
    const V buR = duR * b.p[4], buI = duI * b.p[5];
    const V mmR = vfma(thR, buR, hxR), mmI = vfma(thI, buI, hxI);
    const V rR = vfma(f.e, mmR, b.p[0]), rI = vfma(f.e, mmI, b.p[1]);
    const V tR = vfma(f.n11, rR, f.n12 * rI), tI = vfma(f.n22, rI, -(f.n12 * rR));
    const V sin3[3] = {vfma(f.qR, tR, f.qI * tI), vfma(f.dR, tR, f.dI * tI), vfma(f.pR, tR, vfma(f.pI, tI, vfma(iR, mmR, iI * mmI)))};
    double ss[3];
    ln.sum3n(sin3, ss);
    const double sQ = ss[0], sv = ss[1], sP = ss[2];
    const double b3 = b.s[1], b4 = b.s[2], hp = b.s[3];
    const double b1 = par.inv_wb * hp;
    const double kd0 = (b.s[4] + hp) * inv_gh;
    const double r1 = fma(f.RQ[0], b1, fma(f.RQ[1], b3, fma(f.RQ[2], b4, sQ)));
    const double r2 = fma(f.vdd, kd0, fma(f.Rv[0], b1, fma(f.Rv[1], b3, fma(f.Rv[2], b4, sv))));
    const double cP = fma(f.RP[0], b1, fma(f.RP[1], b3, fma(f.RP[2], b4, sP)));
    const double r3 = fma(-0.25 * par.inv_C, cP, b.s[0]);
    const double dQ = fma(f.N[0], r1, fma(f.N[1], r2, f.N[2] * r3));
    const double dvd = fma(f.N[3], r1, fma(f.N[4], r2, f.N[5] * r3));
    const double KV = fma(f.N[6], r1, fma(f.N[7], r2, f.N[8] * r3));
    const double be1 = fma(piw, dvd, b1);
    const double be3 = fma(-piD, KV, b3);
    const double be4 = fma(piQ, dQ, b4);
    const V rfR = vfma(k.rr, be3, -(k.ri * be4)), rfI = vfma(k.ri, be3, k.rr * be4);
    const V r2R = vfma(iI, be1, vfma(f.hmR, KV, vfma(f.epR, rfR, rR)));
    const V r2I = vfma(-iR, be1, vfma(f.hmI, KV, vfma(f.epI, rfI, rI)));
    const V KiR = vfma(f.n11, r2R, f.n12 * r2I), KiI = vfma(f.n22, r2I, -(f.n12 * r2R));
    const V KuR = vfma(kaR, rfR - KiR, buR), KuI = vfma(kaI, rfI - KiI, buI);
    b.p[0] = KiR;
    b.p[1] = KiI;
    b.p[2] = vfma(gxR, KuR, hxR);
    b.p[3] = vfma(gxI, KuI, hxI);
    b.p[4] = KuR;
    b.p[5] = KuI;
    b.s[0] = KV;
    b.s[1] = fma(-g4h, KV, b3);
    b.s[2] = fma(g5h, dQ, b4);
    b.s[3] = fma(luc[LC_KIPLL_H], dvd, hp);
    b.s[4] = fma(pid, dvd, kd0);
  }

  // Anti-windup mode (A.3) sampled at the sub-step start; same decisions as freeze_bits<Model3ph>.
  template <class L>
  static PVDER_DEV Gains gains(const L& ln, const Params& par, const Consts& k, const In& in, const Vec& y,
                               const double* luc, bool& m_over_out) {
    Gains g;
    const V iR = y.p[0], iI = y.p[1];
    const V mR = vfma(par.Kp_GCC, y.p[4], y.p[2]), mI = vfma(par.Kp_GCC, y.p[5], y.p[3]);
    const bool m_over = ln.any3(vgt(mR * mR + mI * mI, V(par.m_limit10 * par.m_limit10)));
    const double Q = ln.sum3(0.5 * (in.vgI * iR - in.vgR * iI + par.Xt * (iR * iR + iI * iI)));
    const double Vdc = y.s[0], xDC = y.s[1], xQ = y.s[2];
    const double irefR = xDC + par.Kp_DC * (in.Vdcref - Vdc);
    const double irefI = xQ - par.Kp_Q * (in.Qref - Q);
    // decisions on the shared rows come from sums: take lane a's so the three lanes agree
    const int sbits = ln.from_a(((irefR * irefR + irefI * irefI) > par.iref_limit * par.iref_limit ? 1 : 0) |
                                (same_sign(par.Ki_DC * (in.Vdcref - Vdc), xDC) ? 2 : 0) |
                                (same_sign(-par.Ki_Q * (in.Qref - Q), xQ) ? 4 : 0));
    const bool i_over = (sbits & 1) != 0;
    m_over_out = m_over;
    // branch-free (the warp stays converged for the group votes): flags are masked by the group-wide
    // over-limit conditions instead of being computed under them
    const V zero(0.0), kig(luc[LC_GX]), wp(par.wp), du_free(luc[LC_DU]), du_frz(luc[LC_INV_GH]);
    const V uR = y.p[4], uI = y.p[5];
    const V duR = par.wp * (-uR + (k.rr * irefR - k.ri * irefI) - iR);
    const V duI = par.wp * (-uI + (k.ri * irefR + k.rr * irefI) - iI);
    const B f0 = vand(vsame_sign(par.Ki_GCC * uR, y.p[2]), m_over), f1 = vand(vsame_sign(par.Ki_GCC * uI, y.p[3]), m_over);
    const B f2 = vand(vsame_sign(duR, uR), m_over), f3 = vand(vsame_sign(duI, uI), m_over);
    g.g0 = vsel(f0, zero, kig); g.g1 = vsel(f1, zero, kig);
    g.g2 = vsel(f2, zero, wp); g.g3 = vsel(f3, zero, wp);
    g.duR = vsel(f2, du_frz, du_free); g.duI = vsel(f3, du_frz, du_free);
    const bool fdc = i_over && (sbits & 2), fq = i_over && (sbits & 4);
    g.g4 = fdc ? 0.0 : luc[LC_G4H];
    g.g5 = fq ? 0.0 : luc[LC_G5H];
    g.any = ln.any3(vor(vor(f0, f1), vor(f2, f3))) || fdc || fq;
    return g;
  }
};

// ---------------------------------------------------------------------------------------------
// Half-cycle step on the lane-split state (same scheme -- PVDER_SCHEME --, coefficients and incremental
// side-inputs as ros_core).  Stage vectors are folded into the sums that need them as soon as they exist, so
// at most five lane-vectors are live.
// ---------------------------------------------------------------------------------------------
template <bool EXACT, bool FREE, class LN, class TAB>
PVDER_DEV bool ros_core_split(const LN& ln, Split3::Vec& y, const pvder_env_config& cfg, const Inputs& in_s,
                                 const Split3::In& in, const Split3::Consts& k, const TAB& tab,
                                 const Split3::Gains& g, Aux& base, bool discard = false) {
  // discard: this group only keeps its warp converged through the hot step (its sub-step is refined out of line):
  // treated like a stage out of range -- y and base stay untouched, false is returned
  using S = Split3;
  using Vec = S::Vec;
  const Params& par = cfg.par;
  bool oor = discard;
  const double dl0 = y.s[4], V0 = y.s[0];
  pov_from_exp(par, in_s, base.E, base.PoV, base.dPoV);      // inputs (insolation) may have changed
  Vec K1, K2, K3, K4, Y;
  Aux ax;
  S::Fac fac;
  {
    const S::Pt q = S::point(ln, par, k, in, base, y);
    S::template factor<FREE>(ln, par, k, in, base, g, y, q, tab.ghinv, tab.luc, fac);
    S::template rhs<FREE>(par, k, base, g, tab.luc, y, q, K1);
  }
  S::template solve<FREE>(ln, par, k, g, fac, y, tab.luc, K1);
  // the same statement on the six per-phase slots (V) and the five shared slots (double); CC(m, nn) = c_nn/h,
  // or c_nn*gamma on a unit-pivot slot
#define CC(m, nn) (S::unit_##m(i) ? tab.cs##nn : tab.c##nn)
#define PVDER_EACH(ST)                                        \
  _Pragma("unroll") for (int i = 0; i < 6; ++i) { ST(p) }     \
  _Pragma("unroll") for (int i = 0; i < 5; ++i) { ST(s) }
#if PVDER_SCHEME == 4
  // ROS4-L (see ros_core): 3 right-hand sides, 4 solves; stage 4 re-uses stage 3's right-hand side
#define DD(m, nn) (S::unit_##m(i) ? tab.ds##nn : tab.d##nn)
  // stage 2
#define ST(m) Y.m[i] = vfma(tab.a21, K1.m[i], y.m[i]);
  PVDER_EACH(ST)
#undef ST
  aux_advance_sv<EXACT, false, true>(par, in_s, base, dl0, V0, Y.s[4], Y.s[0], ax, oor);
  // the stage's sum of c_ij/h K_j is pre-loaded and the right-hand side accumulated onto it (rhs<.., true>)
#define ST(m) K2.m[i] = CC(m, 21) * K1.m[i];
  PVDER_EACH(ST)
#undef ST
  S::template rhs<FREE, true>(par, k, ax, g, tab.luc, Y, S::point(ln, par, k, in, ax, Y), K2);
  S::template solve<FREE>(ln, par, k, g, fac, y, tab.luc, K2);
  // stage 3: K1, K2 are folded into Y3, the pre-loaded sums of stages 3 (-> K3) and 4 (-> K4) and the new state
  // (-> K1) as soon as K2 exists
#define ST(m)                                                                \
  {                                                                          \
    const auto k1 = K1.m[i], k2 = K2.m[i];                                   \
    Y.m[i] = vfma(tab.a32, k2, vfma(tab.a31, k1, y.m[i]));                   \
    K3.m[i] = vfma(CC(m, 32), k2, CC(m, 31) * k1);                           \
    K4.m[i] = vfma(DD(m, 42), k2, DD(m, 41) * k1);                           \
    K1.m[i] = vfma(tab.m2, k2, vfma(tab.m1, k1, y.m[i]));                    \
  }
  PVDER_EACH(ST)
#undef ST
  aux_advance_sv<EXACT, false, true>(par, in_s, base, dl0, V0, Y.s[4], Y.s[0], ax, oor);
  S::template rhs<FREE, true>(par, k, ax, g, tab.luc, Y, S::point(ln, par, k, in, ax, Y), K3);   // K3 = b_3
#if PVDER_STAGE4_DELTA
  // the fourth solve yields K_4 - K_3 (see ros_core): b_3 is not re-added
  S::template solve<FREE>(ln, par, k, g, fac, y, tab.luc, K3);
#define ST(m)                                   \
  K4.m[i] = vfma(CC(m, 43), K3.m[i], K4.m[i]);  \
  K1.m[i] = vfma(tab.m34, K3.m[i], K1.m[i]);
  PVDER_EACH(ST)
#undef ST
#else
#define ST(m) K4.m[i] = K4.m[i] + K3.m[i];
  PVDER_EACH(ST)
#undef ST
  S::template solve<FREE>(ln, par, k, g, fac, y, tab.luc, K3);
#define ST(m)                                   \
  K4.m[i] = vfma(CC(m, 43), K3.m[i], K4.m[i]);  \
  K1.m[i] = vfma(tab.m3, K3.m[i], K1.m[i]);
  PVDER_EACH(ST)
#undef ST
#endif
  S::template solve<FREE>(ln, par, k, g, fac, y, tab.luc, K4);
#define ST(m) Y.m[i] = vfma(tab.m4, K4.m[i], K1.m[i]);
  PVDER_EACH(ST)
#undef ST
#undef DD
#else   // Rodas4
  // stage 2
#define ST(m) Y.m[i] = vfma(tab.a21, K1.m[i], y.m[i]);
  PVDER_EACH(ST)
#undef ST
  aux_advance_sv<EXACT, false, true>(par, in_s, base, dl0, V0, Y.s[4], Y.s[0], ax, oor);
  S::template rhs<FREE>(par, k, ax, g, tab.luc, Y, S::point(ln, par, k, in, ax, Y), K2);
#define ST(m) K2.m[i] = vfma(CC(m, 21), K1.m[i], K2.m[i]);
  PVDER_EACH(ST)
#undef ST
  S::template solve<FREE>(ln, par, k, g, fac, y, tab.luc, K2);
  // stage 3
#define ST(m) Y.m[i] = vfma(tab.a32, K2.m[i], vfma(tab.a31, K1.m[i], y.m[i]));
  PVDER_EACH(ST)
#undef ST
  aux_advance_sv<EXACT, false, true>(par, in_s, base, dl0, V0, Y.s[4], Y.s[0], ax, oor);
  S::template rhs<FREE>(par, k, ax, g, tab.luc, Y, S::point(ln, par, k, in, ax, Y), K3);
#define ST(m) K3.m[i] = vfma(CC(m, 32), K2.m[i], vfma(CC(m, 31), K1.m[i], K3.m[i]));
  PVDER_EACH(ST)
#undef ST
  S::template solve<FREE>(ln, par, k, g, fac, y, tab.luc, K3);
  // stage 4
#define ST(m) Y.m[i] = vfma(tab.a43, K3.m[i], vfma(tab.a42, K2.m[i], vfma(tab.a41, K1.m[i], y.m[i])));
  PVDER_EACH(ST)
#undef ST
  aux_advance_sv<EXACT, false, true>(par, in_s, base, dl0, V0, Y.s[4], Y.s[0], ax, oor);
  S::template rhs<FREE>(par, k, ax, g, tab.luc, Y, S::point(ln, par, k, in, ax, Y), K4);
#define ST(m) K4.m[i] = vfma(CC(m, 43), K3.m[i], vfma(CC(m, 42), K2.m[i], vfma(CC(m, 41), K1.m[i], K4.m[i])));
  PVDER_EACH(ST)
#undef ST
  S::template solve<FREE>(ln, par, k, g, fac, y, tab.luc, K4);
  // fold K1..K4 into Y5, C5 (-> K2) and C6 (-> K3); K1, K4 are dead afterwards
#define ST(m)                                                                                                  \
  {                                                                                                            \
    const auto k1 = K1.m[i], k2 = K2.m[i], k3 = K3.m[i], k4 = K4.m[i];                                         \
    Y.m[i] = vfma(tab.a54, k4, vfma(tab.a53, k3, vfma(tab.a52, k2, vfma(tab.a51, k1, y.m[i]))));               \
    K2.m[i] = vfma(CC(m, 54), k4, vfma(CC(m, 53), k3, vfma(CC(m, 52), k2, CC(m, 51) * k1)));                           \
    K3.m[i] = vfma(CC(m, 64), k4, vfma(CC(m, 63), k3, vfma(CC(m, 62), k2, CC(m, 61) * k1)));                           \
  }
  PVDER_EACH(ST)
#undef ST
  // stage 5
  aux_advance_sv<EXACT, false, true>(par, in_s, base, dl0, V0, Y.s[4], Y.s[0], ax, oor);
  S::template rhs<FREE>(par, k, ax, g, tab.luc, Y, S::point(ln, par, k, in, ax, Y), K4);
#define ST(m) K2.m[i] = K2.m[i] + K4.m[i];
  PVDER_EACH(ST)
#undef ST
  S::template solve<FREE>(ln, par, k, g, fac, y, tab.luc, K2);      // K2 = K5
  // stage 6 (Y6 = Y5 + K5; y+ = Y6 + K6: stiffly accurate)
#define ST(m)                                   \
  Y.m[i] = Y.m[i] + K2.m[i];                    \
  K3.m[i] = vfma(CC(m, 65), K2.m[i], K3.m[i]);
  PVDER_EACH(ST)
#undef ST
  aux_advance_sv<EXACT, false, true>(par, in_s, base, dl0, V0, Y.s[4], Y.s[0], ax, oor);
  S::template rhs<FREE>(par, k, ax, g, tab.luc, Y, S::point(ln, par, k, in, ax, Y), K4);
#define ST(m) K3.m[i] = K3.m[i] + K4.m[i];
  PVDER_EACH(ST)
#undef ST
  S::template solve<FREE>(ln, par, k, g, fac, y, tab.luc, K3);      // K3 = K6
#define ST(m) Y.m[i] = Y.m[i] + K3.m[i];
  PVDER_EACH(ST)
#undef ST
#endif
#undef PVDER_EACH
#undef CC
  aux_advance_sv<EXACT, true, true>(par, in_s, base, dl0, V0, Y.s[4], Y.s[0], ax, oor);
  // group-wide decision (the three lanes must agree on the redo); a group that only keeps the warp company commits nothing
  if (!EXACT && ln.any3(oor)) return false;
  if (EXACT && discard) return false;
  y = Y;
  base = ax;
  return true;
}

// Out-of-line slow path (see ros_slow in pvder_env_step.cuh): level 0 = the half-cycle step redone with library
// transcendentals at every stage; level > 0 = the sub-step as 2^level steps of h / 2^level with tab->fine[level - 1], the
// clamp mode re-sampled before every fine step.  Every fine step uses library transcendentals too: the function stays
// straight-line code inside one loop whose trip count is the same for all lanes that enter it.
//
// CONVERGENCE CONTRACT (learned the hard way on B200: a version that let groups with different needs share one call --
// different loop lengths, a retry only for some -- hung the kernel or died with "illegal instruction" inside
// __shfl_sync as soon as one env of a warp left the beaten path).  The caller enters with the ballot mask of the groups
// that need exactly THIS level, so all lanes of `ln.mask` execute the same instructions in the same order; every shuffle
// and vote in here names exactly those lanes.  Everything travels by value so that nothing in the caller's hot loop has
// its address taken (that would pin it to local memory).
struct SplitStepResult {
  Split3::Vec y;
  Aux base;
  int clamped;
};
PVDER_NOINLINE SplitStepResult ros_slow_split(LanesT<true> ln, Split3::Vec y, const pvder_env_config* cfg, Inputs in_s,
                                              Split3::In in, Split3::Consts k, const RodasTab* tab,
                                              Split3::Gains g, Aux base, int level) {
  SplitStepResult r;
  r.clamped = 0;
  if (level <= 0) {      // (level is the same in every lane of the mask)
    ros_core_split<true, false>(ln, y, *cfg, in_s, in, k, static_cast<const RodasCoef&>(*tab), g, base);
  } else {
    const RodasCoef& ft = tab->fine[level - 1];
    const int nf = 1 << level;
#pragma unroll 1
    for (int j = 0; j < nf; ++j) {
      // gains carry h-scaled entries: rebuilt with this level's constants (also for the first fine step)
      bool m_over;
      g = Split3::gains(ln, cfg->par, k, in, y, ft.luc, m_over);
      r.clamped |= g.any ? 1 : 0;
      ros_core_split<true, false>(ln, y, *cfg, in_s, in, k, ft, g, base);
    }
  }
  r.y = y;
  r.base = base;
  return r;
}

// Registers of one environment as seen by one of its three lanes.
struct EnvRegsSplit {
  Split3::Vec y;
  double Qref, Vdcref, Vgrid, Sinsol, ret, last_reward;
  int k, steps, episode, status, done, windup, exact;
};

// Outputs from the lane-split state: the same individually rounded operations as
// compute_outputs_p<3>, with the three phase terms summed in the same order.
template <class LN>
PVDER_DEV void compute_outputs_split(const LN& ln, const pvder_env_config& cfg, const Split3::Consts& kc,
                                     const Split3::Vec& y, double Qref, double Vdcref, double Vgrid, double Sinsol,
                                     int k, Outputs& o) {
  using V = Split3::V;
  const Params& par = cfg.par;
  const Inputs in = make_inputs(cfg, Vgrid, Qref, Vdcref, Sinsol);
  const Split3::In inp = Split3::inputs(ln, kc, in);
  const V jR = y.p[0], jI = y.p[1];
  const V vkR = vadd_rn(inp.vgR, vadd_rn(vmul_rn(par.Rt, jR), -vmul_rn(par.Xt, jI)));
  const V vkI = vadd_rn(inp.vgI, vadd_rn(vmul_rn(par.Xt, jR), vmul_rn(par.Rt, jI)));
  const double Ppcc = ln.sum3_ordered(vmul_rn(0.5, vadd_rn(vmul_rn(vkR, jR), vmul_rn(vkI, jI))));
  const double Qpcc = ln.sum3_ordered(vmul_rn(0.5, vadd_rn(vmul_rn(vkI, jR), -vmul_rn(vkR, jI))));
  const double v2 = ln.sum3_ordered(vadd_rn(vmul_rn(vkR, vkR), vmul_rn(vkI, vkI)));
  finish_outputs<3>(cfg, in, ln.from_a(jR), ln.from_a(jI), ln.from_a(vkR), ln.from_a(vkI), Ppcc, Qpcc, v2, y.s[0], Qref,
                    Vdcref, k, o);
}

template <class LN>
PVDER_DEV void init_env_split(const LN& ln, const pvder_env_config& cfg, Split3::Vec& y, double& Qref,
                              double& Vdcref, double& Vgrid, double& Sinsol) {
#pragma unroll
  for (int i = 0; i < 6; ++i) y.p[i] = ln.pc(cfg.y0[i], cfg.y0[6 + i], cfg.y0[12 + i]);
#pragma unroll
  for (int i = 0; i < 5; ++i) y.s[i] = cfg.y0[18 + i];
  Qref = cfg.Q_ref0;
  Vdcref = cfg.Vdc_ref0;
  Vgrid = 1.0;
  Sinsol = 100.0;
}

// Trajectory recording (see record_substep): every lane writes the rows of its phase, lane a the shared
// tail and the event values.  traj is null in the lanes that do not record.
#ifdef __CUDACC__
template <class LN>
PVDER_DEV void record_substep_split(const LN& ln, double* traj, int64_t traj_ld, int s, const Split3::Vec& y, double Vgrid,
                                    double Sinsol) {
  double* row = traj + (int64_t)s * 25 * traj_ld;
#pragma unroll
  for (int i = 0; i < 6; ++i) row[(int64_t)(6 * ln.p + i) * traj_ld] = y.p[i];
  if (ln.p == 0) {
#pragma unroll
    for (int i = 0; i < 5; ++i) row[(int64_t)(18 + i) * traj_ld] = y.s[i];
    row[(int64_t)23 * traj_ld] = Vgrid;
    row[(int64_t)24 * traj_ld] = Sinsol;
  }
}
#else
template <class LN>
inline void record_substep_split(const LN&, double* traj, int64_t traj_ld, int s, const Split3::Vec& y, double Vgrid,
                                 double Sinsol) {
  double* row = traj + (int64_t)s * 25 * traj_ld;
  for (int i = 0; i < 6; ++i)
    for (int k = 0; k < 3; ++k) row[(int64_t)(6 * k + i) * traj_ld] = y.p[i].v[k];
  for (int i = 0; i < 5; ++i) row[(int64_t)(18 + i) * traj_ld] = y.s[i];
  row[(int64_t)23 * traj_ld] = Vgrid;
  row[(int64_t)24 * traj_ld] = Sinsol;
}
#endif

// One env step of one env on three lanes (every lane runs the same bookkeeping; see advance_env for
// the reference line numbers).  The warp stays converged through the integration: when at least one env
// of the warp steps, ALL its lanes integrate -- an env that must not step (done, bad action, padding)
// works on values that `restore(r)` afterwards reloads from memory (rare, cold).  That keeps every
// shuffle on the compile-time full mask; only the out-of-line slow path runs under a ballot mask.
template <class Restore>
PVDER_DEV bool advance_env_split(const Lanes3& ln, const pvder_env_config& cfg, const RodasTab& tab, EnvRegsSplit& r,
                                 int act, bool active, const double* vtab, const double* stab, int64_t ld, int64_t e,
                                 uint32_t env_glob, Outputs& o, int& done_out, int& hist_inc, bool& hist_clear,
                                 Restore restore, double* traj = nullptr, int64_t traj_ld = 0) {
  using S = Split3;
  const Params& par = cfg.par;
  const S::Consts kc = S::consts(ln);
  hist_inc = -1;
  hist_clear = false;
  // (selects throughout: see the note on group-divergent branches at the sub-step loop)
  const bool stepping = active && !r.done;          // PVDER_env.py:145-154: step after done is a no-op
  const bool bad_action = stepping && (unsigned)act >= (unsigned)PVDER_N_ACTIONS;   // PVDER_env.py:201
  const bool run = stepping && !bad_action;
  // a rejected action changes nothing; the next valid step clears the flag
  r.status = bad_action ? PVDER_STATUS_BAD_ACTION : ((run && r.status == PVDER_STATUS_BAD_ACTION) ? PVDER_STATUS_OK : r.status);
  hist_inc = run ? act : -1;                        // env_utilities.py:25-30
  r.steps += run ? 1 : 0;                           // PVDER_env.py:156
  {
    const double dQ = (act == 1) ? cfg.delQ_pu : ((act == 2) ? -cfg.delQ_pu : 0.0);
    const double dV = (act == 3) ? cfg.delVdc_pu : ((act == 4) ? -cfg.delVdc_pu : 0.0);
    const double q = __dadd_rn(r.Qref, dQ), v = __dadd_rn(r.Vdcref, dV);
    r.Qref = run ? q : r.Qref;                      // PVDER_env.py:225
    r.Vdcref = run ? v : r.Vdcref;                  // PVDER_env.py:229
  }
  const bool any_run = ln.any_warp(run);                     // warp-uniform
  if (any_run) {
    const int status_in = r.status;
    Aux base;
    Inputs in_s = make_inputs(cfg, r.Vgrid, r.Qref, r.Vdcref, r.Sinsol);   // changes only when an event fires
    S::In in = S::inputs(ln, kc, in_s);
    aux_exact_sv(par, in_s, r.y.s[4], r.y.s[0], base);
    // One loop over the half-cycle sub-steps, the warp converged through its hot step (whose shuffles use the compile-time
    // full mask).  A group whose sub-step is refined (fine-step level > 0: the inputs changed at its start, the PLL pull-in
    // after reset, base_level) or left the range of the incremental side-inputs takes the out-of-line path under the
    // ballot mask of the groups at the same level (ros_slow_split: convergence contract there); in the hot step such a
    // group only keeps the warp company (discard).  This is the loop shape of round 1, which is the one that has been
    // run through every failure mode on the device; the call-free segment loop of the one-thread kernels (advance_env)
    // measured 8 % faster here but could not be made to survive an env that blows up (see ros_slow_split).
    const int k0 = r.k;
    int j_next = (r.k < cfg.ev_start_k) ? 0 : (r.k - cfg.ev_start_k) / cfg.ev_step_k + 1;
    int next_k = cfg.ev_start_k + j_next * cfg.ev_step_k;
    const bool ev_here = r.k >= cfg.ev_start_k && (r.k - cfg.ev_start_k) % cfg.ev_step_k == 0 &&
                         (r.k - cfg.ev_start_k) / cfg.ev_step_k < cfg.ev_count;
    int lvl_in = ((cfg.refine_on_action && act != 0 && run) || ev_here) ? cfg.refine_input_level : 0;
    bool dead = false;   // the state went non-finite or absurd: the env is finished (status NONFINITE below) and parked on
                         // the finite reset state, so that its lanes execute what healthy lanes execute from then on
    for (int s = 0; s < cfg.n_sub_per_step; ++s) {
      const int lvl_st = (r.k < cfg.startup_substeps) ? cfg.startup_level : cfg.base_level;
      const int lvl = dead ? 0 : (lvl_in > lvl_st ? lvl_in : lvl_st);
      lvl_in = 0;
      bool m_over;
      const S::Gains g = S::gains(ln, par, kc, in, r.y, tab.luc, m_over);
      bool clamped = g.any;
      // warp-uniform choice: with no clamp active anywhere in the warp the gain-dependent coefficients
      // come from the constant bank (fewer live registers, no spills in the common case)
      const bool skip = lvl != 0 || dead;
      const bool ok = ln.any_warp(g.any) ? ros_core_split<false, false>(ln, r.y, cfg, in_s, in, kc, tab, g, base, skip)
                                         : ros_core_split<false, true>(ln, r.y, cfg, in_s, in, kc, tab, g, base, skip);
      const bool slow = !ok && !dead;
      if (ln.any_warp(slow)) {
#pragma unroll 1
        for (int L = 0; L <= PVDER_FINE_LEVELS; ++L) {
          const bool mine = slow && lvl == L;
          const LanesT<true> lx = ln.sub(mine);
          if (mine) {
            const SplitStepResult res = ros_slow_split(lx, r.y, &cfg, in_s, in, kc, &tab, g, base, L);
            r.y = res.y;
            base = res.base;
            r.exact += 1 << L;
            clamped |= res.clamped != 0;
          }
        }
#ifndef PVDER_SPLIT_NO_QUARANTINE
        // QUARANTINE (group-wide decision, after a slow sub-step: that is where a runaway state shows up first).  A
        // state that is non-finite or absurd (|Vdc| > 1e3 pu, |delta| > 1e8 rad) cannot recover.
        auto nf = vnonfinite(r.y.p[0]);
#pragma unroll
        for (int i = 1; i < 6; ++i) nf = vor(nf, vnonfinite(r.y.p[i]));
        bool nfs = !(fabs(r.y.s[0]) < 1e3) || !(fabs(r.y.s[4]) < 1e8);
#pragma unroll
        for (int i = 1; i < 4; ++i) nfs |= !(bool)isfinite(r.y.s[i]);
        // (both votes by every lane: `a || b` would skip the second one in the groups where the first is true -- a vote
        // that only some lanes of its mask execute is exactly what hung this kernel on B200)
        const bool nf_any = ln.any3(nf), nfs_any = ln.any3(nfs);
        const bool failed_now = slow && (nf_any || nfs_any);
        if (ln.any_warp(failed_now)) {
          S::Vec y0;
          double q0, q1, q2, q3;
          init_env_split(ln, cfg, y0, q0, q1, q2, q3);
#pragma unroll
          for (int i = 0; i < 6; ++i) r.y.p[i] = vsel_group(failed_now, y0.p[i], r.y.p[i]);
#pragma unroll
          for (int i = 0; i < 5; ++i) r.y.s[i] = failed_now ? y0.s[i] : r.y.s[i];
          Aux b0;
          aux_exact_sv(par, in_s, r.y.s[4], r.y.s[0], b0);      // every lane, at its own (sane) state
          base.sn = failed_now ? b0.sn : base.sn;
          base.cs = failed_now ? b0.cs : base.cs;
          base.E = failed_now ? b0.E : base.E;
          dead = dead || failed_now;
        }
#endif
      }
      if (clamped && !dead) r.windup += 1;
      if (traj && run) record_substep_split(ln, traj, traj_ld, r.k - k0, r.y, r.Vgrid, r.Sinsol);   // not the keep-converged dummy work
      r.k += 1;
      if (r.k == next_k && j_next < cfg.ev_count) {
        apply_event(cfg, vtab, stab, ld, e, env_glob, (uint32_t)r.episode, j_next, r.Vgrid, r.Sinsol);
        in_s = make_inputs(cfg, r.Vgrid, r.Qref, r.Vdcref, r.Sinsol);
        in = S::inputs(ln, kc, in_s);
        j_next += 1;
        next_k += cfg.ev_step_k;
        lvl_in = cfg.refine_input_level;
      }
    }
    auto bad = vnonfinite(r.y.p[0]);
#pragma unroll
    for (int i = 1; i < 6; ++i) bad = vor(bad, vnonfinite(r.y.p[i]));
    bool nonfinite = false;
#pragma unroll
    for (int i = 0; i < 5; ++i) nonfinite |= !(bool)isfinite(r.y.s[i]);
    {
      const bool bad_any = ln.any3(bad), nf_any = ln.any3(nonfinite);   // both votes by every lane (no short-circuit)
      nonfinite = bad_any || nf_any || dead;
    }
    r.status = nonfinite ? PVDER_STATUS_NONFINITE : r.status;
    if (ln.any_warp(!run)) {   // an env that was only keeping the warp company: undo (every lane reloads, the others discard)
      EnvRegsSplit q = r;
      restore(q);
      q.status = status_in;
      const bool back = !run;
#pragma unroll
      for (int i = 0; i < 6; ++i) r.y.p[i] = vsel_group(back, q.y.p[i], r.y.p[i]);
#pragma unroll
      for (int i = 0; i < 5; ++i) r.y.s[i] = back ? q.y.s[i] : r.y.s[i];
      r.Qref = back ? q.Qref : r.Qref; r.Vdcref = back ? q.Vdcref : r.Vdcref;
      r.Vgrid = back ? q.Vgrid : r.Vgrid; r.Sinsol = back ? q.Sinsol : r.Sinsol;
      r.ret = back ? q.ret : r.ret; r.last_reward = back ? q.last_reward : r.last_reward;
      r.k = back ? q.k : r.k; r.steps = back ? q.steps : r.steps; r.episode = back ? q.episode : r.episode;
      r.status = back ? q.status : r.status; r.done = back ? q.done : r.done;
      r.windup = back ? q.windup : r.windup; r.exact = back ? q.exact : r.exact;
    }
  }

  compute_outputs_split(ln, cfg, kc, r.y, r.Qref, r.Vdcref, r.Vgrid, r.Sinsol, r.k, o);
  // (selects, not branches: the lanes of a warp stay converged up to the vote below whatever their envs did)
  const bool failed = run && r.status == PVDER_STATUS_NONFINITE;   // PVDER_env.py:170-172: -100, episode ends
  o.reward = failed ? -100.0 : (run ? o.reward : r.last_reward);   // not run: cached tuple, PVDER_env.py:196
  o.reward_i = failed ? -100 : (run ? o.reward_i : (int)r.last_reward);
  done_out = (failed || (run && r.k >= cfg.done_substep)) ? 1 : r.done;   // PVDER_env.py:183
  r.last_reward = run ? o.reward : r.last_reward;
  r.ret += run ? o.reward : 0.0;                // env_utilities.py:32-38
  r.done = run ? done_out : r.done;
  // Auto-reset (vector-env convention: the observation is the first of the new episode, reward/done are the final
  // ones).  The output sums need the whole warp, so everything happens under a warp-wide vote and is selected per env.
  const bool reset_now = run && done_out && cfg.auto_reset;
  if (ln.any_warp(reset_now)) {
    S::Vec y0;
    double Q0, V0, G0, S0;
    init_env_split(ln, cfg, y0, Q0, V0, G0, S0);
    const int ep = r.episode + 1;
    if (cfg.ev_start_k == 0 && cfg.ev_count > 0)
      apply_event(cfg, vtab, stab, ld, e, env_glob, (uint32_t)ep, 0, G0, S0);
#pragma unroll
    for (int i = 0; i < 6; ++i) r.y.p[i] = vsel_group(reset_now, y0.p[i], r.y.p[i]);
#pragma unroll
    for (int i = 0; i < 5; ++i) r.y.s[i] = reset_now ? y0.s[i] : r.y.s[i];
    r.Qref = reset_now ? Q0 : r.Qref; r.Vdcref = reset_now ? V0 : r.Vdcref;
    r.Vgrid = reset_now ? G0 : r.Vgrid; r.Sinsol = reset_now ? S0 : r.Sinsol;
    r.episode = reset_now ? ep : r.episode;
    r.k = reset_now ? 0 : r.k; r.steps = reset_now ? 0 : r.steps; r.done = reset_now ? 0 : r.done;
    r.ret = reset_now ? 0.0 : r.ret; r.status = reset_now ? PVDER_STATUS_OK : r.status;
    r.windup = reset_now ? 0 : r.windup; r.exact = reset_now ? 0 : r.exact;
    hist_inc = reset_now ? -1 : hist_inc;
    hist_clear = reset_now;
    Outputs o2;
    compute_outputs_split(ln, cfg, kc, r.y, r.Qref, r.Vdcref, r.Vgrid, r.Sinsol, r.k, o2);
#pragma unroll
    for (int j = 0; j < PVDER_OBS_DIM; ++j) o.obs[j] = reset_now ? o2.obs[j] : o.obs[j];
  }
  return run;
}

}  // namespace pvder
