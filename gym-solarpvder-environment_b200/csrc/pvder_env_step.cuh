// Per-environment device logic of the PVDER-v0 step: Rodas4 half-cycle integrator, anti-windup
// mode sampling, event draw, outputs (obs / reward).  Shared by the CUDA kernels
// (pvder_kernels.cu) and -- compiled as plain C++ -- by the CPU-side test harness
// tests/host_emul (test infrastructure only; the product path is the CUDA build).
//
//   advance_env    <- PVDER.step            reference gym_PVDER/envs/PVDER_env.py:138-196
//   compute_outputs<- PVDER.state :531-542, PVDER.reward_calc :231-301
//   draw_event     <- generate_simulation_events :400-411 (pvder create_random_events, A.8)
#pragma once
#include "pvder_common.cuh"

namespace pvder {

// ---- Rodas4 (Hairer & Wanner, Solving ODEs II, sec. IV.10), form  (I/(h g) - J) K_i = f(Y_i) + sum c_ij/h K_j
constexpr double RG = 0.25;
constexpr double A21 = 0.1544000000000000e+01;
constexpr double A31 = 0.9466785280815826e+00, A32 = 0.2557011698983284e+00;
constexpr double A41 = 0.3314825187068521e+01, A42 = 0.2896124015972201e+01, A43 = 0.9986419139977817e+00;
constexpr double A51 = 0.1221224509226641e+01, A52 = 0.6019134481288629e+01, A53 = 0.1253708332932087e+02,
                 A54 = -0.6878860361058950e+00;
constexpr double C21 = -0.5668800000000000e+01;
constexpr double C31 = -0.2430093356833875e+01, C32 = -0.2063599157091915e+00;
constexpr double C41 = -0.1073529058151375e+00, C42 = -0.9594562251023355e+01, C43 = -0.2047028614809616e+02;
constexpr double C51 = 0.7496443313967647e+01, C52 = -0.1024680431464352e+02, C53 = -0.3399990352819905e+02,
                 C54 = 0.1170890893206160e+02;
constexpr double C61 = 0.8083246795921522e+01, C62 = -0.7981132988064893e+01, C63 = -0.3152159432874371e+02,
                 C64 = 0.1631930543123136e+02, C65 = -0.6058818238834054e+01;

template <class M, bool FRZ>
PVDER_DEV void rodas4_step(double (&y)[M::NS], const Params& par, const Inputs& in,
                                            unsigned frz, double hinv) {
  constexpr int NS = M::NS;
  typename M::LU lu;
  M::template factor<FRZ>(y, par, in, frz, hinv * (1.0 / RG), lu);
  double K1[NS], K2[NS], K3[NS], K4[NS], K5[NS], Y[NS];
  // stage 1
  M::template rhs<FRZ>(y, par, in, frz, K1);
  M::solve(lu, K1);
  // stage 2
#pragma unroll
  for (int i = 0; i < NS; ++i) Y[i] = fma(A21, K1[i], y[i]);
  M::template rhs<FRZ>(Y, par, in, frz, K2);
  {
    const double c1 = C21 * hinv;
#pragma unroll
    for (int i = 0; i < NS; ++i) K2[i] = fma(c1, K1[i], K2[i]);
  }
  M::solve(lu, K2);
  // stage 3
#pragma unroll
  for (int i = 0; i < NS; ++i) Y[i] = fma(A32, K2[i], fma(A31, K1[i], y[i]));
  M::template rhs<FRZ>(Y, par, in, frz, K3);
  {
    const double c1 = C31 * hinv, c2 = C32 * hinv;
#pragma unroll
    for (int i = 0; i < NS; ++i) K3[i] = fma(c2, K2[i], fma(c1, K1[i], K3[i]));
  }
  M::solve(lu, K3);
  // stage 4
#pragma unroll
  for (int i = 0; i < NS; ++i) Y[i] = fma(A43, K3[i], fma(A42, K2[i], fma(A41, K1[i], y[i])));
  M::template rhs<FRZ>(Y, par, in, frz, K4);
  {
    const double c1 = C41 * hinv, c2 = C42 * hinv, c3 = C43 * hinv;
#pragma unroll
    for (int i = 0; i < NS; ++i) K4[i] = fma(c3, K3[i], fma(c2, K2[i], fma(c1, K1[i], K4[i])));
  }
  M::solve(lu, K4);
  // stage 5
#pragma unroll
  for (int i = 0; i < NS; ++i)
    Y[i] = fma(A54, K4[i], fma(A53, K3[i], fma(A52, K2[i], fma(A51, K1[i], y[i]))));
  M::template rhs<FRZ>(Y, par, in, frz, K5);
  {
    const double c1 = C51 * hinv, c2 = C52 * hinv, c3 = C53 * hinv, c4 = C54 * hinv;
#pragma unroll
    for (int i = 0; i < NS; ++i)
      K5[i] = fma(c4, K4[i], fma(c3, K3[i], fma(c2, K2[i], fma(c1, K1[i], K5[i]))));
  }
  M::solve(lu, K5);
  // stage 6 (Y6 = Y5 + K5; y_new = Y6 + K6: stiffly accurate)
#pragma unroll
  for (int i = 0; i < NS; ++i) Y[i] += K5[i];
  double K6[NS];
  M::template rhs<FRZ>(Y, par, in, frz, K6);
  {
    const double c1 = C61 * hinv, c2 = C62 * hinv, c3 = C63 * hinv, c4 = C64 * hinv, c5 = C65 * hinv;
#pragma unroll
    for (int i = 0; i < NS; ++i)
      K6[i] = fma(c5, K5[i], fma(c4, K4[i], fma(c3, K3[i], fma(c2, K2[i], fma(c1, K1[i], K6[i])))));
  }
  M::solve(lu, K6);
#pragma unroll
  for (int i = 0; i < NS; ++i) y[i] = Y[i] + K6[i];
}

// pvder's clamping test np.sign(a) == np.sign(b)
PVDER_DEV bool same_sign(double a, double b) {
  const int sa = (a > 0.0) - (a < 0.0);
  const int sb = (b > 0.0) - (b < 0.0);
  return sa == sb;
}

PVDER_DEV void phase_rot(int P, int k, double& rr, double& ri) {
  if (P == 1 || k == 0) { rr = 1.0; ri = 0.0; }
  else if (k == 1) { rr = -0.5; ri = -0.86602540378443864676; }
  else { rr = -0.5; ri = 0.86602540378443864676; }
}

// Anti-windup mode, sampled once per half-cycle sub-step.  Bit order = M::NFRZ
// rows: per phase xR,xI,uR,uI ; then xDC, xQ.
template <class M>
PVDER_DEV unsigned freeze_bits(const double (&y)[M::NS], const Params& par, const Inputs& in) {
  constexpr int P = M::PHASES;
  constexpr int B = 6 * P;
  double Q = 0.0;
  bool m_over = false;
#pragma unroll
  for (int k = 0; k < P; ++k) {
    double rr, ri;
    phase_rot(P, k, rr, ri);
    const double iR = y[6 * k], iI = y[6 * k + 1];
    const double mR = fma(par.Kp_GCC, y[6 * k + 4], y[6 * k + 2]);
    const double mI = fma(par.Kp_GCC, y[6 * k + 5], y[6 * k + 3]);
    m_over |= (mR * mR + mI * mI) > par.m_limit10 * par.m_limit10;
    Q += 0.5 * ((in.vg * ri) * iR - (in.vg * rr) * iI + par.Xt * (iR * iR + iI * iI));
  }
  const double Vdc = y[B], xDC = y[B + 1], xQ = y[B + 2];
  const double irefR = xDC + par.Kp_DC * (in.Vdcref - Vdc);
  const double irefI = xQ - par.Kp_Q * (in.Qref - Q);
  const bool i_over = (irefR * irefR + irefI * irefI) > par.iref_limit * par.iref_limit;
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
    if (same_sign(par.Ki_DC * (in.Vdcref - Vdc), xDC)) bits |= 1u << (4 * P);
    if (same_sign(-par.Ki_Q * (in.Qref - Q), xQ)) bits |= 1u << (4 * P + 1);
  }
  return bits;
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

template <class M>
PVDER_DEV void compute_outputs(const pvder_env_config& cfg, const double (&y)[M::NS], double Qref,
                                                double Vdcref, double Vgrid, double Sinsol, int k, Outputs& o) {
  constexpr int P = M::PHASES;
  constexpr int B = 6 * P;
  const Params& par = cfg.par;
  const double vg = __dmul_rn(Vgrid, par.vgs);
  double Ppcc = 0.0, Qpcc = 0.0, v2 = 0.0, vaR = 0.0, vaI = 0.0;
#pragma unroll
  for (int ph = 0; ph < P; ++ph) {
    double rr, ri;
    phase_rot(P, ph, rr, ri);
    const double jR = y[6 * ph], jI = y[6 * ph + 1];
    const double vkR = __dadd_rn(__dmul_rn(vg, rr), __dadd_rn(__dmul_rn(par.Rt, jR), -__dmul_rn(par.Xt, jI)));
    const double vkI = __dadd_rn(__dmul_rn(vg, ri), __dadd_rn(__dmul_rn(par.Xt, jR), __dmul_rn(par.Rt, jI)));
    Ppcc = __dadd_rn(Ppcc, __dmul_rn(0.5, __dadd_rn(__dmul_rn(vkR, jR), __dmul_rn(vkI, jI))));
    Qpcc = __dadd_rn(Qpcc, __dmul_rn(0.5, __dadd_rn(__dmul_rn(vkI, jR), -__dmul_rn(vkR, jI))));
    v2 = __dadd_rn(v2, __dadd_rn(__dmul_rn(vkR, vkR), __dmul_rn(vkI, vkI)));
    if (ph == 0) { vaR = vkR; vaI = vkI; }
  }
  const double SQRT2 = 1.4142135623730951;
  const double Vrms = (P == 1) ? __ddiv_rn(__dsqrt_rn(v2), SQRT2) : __ddiv_rn(__dsqrt_rn(__ddiv_rn(v2, 3.0)), SQRT2);
  Inputs in{vg, Qref, Vdcref, __dmul_rn(par.np_iph100, __ddiv_rn(Sinsol, 100.0))};
  double Ppv, dPpv;
  ppv_eval(par, in, y[B], Ppv, dPpv);
  o.obs[0] = y[0]; o.obs[1] = y[1]; o.obs[2] = vaR; o.obs[3] = vaI; o.obs[4] = Ppcc; o.obs[5] = Qpcc;
  o.obs[6] = y[B]; o.obs[7] = Ppv; o.obs[8] = Vdcref; o.obs[9] = Qref;
  o.obs[10] = __ddiv_rn(__ddiv_rn((double)k, cfg.substeps_per_sec), cfg.max_sim_time);
  double x, target, hi;
  if (cfg.goal == PVDER_GOAL_VOLTAGE) { x = Vrms; target = par.Vrms_ref; hi = 0.05; }
  else if (cfg.goal == PVDER_GOAL_Q) { x = Qpcc; target = par.q_target; hi = 0.05; }
  else { x = Ppcc; target = par.p_target; hi = 0.03; }
  if (cfg.discrete_reward) {
    if (cfg.goal == PVDER_GOAL_Q && target == 0.0) target = 1e-6;
    const double err = __ddiv_rn(fabs(__dadd_rn(x, -target)), fabs(target));
    o.reward_i = (err <= 0.01) ? 1 : ((err >= hi) ? -5 : -1);
    o.reward = (double)o.reward_i;
  } else {
    const double d = __dadd_rn(x, -target);
    o.reward = -__dmul_rn(d, d);
    o.reward_i = 0;
  }
}


// Registers of one environment.
template <class M>
struct EnvRegs {
  double y[M::NS];
  double Qref, Vdcref, Vgrid, Sinsol, ret, last_reward;
  int k, steps, episode, status, done, windup;
};

template <class M>
PVDER_DEV void init_env(const pvder_env_config& cfg, double (&y)[M::NS], double& Qref, double& Vdcref,
                        double& Vgrid, double& Sinsol) {
#pragma unroll
  for (int i = 0; i < M::NS; ++i) y[i] = cfg.y0[i];
  Qref = cfg.Q_ref0;
  Vdcref = cfg.Vdc_ref0;
  Vgrid = 1.0;
  Sinsol = 100.0;
}

// One env step for one env (everything between the state load and the state store).
// Returns true when the env advanced (state must be written back).  hist_inc: action whose
// histogram counter must be incremented (-1: none); hist_clear: auto-reset happened.
template <class M>
PVDER_DEV bool advance_env(const pvder_env_config& cfg, EnvRegs<M>& r, int act, bool active, const double* vtab,
                           const double* stab, int64_t ld, int64_t e, uint32_t env_glob, Outputs& o, int& done_out,
                           int& hist_inc, bool& hist_clear) {
  constexpr int NS = M::NS;
  const Params& par = cfg.par;
  hist_inc = -1;
  hist_clear = false;
  bool run = active && !r.done;                     // PVDER_env.py:145-154: step after done is a no-op
  if (run && (unsigned)act >= (unsigned)PVDER_N_ACTIONS) {   // PVDER_env.py:201
    r.status = PVDER_STATUS_BAD_ACTION;
    run = false;
  }
  if (run) {
    hist_inc = act;                                          // env_utilities.py:25-30
    r.steps += 1;                                            // PVDER_env.py:156
    const double dQ = (act == 1) ? cfg.delQ_pu : ((act == 2) ? -cfg.delQ_pu : 0.0);
    const double dV = (act == 3) ? cfg.delVdc_pu : ((act == 4) ? -cfg.delVdc_pu : 0.0);
    r.Qref = __dadd_rn(r.Qref, dQ);                          // PVDER_env.py:225
    r.Vdcref = __dadd_rn(r.Vdcref, dV);                      // PVDER_env.py:229
    int j_next = (r.k < cfg.ev_start_k) ? 0 : (r.k - cfg.ev_start_k) / cfg.ev_step_k + 1;
    int next_k = cfg.ev_start_k + j_next * cfg.ev_step_k;
    const double hinv = cfg.substeps_per_sec * (double)cfg.micro;
    for (int s = 0; s < cfg.n_sub_per_step; ++s) {
      Inputs in{__dmul_rn(r.Vgrid, par.vgs), r.Qref, r.Vdcref,
                __dmul_rn(par.np_iph100, __ddiv_rn(r.Sinsol, 100.0))};
      const unsigned frz = freeze_bits<M>(r.y, par, in);
      if (frz) {
        r.windup += 1;
        for (int m = 0; m < cfg.micro; ++m) rodas4_step<M, true>(r.y, par, in, frz, hinv);
      } else {
        for (int m = 0; m < cfg.micro; ++m) rodas4_step<M, false>(r.y, par, in, 0u, hinv);
      }
      r.k += 1;
      if (r.k == next_k && j_next < cfg.ev_count) {
        apply_event(cfg, vtab, stab, ld, e, env_glob, (uint32_t)r.episode, j_next, r.Vgrid, r.Sinsol);
        j_next += 1;
        next_k += cfg.ev_step_k;
      }
    }
    bool finite = true;
#pragma unroll
    for (int i = 0; i < NS; ++i) finite &= (bool)isfinite(r.y[i]);
    if (!finite) r.status = PVDER_STATUS_NONFINITE;
  }

  compute_outputs<M>(cfg, r.y, r.Qref, r.Vdcref, r.Vgrid, r.Sinsol, r.k, o);
  done_out = r.done;
  if (run) {
    if (r.status == PVDER_STATUS_NONFINITE) {   // PVDER_env.py:170-172: intended -100 penalty, episode ends
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
    r.k = 0; r.steps = 0; r.done = 0; r.ret = 0.0; r.status = PVDER_STATUS_OK; r.windup = 0;
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
