// TEST INFRASTRUCTURE ONLY.  Plain-C++ build of the per-environment device logic
// (csrc/pvder_env_step.cuh + the generated model headers) so the generated right-hand side,
// symbolic LU, Rosenbrock stepper, event draw and output/reward code can be checked against the
// oracle on a machine without a GPU.  Never loaded by the product package.
#include <cstring>

#include "../../gym-solarpvder-environment_b200/csrc/pvder_common.cuh"
#include "../../gym-solarpvder-environment_b200/csrc/pvder_model_1ph.cuh"
#include "../../gym-solarpvder-environment_b200/csrc/pvder_model_3ph.cuh"
#include "../../gym-solarpvder-environment_b200/csrc/pvder_model_3ph_bal.cuh"
#include "../../gym-solarpvder-environment_b200/csrc/pvder_env_step.cuh"
#include "../../gym-solarpvder-environment_b200/csrc/pvder_split3.cuh"

using namespace pvder;

static double* g_traj = nullptr;   // optional trajectory buffer [n_sub][NS_STORE + 2][n] of the next emul_step call

static void step_one_split(const pvder_env_config& cfg, double* sd, int32_t* si, int64_t ld, const int32_t* action,
                           const double* vtab, const double* stab, double* obs64, double* reward, int32_t* reward_i,
                           uint8_t* done, int64_t n, int64_t off, int64_t e);

// AUTO3 (M = Model3phBal): like the CUDA launch pair of PVDER_3PH_AUTO -- a balanced stored state is stepped on phase
// a; an env that is not balanced, or whose duty-cycle clamp engages during the step, is redone by the lane-split model.
template <class M, bool AUTO3 = false>
static void step_all(const pvder_env_config& cfg, double* sd, int32_t* si, int64_t ld, const int32_t* action,
                     const double* vtab, const double* stab, double* obs64, double* reward, int32_t* reward_i,
                     uint8_t* done, int64_t n, int64_t off) {
  constexpr int NS = M::NS_STORE;
  const RodasTab tab = make_rodas_tab<M>(cfg.par, cfg.substeps_per_sec);
  for (int64_t e = 0; e < n; ++e) {
    EnvRegs<M> r;
    if constexpr (AUTO3) {
      double z[23];
      for (int i = 0; i < 23; ++i) z[i] = sd[(int64_t)i * ld + e];
      if (!is_balanced(z)) {
        step_one_split(cfg, sd, si, ld, action, vtab, stab, obs64, reward, reward_i, done, n, off, e);
        continue;
      }
    }
    load_state<M>(sd, ld, e, r.y);
    r.Qref = sd[(int64_t)PVDER_SD_QREF(NS) * ld + e];
    r.Vdcref = sd[(int64_t)PVDER_SD_VDCREF(NS) * ld + e];
    r.Vgrid = sd[(int64_t)PVDER_SD_VGRID(NS) * ld + e];
    r.Sinsol = sd[(int64_t)PVDER_SD_SINSOL(NS) * ld + e];
    r.ret = sd[(int64_t)PVDER_SD_RETURN(NS) * ld + e];
    r.last_reward = sd[(int64_t)PVDER_SD_REWARD(NS) * ld + e];
    r.k = si[(int64_t)PVDER_SI_K * ld + e];
    r.steps = si[(int64_t)PVDER_SI_STEPS * ld + e];
    r.episode = si[(int64_t)PVDER_SI_EPISODE * ld + e];
    r.status = si[(int64_t)PVDER_SI_STATUS * ld + e];
    r.done = si[(int64_t)PVDER_SI_DONE * ld + e];
    r.windup = si[(int64_t)PVDER_SI_WINDUP * ld + e];
    r.exact = si[(int64_t)PVDER_SI_EXACT * ld + e];
    Outputs o;
    int done_out, hist_inc;
    bool hist_clear;
    const int status_in = r.status;
    const bool run = advance_env<M>(cfg, tab, r, action[e], true, vtab, stab, ld, e, (uint32_t)(off + e), o, done_out,
                                    hist_inc, hist_clear, g_traj ? g_traj + e : nullptr, n);
    if constexpr (AUTO3) {
      if (run && status_in != PVDER_STATUS_UNBALANCED && r.status == PVDER_STATUS_UNBALANCED) {
        step_one_split(cfg, sd, si, ld, action, vtab, stab, obs64, reward, reward_i, done, n, off, e);
        continue;
      }
    }
    if (reward) reward[e] = o.reward;
    if (reward_i) reward_i[e] = o.reward_i;
    if (done) done[e] = (uint8_t)done_out;
    if (run) {
      store_state<M>(sd, ld, e, r.y);
      sd[(int64_t)PVDER_SD_QREF(NS) * ld + e] = r.Qref;
      sd[(int64_t)PVDER_SD_VDCREF(NS) * ld + e] = r.Vdcref;
      sd[(int64_t)PVDER_SD_VGRID(NS) * ld + e] = r.Vgrid;
      sd[(int64_t)PVDER_SD_SINSOL(NS) * ld + e] = r.Sinsol;
      sd[(int64_t)PVDER_SD_RETURN(NS) * ld + e] = r.ret;
      sd[(int64_t)PVDER_SD_REWARD(NS) * ld + e] = r.last_reward;
      si[(int64_t)PVDER_SI_K * ld + e] = r.k;
      si[(int64_t)PVDER_SI_STEPS * ld + e] = r.steps;
      si[(int64_t)PVDER_SI_EPISODE * ld + e] = r.episode;
      si[(int64_t)PVDER_SI_DONE * ld + e] = r.done;
      si[(int64_t)PVDER_SI_WINDUP * ld + e] = r.windup;
      si[(int64_t)PVDER_SI_EXACT * ld + e] = r.exact;
      if (hist_inc >= 0) si[(int64_t)(PVDER_SI_HIST + hist_inc) * ld + e] += 1;
      if (hist_clear)
        for (int h = 0; h < PVDER_N_ACTIONS; ++h) si[(int64_t)(PVDER_SI_HIST + h) * ld + e] = 0;
    }
    si[(int64_t)PVDER_SI_STATUS * ld + e] = r.status;
    if (obs64)
      for (int j = 0; j < PVDER_OBS_DIM; ++j) obs64[e * PVDER_OBS_DIM + j] = o.obs[j];
  }
}

template <class M>
static void reset_all(const pvder_env_config& cfg, double* sd, int32_t* si, int64_t ld, int init, double* obs64,
                      int64_t n, int64_t off) {
  constexpr int NS = M::NS;
  for (int64_t e = 0; e < n; ++e) {
    double y[NS], Qref, Vdcref, Vgrid, Sinsol;
    init_env<M>(cfg, y, Qref, Vdcref, Vgrid, Sinsol);
    const int episode = init ? 0 : si[(int64_t)PVDER_SI_EPISODE * ld + e] + 1;
    if (cfg.ev_start_k == 0 && cfg.ev_count > 0)
      apply_event(cfg, nullptr, nullptr, ld, e, (uint32_t)(off + e), (uint32_t)episode, 0, Vgrid, Sinsol);
    for (int i = 0; i < NS; ++i) sd[(int64_t)i * ld + e] = y[i];
    sd[(int64_t)PVDER_SD_QREF(NS) * ld + e] = Qref;
    sd[(int64_t)PVDER_SD_VDCREF(NS) * ld + e] = Vdcref;
    sd[(int64_t)PVDER_SD_VGRID(NS) * ld + e] = Vgrid;
    sd[(int64_t)PVDER_SD_SINSOL(NS) * ld + e] = Sinsol;
    sd[(int64_t)PVDER_SD_RETURN(NS) * ld + e] = 0.0;
    sd[(int64_t)PVDER_SD_REWARD(NS) * ld + e] = 0.0;
    for (int f = 0; f < PVDER_SI_FIELDS; ++f) si[(int64_t)f * ld + e] = 0;
    si[(int64_t)PVDER_SI_EPISODE * ld + e] = episode;
    Outputs o;
    compute_outputs<M>(cfg, y, Qref, Vdcref, Vgrid, Sinsol, 0, o);
    if (obs64)
      for (int j = 0; j < PVDER_OBS_DIM; ++j) obs64[e * PVDER_OBS_DIM + j] = o.obs[j];
  }
}

template <class M>
static void rhs_one(const pvder_env_config& cfg, const double* yin, const double* inp4, unsigned frz, double* f) {
  double y[M::NS], ff[M::NS];
  for (int i = 0; i < M::NS; ++i) y[i] = yin[i];
  Inputs in{inp4[0], inp4[1], inp4[2], inp4[3], inp4[0] * cfg.vg_ratio_b, inp4[0] * cfg.vg_ratio_c,
            (inp4[3] + cfg.par.np_irs) * cfg.par.pv_scale};
  Aux ax;
  aux_exact<M>(cfg.par, in, y, ax);
  // h*gamma = 1: the unit-pivot rows come out unscaled (tests compare with the oracle's f)
  const RodasTab tab = make_rodas_tab<M>(cfg.par, RG);
  double gn[M::NGAIN];
  make_gains<M>(cfg.par, tab, frz, gn);
  M::rhs(y, cfg.par, in, ax, gn, ff);
  for (int i = 0; i < M::NS; ++i) f[i] = ff[i];
}

template <class M>
static void wsolve_one(const pvder_env_config& cfg, const double* yin, const double* inp4, unsigned frz, double ghinv,
                       double* b) {
  double y[M::NS], bb[M::NS];
  for (int i = 0; i < M::NS; ++i) { y[i] = yin[i]; bb[i] = b[i]; }
  Inputs in{inp4[0], inp4[1], inp4[2], inp4[3], inp4[0] * cfg.vg_ratio_b, inp4[0] * cfg.vg_ratio_c,
            (inp4[3] + cfg.par.np_irs) * cfg.par.pv_scale};
  typename M::LU lu;
  Aux ax;
  aux_exact<M>(cfg.par, in, y, ax);
  // the solve works on the scaled system: unit-pivot rows of the right-hand side are pre-scaled by h*gamma
  const RodasTab tab = make_rodas_tab<M>(cfg.par, ghinv * RG);
  double gn[M::NGAIN];
  make_gains<M>(cfg.par, tab, frz, gn);
  for (int i = 0; i < M::NS; ++i)
    if (M::unit_row(i)) bb[i] *= 1.0 / tab.ghinv;
  M::factor(y, cfg.par, in, ax, gn, tab.ghinv, tab.luc, lu);
  M::solve(lu, tab.luc, bb);
  for (int i = 0; i < M::NS; ++i) b[i] = bb[i];
}

template <class M>
static unsigned frz_one(const pvder_env_config& cfg, const double* yin, const double* inp4) {
  double y[M::NS];
  for (int i = 0; i < M::NS; ++i) y[i] = yin[i];
  Inputs in{inp4[0], inp4[1], inp4[2], inp4[3], inp4[0] * cfg.vg_ratio_b, inp4[0] * cfg.vg_ratio_c,
            (inp4[3] + cfg.par.np_irs) * cfg.par.pv_scale};
  bool m_over;
  return freeze_bits<M>(y, cfg.par, in, m_over);
}


// ---- lane-split three-phase model (pvder_split3.cuh) with the three lanes emulated by V3 ----------
static void load_split(const double* sd, int64_t ld, int64_t e, Split3::Vec& y) {
  for (int i = 0; i < 6; ++i) y.p[i] = V3(sd[(int64_t)i * ld + e], sd[(int64_t)(6 + i) * ld + e], sd[(int64_t)(12 + i) * ld + e]);
  for (int i = 0; i < 5; ++i) y.s[i] = sd[(int64_t)(18 + i) * ld + e];
}

static void store_split(double* sd, int64_t ld, int64_t e, const Split3::Vec& y) {
  for (int i = 0; i < 6; ++i)
    for (int k = 0; k < 3; ++k) sd[(int64_t)(6 * k + i) * ld + e] = y.p[i].v[k];
  for (int i = 0; i < 5; ++i) sd[(int64_t)(18 + i) * ld + e] = y.s[i];
}

static void step_one_split(const pvder_env_config& cfg, double* sd, int32_t* si, int64_t ld, const int32_t* action,
                           const double* vtab, const double* stab, double* obs64, double* reward, int32_t* reward_i,
                           uint8_t* done, int64_t n, int64_t off, int64_t e) {
  constexpr int NS = 23;
  const RodasTab tab = make_rodas_tab<Split3>(cfg.par, cfg.substeps_per_sec);
  const Lanes3 ln;
  {
    EnvRegsSplit r;
    load_split(sd, ld, e, r.y);
    r.Qref = sd[(int64_t)PVDER_SD_QREF(NS) * ld + e];
    r.Vdcref = sd[(int64_t)PVDER_SD_VDCREF(NS) * ld + e];
    r.Vgrid = sd[(int64_t)PVDER_SD_VGRID(NS) * ld + e];
    r.Sinsol = sd[(int64_t)PVDER_SD_SINSOL(NS) * ld + e];
    r.ret = sd[(int64_t)PVDER_SD_RETURN(NS) * ld + e];
    r.last_reward = sd[(int64_t)PVDER_SD_REWARD(NS) * ld + e];
    r.k = si[(int64_t)PVDER_SI_K * ld + e];
    r.steps = si[(int64_t)PVDER_SI_STEPS * ld + e];
    r.episode = si[(int64_t)PVDER_SI_EPISODE * ld + e];
    r.status = si[(int64_t)PVDER_SI_STATUS * ld + e];
    r.done = si[(int64_t)PVDER_SI_DONE * ld + e];
    r.windup = si[(int64_t)PVDER_SI_WINDUP * ld + e];
    r.exact = si[(int64_t)PVDER_SI_EXACT * ld + e];
    Outputs o;
    int done_out, hist_inc;
    bool hist_clear;
    const bool run = advance_env_split(ln, cfg, tab, r, action[e], true, vtab, stab, ld, e, (uint32_t)(off + e), o,
                                       done_out, hist_inc, hist_clear, [](EnvRegsSplit&) {},
                                       g_traj ? g_traj + e : nullptr, n);
    if (reward) reward[e] = o.reward;
    if (reward_i) reward_i[e] = o.reward_i;
    if (done) done[e] = (uint8_t)done_out;
    if (run) {
      store_split(sd, ld, e, r.y);
      sd[(int64_t)PVDER_SD_QREF(NS) * ld + e] = r.Qref;
      sd[(int64_t)PVDER_SD_VDCREF(NS) * ld + e] = r.Vdcref;
      sd[(int64_t)PVDER_SD_VGRID(NS) * ld + e] = r.Vgrid;
      sd[(int64_t)PVDER_SD_SINSOL(NS) * ld + e] = r.Sinsol;
      sd[(int64_t)PVDER_SD_RETURN(NS) * ld + e] = r.ret;
      sd[(int64_t)PVDER_SD_REWARD(NS) * ld + e] = r.last_reward;
      si[(int64_t)PVDER_SI_K * ld + e] = r.k;
      si[(int64_t)PVDER_SI_STEPS * ld + e] = r.steps;
      si[(int64_t)PVDER_SI_EPISODE * ld + e] = r.episode;
      si[(int64_t)PVDER_SI_DONE * ld + e] = r.done;
      si[(int64_t)PVDER_SI_WINDUP * ld + e] = r.windup;
      si[(int64_t)PVDER_SI_EXACT * ld + e] = r.exact;
      if (hist_inc >= 0) si[(int64_t)(PVDER_SI_HIST + hist_inc) * ld + e] += 1;
      if (hist_clear)
        for (int h = 0; h < PVDER_N_ACTIONS; ++h) si[(int64_t)(PVDER_SI_HIST + h) * ld + e] = 0;
    }
    si[(int64_t)PVDER_SI_STATUS * ld + e] = r.status;
    if (obs64)
      for (int j = 0; j < PVDER_OBS_DIM; ++j) obs64[e * PVDER_OBS_DIM + j] = o.obs[j];
  }
}

static void step_all_split(const pvder_env_config& cfg, double* sd, int32_t* si, int64_t ld, const int32_t* action,
                           const double* vtab, const double* stab, double* obs64, double* reward, int32_t* reward_i,
                           uint8_t* done, int64_t n, int64_t off) {
  for (int64_t e = 0; e < n; ++e) step_one_split(cfg, sd, si, ld, action, vtab, stab, obs64, reward, reward_i, done, n, off, e);
}

// rhs (mode 0) or W^-1 b (mode 1) of the lane-split model at a 23-state point; frz = freeze mask bits
// in the order of Model3ph (per phase xR,xI,uR,uI; then xDC, xQ)
static bool free_path = true;   // frz == 0: exercise the FREE (constant-bank gains) instantiation or the general one

static void split_one(const pvder_env_config& cfg, const double* yin, const double* inp4, unsigned frz, double ghinv,
                      int mode, double* out) {
  const Lanes3 ln;
  Split3::Vec y, b;
  load_split(yin, 1, 0, y);
  Inputs in_s{inp4[0], inp4[1], inp4[2], inp4[3], inp4[0] * cfg.vg_ratio_b, inp4[0] * cfg.vg_ratio_c,
            (inp4[3] + cfg.par.np_irs) * cfg.par.pv_scale};
  const Split3::Consts kc = Split3::consts(ln);
  const Split3::In in = Split3::inputs(ln, kc, in_s);
  Aux ax;
  aux_exact_sv(cfg.par, in_s, y.s[4], y.s[0], ax);
  double luc[16];
  Split3::lu_consts(cfg.par, ghinv, luc);
  Split3::Gains g;
  auto bit = [&](int b_) { return (frz >> b_) & 1u; };
  // mode 0 (rhs): h*gamma = 1 so that the unit-pivot rows come out unscaled; mode 1: the solve works on the scaled
  // system (gains of the unit-pivot rows and their right-hand sides pre-scaled by h*gamma = 1/ghinv)
  if (mode == 0) ghinv = 1.0;
  Split3::lu_consts(cfg.par, ghinv, luc);
  const double hg = 1.0 / ghinv;
  auto gv = [&](int j, double val) { return V3(bit(j) ? 0.0 : val, bit(4 + j) ? 0.0 : val, bit(8 + j) ? 0.0 : val); };
  g.g0 = gv(0, cfg.par.Ki_GCC * hg); g.g1 = gv(1, cfg.par.Ki_GCC * hg); g.g2 = gv(2, cfg.par.wp); g.g3 = gv(3, cfg.par.wp);
  g.duR = V3(1.0) / (g.g2 + V3(ghinv));
  g.duI = V3(1.0) / (g.g3 + V3(ghinv));
  g.g4 = bit(12) ? 0.0 : cfg.par.Ki_DC * hg;
  g.g5 = bit(13) ? 0.0 : cfg.par.Ki_Q * hg;
  g.any = frz != 0;
  const Split3::Pt q = Split3::point(ln, cfg.par, kc, in, ax, y);
  if (mode == 0) {
    if (frz == 0 && free_path) Split3::rhs<true>(cfg.par, kc, ax, g, luc, y, q, b);
    else Split3::rhs<false>(cfg.par, kc, ax, g, luc, y, q, b);
  } else {
    load_split(out, 1, 0, b);
    for (int i = 0; i < 6; ++i)
      if (Split3::unit_p(i)) b.p[i] = b.p[i] * V3(hg);
    for (int i = 0; i < 5; ++i)
      if (Split3::unit_s(i)) b.s[i] *= hg;
    Split3::Fac fac;
    if (frz == 0 && free_path) {
      Split3::factor<true>(ln, cfg.par, kc, in, ax, g, y, q, ghinv, luc, fac);
      Split3::solve<true>(ln, cfg.par, kc, g, fac, y, luc, b);
    } else {
      Split3::factor<false>(ln, cfg.par, kc, in, ax, g, y, q, ghinv, luc, fac);
      Split3::solve<false>(ln, cfg.par, kc, g, fac, y, luc, b);
    }
  }
  store_split(out, 1, 0, b);
}

static unsigned split_freeze_bits(const pvder_env_config& cfg, const double* yin, const double* inp4, double ghinv) {
  const Lanes3 ln;
  Split3::Vec y;
  load_split(yin, 1, 0, y);
  Inputs in_s{inp4[0], inp4[1], inp4[2], inp4[3], inp4[0] * cfg.vg_ratio_b, inp4[0] * cfg.vg_ratio_c,
            (inp4[3] + cfg.par.np_irs) * cfg.par.pv_scale};
  const Split3::Consts kc = Split3::consts(ln);
  const Split3::In in = Split3::inputs(ln, kc, in_s);
  double luc[16];
  Split3::lu_consts(cfg.par, ghinv, luc);
  bool m_over;
  const Split3::Gains g = Split3::gains(ln, cfg.par, kc, in, y, luc, m_over);
  unsigned bits = 0;
  for (int k = 0; k < 3; ++k) {
    if (g.g0.v[k] == 0.0) bits |= 1u << (4 * k);
    if (g.g1.v[k] == 0.0) bits |= 1u << (4 * k + 1);
    if (g.g2.v[k] == 0.0) bits |= 1u << (4 * k + 2);
    if (g.g3.v[k] == 0.0) bits |= 1u << (4 * k + 3);
  }
  if (g.g4 == 0.0) bits |= 1u << 12;
  if (g.g5 == 0.0) bits |= 1u << 13;
  return bits;
}

extern "C" {

int emul_step(const pvder_env_config* cfg, double* sd, int32_t* si, int64_t ld, const int32_t* action,
              const double* vtab, const double* stab, double* obs64, double* reward, int32_t* reward_i, uint8_t* done,
              int64_t n, int64_t off) {
  if (cfg->phases == 1) step_all<Model1ph>(*cfg, sd, si, ld, action, vtab, stab, obs64, reward, reward_i, done, n, off);
  else if (cfg->balanced3 == PVDER_3PH_BALANCED)
    step_all<Model3phBal>(*cfg, sd, si, ld, action, vtab, stab, obs64, reward, reward_i, done, n, off);
  else if (cfg->balanced3 == PVDER_3PH_AUTO)
    step_all<Model3phBal, true>(*cfg, sd, si, ld, action, vtab, stab, obs64, reward, reward_i, done, n, off);
  else if (cfg->balanced3 == PVDER_3PH_SPLIT)
    step_all_split(*cfg, sd, si, ld, action, vtab, stab, obs64, reward, reward_i, done, n, off);
  else step_all<Model3ph>(*cfg, sd, si, ld, action, vtab, stab, obs64, reward, reward_i, done, n, off);
  return 0;
}

void emul_set_traj(double* traj) { g_traj = traj; }

int emul_reset(const pvder_env_config* cfg, double* sd, int32_t* si, int64_t ld, int init, double* obs64, int64_t n,
               int64_t off) {
  if (cfg->phases == 1) reset_all<Model1ph>(*cfg, sd, si, ld, init, obs64, n, off);
  else reset_all<Model3ph>(*cfg, sd, si, ld, init, obs64, n, off);
  return 0;
}

void emul_rhs(const pvder_env_config* cfg, const double* y, const double* inp4, unsigned frz, double* f) {
  if (cfg->phases == 1) rhs_one<Model1ph>(*cfg, y, inp4, frz, f);
  else rhs_one<Model3ph>(*cfg, y, inp4, frz, f);
}

void emul_wsolve(const pvder_env_config* cfg, const double* y, const double* inp4, unsigned frz, double ghinv,
                 double* b) {
  if (cfg->phases == 1) wsolve_one<Model1ph>(*cfg, y, inp4, frz, ghinv, b);
  else wsolve_one<Model3ph>(*cfg, y, inp4, frz, ghinv, b);
}

unsigned emul_freeze_bits(const pvder_env_config* cfg, const double* y, const double* inp4) {
  return cfg->phases == 1 ? frz_one<Model1ph>(*cfg, y, inp4) : frz_one<Model3ph>(*cfg, y, inp4);
}

// The integrator tableau the kernels are compiled with, un-scaled (c_ij and d_4j multiplied back by h):
// out = [PVDER_SCHEME, gamma, a21 a31 a32 a41 a42 a43 a51 a52 a53 a54, c21 c31 c32 c41 c42 c43 c51 .. c54 c61 .. c65,
//        m1 m2 m3 m4, d41 d42, cs21 / c21 (= h gamma)]
void emul_scheme(const pvder_env_config* cfg, double* out) {
  const double hinv = 120.0;
  const RodasTab t = make_rodas_tab<Model1ph>(cfg->par, hinv);
  const double h = 1.0 / hinv;
  int k = 0;
  out[k++] = (double)PVDER_SCHEME;
  out[k++] = RG;
  const double a[10] = {t.a21, t.a31, t.a32, t.a41, t.a42, t.a43, t.a51, t.a52, t.a53, t.a54};
  for (double v : a) out[k++] = v;
  const double c[15] = {t.c21, t.c31, t.c32, t.c41, t.c42, t.c43, t.c51, t.c52, t.c53, t.c54, t.c61, t.c62, t.c63, t.c64, t.c65};
  for (double v : c) out[k++] = v * h;
  out[k++] = t.m1; out[k++] = t.m2; out[k++] = t.m3; out[k++] = t.m4;
  out[k++] = t.d41 * h; out[k++] = t.d42 * h;
  out[k++] = t.cs21 / t.c21;
}

void emul_split_free_path(int on) { free_path = on != 0; }

void emul_split_rhs(const pvder_env_config* cfg, const double* y, const double* inp4, unsigned frz, double* f) {
  split_one(*cfg, y, inp4, frz, 480.0, 0, f);
}

void emul_split_wsolve(const pvder_env_config* cfg, const double* y, const double* inp4, unsigned frz, double ghinv,
                       double* b) {
  split_one(*cfg, y, inp4, frz, ghinv, 1, b);
}

unsigned emul_split_freeze_bits(const pvder_env_config* cfg, const double* y, const double* inp4) {
  return split_freeze_bits(*cfg, y, inp4, 480.0);
}

void emul_events(const pvder_env_config* cfg, const int32_t* episode, double* vtab, double* stab, int64_t ld,
                 int64_t n, int64_t off) {
  for (int64_t e = 0; e < n; ++e) {
    double Vgrid = 1.0, Sinsol = 100.0;
    for (int j = 0; j < cfg->ev_count; ++j) {
      if (cfg->ev_voltage_enable || cfg->ev_insol_enable)
        draw_event(*cfg, (uint32_t)(off + e), episode ? (uint32_t)episode[e] : 0u, (uint32_t)j, Vgrid, Sinsol);
      vtab[(int64_t)j * ld + e] = Vgrid;
      stab[(int64_t)j * ld + e] = Sinsol;
    }
  }
}

}  // extern "C"
