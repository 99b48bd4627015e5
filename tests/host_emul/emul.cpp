// TEST INFRASTRUCTURE ONLY.  Plain-C++ build of the per-environment device logic
// (csrc/pvder_env_step.cuh + the generated model headers) so the generated right-hand side,
// symbolic LU, Rodas4 stepper, event draw and output/reward code can be checked against the
// oracle on a machine without a GPU.  Never loaded by the product package.
#include <cstring>

#include "../../gym-solarpvder-environment_b200/csrc/pvder_common.cuh"
#include "../../gym-solarpvder-environment_b200/csrc/pvder_model_1ph.cuh"
#include "../../gym-solarpvder-environment_b200/csrc/pvder_model_3ph.cuh"
#include "../../gym-solarpvder-environment_b200/csrc/pvder_model_3ph_bal.cuh"
#include "../../gym-solarpvder-environment_b200/csrc/pvder_env_step.cuh"

using namespace pvder;

template <class M, bool AUTO3 = false>
static void step_all(const pvder_env_config& cfg, double* sd, int32_t* si, int64_t ld, const int32_t* action,
                     const double* vtab, const double* stab, double* obs64, double* reward, int32_t* reward_i,
                     uint8_t* done, int64_t n, int64_t off) {
  constexpr int NS = M::NS_STORE;
  const RodasTab tab = make_rodas_tab<M>(cfg.par, cfg.substeps_per_sec * (double)cfg.micro);
  for (int64_t e = 0; e < n; ++e) {
    EnvRegs<M> r;
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
    bool run;
    if constexpr (AUTO3)
      run = advance_env_auto3(cfg, tab, r, action[e], true, vtab, stab, ld, e, (uint32_t)(off + e), o, done_out, hist_inc,
                              hist_clear);
    else
      run = advance_env<M>(cfg, tab, r, action[e], true, vtab, stab, ld, e, (uint32_t)(off + e), o, done_out, hist_inc,
                           hist_clear);
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
  Inputs in{inp4[0], inp4[1], inp4[2], inp4[3]};
  Aux ax;
  aux_exact<M>(cfg.par, in, y, ax);
  double gn[M::NFRZ];
  make_gains<M>(cfg.par, frz, gn);
  M::rhs(y, cfg.par, in, ax, gn, ff);
  for (int i = 0; i < M::NS; ++i) f[i] = ff[i];
}

template <class M>
static void wsolve_one(const pvder_env_config& cfg, const double* yin, const double* inp4, unsigned frz, double ghinv,
                       double* b) {
  double y[M::NS], bb[M::NS];
  for (int i = 0; i < M::NS; ++i) { y[i] = yin[i]; bb[i] = b[i]; }
  Inputs in{inp4[0], inp4[1], inp4[2], inp4[3]};
  typename M::LU lu;
  Aux ax;
  aux_exact<M>(cfg.par, in, y, ax);
  double luc[16];
  M::lu_consts(cfg.par, ghinv, luc);
  double gn[M::NFRZ];
  make_gains<M>(cfg.par, frz, gn);
  M::factor(y, cfg.par, in, ax, gn, ghinv, luc, lu);
  M::solve(lu, luc, bb);
  for (int i = 0; i < M::NS; ++i) b[i] = bb[i];
}

template <class M>
static unsigned frz_one(const pvder_env_config& cfg, const double* yin, const double* inp4) {
  double y[M::NS];
  for (int i = 0; i < M::NS; ++i) y[i] = yin[i];
  Inputs in{inp4[0], inp4[1], inp4[2], inp4[3]};
  bool m_over;
  return freeze_bits<M>(y, cfg.par, in, m_over);
}

extern "C" {

int emul_step(const pvder_env_config* cfg, double* sd, int32_t* si, int64_t ld, const int32_t* action,
              const double* vtab, const double* stab, double* obs64, double* reward, int32_t* reward_i, uint8_t* done,
              int64_t n, int64_t off) {
  if (cfg->phases == 1) step_all<Model1ph>(*cfg, sd, si, ld, action, vtab, stab, obs64, reward, reward_i, done, n, off);
  else if (cfg->balanced3 == PVDER_3PH_BALANCED)
    step_all<Model3phBal>(*cfg, sd, si, ld, action, vtab, stab, obs64, reward, reward_i, done, n, off);
  else if (cfg->balanced3 == PVDER_3PH_AUTO)
    step_all<Model3ph, true>(*cfg, sd, si, ld, action, vtab, stab, obs64, reward, reward_i, done, n, off);
  else step_all<Model3ph>(*cfg, sd, si, ld, action, vtab, stab, obs64, reward, reward_i, done, n, off);
  return 0;
}

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
