/* pvder_b200.h -- C ABI of the B200-native PVDER-v0 hot path (libpvder_b200.so).
 *
 * The reference is pure Python and has no FFI; its "operator interface" for this path is the
 * Gym Env API of gym_PVDER/envs/PVDER_env.py.  Each entry point below names the reference
 * method(s) it replaces (file:line under /root/reference).  All pointers are plain C pointers;
 * "device" pointers are CUDA device addresses owned by the caller, `stream` is a cudaStream_t
 * passed as void* (NULL = default stream).  No call synchronises unless it says so.
 *
 * Per-env state is SoA with leading dimension `ld` (>= n_envs), env index fastest:
 *   sd : double  [PVDER_SD_FIELDS(ns)][ld]   ns = 11 (single-phase) or 23 (three-phase)
 *   si : int32_t [PVDER_SI_FIELDS][ld]
 */
#ifndef PVDER_B200_H_
#define PVDER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVDER_ABI_VERSION 3
#define PVDER_OBS_DIM 11          /* PVDER_env.py:44-51 observed_quantities */
#define PVDER_N_ACTIONS 5         /* PVDER_env.py:50 Discrete(5) */
#define PVDER_MAX_STATES 23
#define PVDER_FINE_LEVELS 3       /* fine-step levels: sub-steps of h/2, h/4, h/8 (refine_input_level, startup_level) */

/* ---- sd field offsets (rows of the double state matrix), ns = number of ODE states ---- */
#define PVDER_SD_Y 0              /* rows 0..ns-1: ODE state, order SURVEY.md A.1, last = delta = wte - w*t */
#define PVDER_SD_QREF(ns) ((ns) + 0)      /* PV_model.Q_ref   (PVDER_env.py:225) */
#define PVDER_SD_VDCREF(ns) ((ns) + 1)    /* PV_model.Vdc_ref (PVDER_env.py:229) */
#define PVDER_SD_VGRID(ns) ((ns) + 2)     /* grid-voltage event value in force */
#define PVDER_SD_SINSOL(ns) ((ns) + 3)    /* insolation event value in force */
#define PVDER_SD_RETURN(ns) ((ns) + 4)    /* episode return (env_utilities.py:32-38) */
#define PVDER_SD_REWARD(ns) ((ns) + 5)    /* last reward (cached tuple of PVDER_env.py:145-152,196) */
#define PVDER_SD_FIELDS(ns) ((ns) + 6)

/* ---- si field offsets ---- */
#define PVDER_SI_K 0              /* half-cycle counter: t = k/120 s  (sim.tStart, PVDER_env.py:180) */
#define PVDER_SI_STEPS 1          /* env._steps (PVDER_env.py:156) */
#define PVDER_SI_EPISODE 2        /* episode counter (RNG key) */
#define PVDER_SI_STATUS 3         /* PVDER_STATUS_* */
#define PVDER_SI_DONE 4           /* env.done (PVDER_env.py:183-184) */
#define PVDER_SI_HIST 5           /* 5 rows: action histogram (env_utilities.py:25-30) */
#define PVDER_SI_WINDUP 10        /* sub-steps taken with an anti-windup clamp active */
#define PVDER_SI_EXACT 11         /* sub-steps redone with library sin/cos/exp (outside the incremental range) */
#define PVDER_SI_REDO_LIST 12     /* PVDER_3PH_AUTO scratch: local indices of the envs the balanced kernel handed to the three-lane one */
#define PVDER_SI_REDO_CTRL 13     /* PVDER_3PH_AUTO scratch: [0] number of list entries, [1] ticket of the consuming kernel (both 0 between steps) */
#define PVDER_SI_FIELDS 14

enum { PVDER_GOAL_VOLTAGE = 0, PVDER_GOAL_Q = 1, PVDER_GOAL_POWER = 2 };      /* PVDER_env.py:78-93 */
/* reward terms (PVDER_env.py:67-71 reward_list 'valid'; evaluated at :251-299) */
enum { PVDER_TERM_VOLTAGE = 0, PVDER_TERM_Q = 1, PVDER_TERM_POWER = 2, PVDER_TERM_VDC = 3 };
#define PVDER_MAX_REWARD_TERMS 4
enum { PVDER_EVENTS_NONE = 0, PVDER_EVENTS_PHILOX = 1, PVDER_EVENTS_TABLE = 2 };
enum { PVDER_STATUS_OK = 0, PVDER_STATUS_BAD_ACTION = 1, PVDER_STATUS_NONFINITE = 2,
       PVDER_STATUS_UNBALANCED = 3 /* PVDER_3PH_BALANCED only: the per-phase duty-cycle clamp (A.3) engaged, which the
                                      phase-a reduction cannot represent: reward -100, done (PVDER_3PH_AUTO hands such an
                                      env to the general model instead) */ };
/* three-phase integration mode (pvder_env_config.balanced3) */
enum { PVDER_3PH_GENERAL = 0,   /* 23 states, one thread per env */
       PVDER_3PH_BALANCED = 1,  /* balanced set carried by phase a (11 states) */
       PVDER_3PH_AUTO = 2,      /* per env: balanced reduction while the stored state is a balanced set and no per-phase
                                   duty-cycle clamp engages; any other env is stepped by the three-lane kernel */
       PVDER_3PH_SPLIT = 3 };   /* 23 states, three lanes per env (one per phase, warp-shuffle reductions) */
enum { PVDER_OK = 0, PVDER_ERR_INVALID = -1, PVDER_ERR_CUDA = -2, PVDER_ERR_NOMEM = -3 };

/* Per-unit DER parameters (values from config_der.json:2-21 / :66-84). */
typedef struct pvder_params {
  double Rf, Rt, Xt, inv_Lf, inv_wb;
  double Kp_GCC, Ki_GCC, Kp_DC, Ki_DC, Kp_Q, Ki_Q, wp, Kp_PLL, Ki_PLL;
  double inv_C, w0, dw;
  double vgs;        /* LV-side grid phasor magnitude at Vgrid = 1.0 pu */
  double np_iph100;  /* Np*(Iscr + Kv (T - T0)): array photo-current at Sinsol = 100 */
  double np_irs;     /* Np*Irs */
  double kappa;      /* q Vdcbase / (k T A Ns) */
  double pv_scale;   /* Vdcbase / Sbase */
  double Vrms_ref;   /* PV_model.Vrms_ref (PVDER_env.py:280) */
  double iref_limit; /* anti-windup current limit (A.3) */
  double m_limit10;  /* 10 * m_limit (A.3) */
  double p_target;   /* P_ref / Sbase (PVDER_env.py:69, :245) */
  double q_target;   /* Q_ref / Sbase (PVDER_env.py:69, :241) */
  double Lf, Rf_Rt;  /* reserved for the steady-state solve */
} pvder_params;

typedef struct pvder_env_config {
  pvder_params par;
  int32_t phases;            /* 1: model_1 / derId 10, 3: model_2 / derId 50 (PVDER_env.py:56-58) */
  int32_t n_sub_per_step;    /* half-cycle sub-steps per env step = 2 * n_sim_time_steps_per_env_step */
  int32_t base_level;        /* fine-step level of every sub-step: 2^base_level integrator steps per half-cycle (0) */
  int32_t done_substep;      /* done when k >= done_substep  (tStop >= max_sim_time, PVDER_env.py:183) */
  int32_t discrete_reward;   /* DISCRETE_REWARD (PVDER_env.py:590-600) */
  int32_t goal;              /* PVDER_GOAL_* = goals_list[0] (PVDER_env.py:234) */
  int32_t auto_reset;        /* vector-env option: reset in place when done */
  int32_t event_mode;        /* PVDER_EVENTS_* */
  int32_t ev_start_k, ev_step_k, ev_count;   /* event instants on the 1/120 s grid (PVDER_env.py:60-61) */
  int32_t ev_voltage_enable, ev_insol_enable;
  int32_t balanced3;         /* phases == 3: PVDER_3PH_*: general 23-state integration, balanced set on phase a
                                (b, c = rotated copies), or per-env auto-detection */
  /* Fine steps (level L = 2^L integrator steps of h/2^L instead of one of h = 1/120 s; 0 = off, max PVDER_FINE_LEVELS).
     The reference's LSODA (PVDER_env.py:166) adapts its step to input steps and to the PLL pull-in after
     reset; the fixed half-cycle grid refines exactly those sub-steps. */
  int32_t refine_input_level; /* the sub-step whose inputs changed at its start: an event instant, or (refine_on_action)
                                 an action that moved Q_ref / Vdc_ref */
  int32_t refine_on_action;
  int32_t startup_substeps;   /* sub-steps k < startup_substeps of every episode (PLL pull-in: wte0 = 6.28 is ~90 degrees */
  int32_t startup_level;      /* from lock, config_der.json:18) are taken at startup_level */
  int32_t reward_terms[PVDER_MAX_REWARD_TERMS]; /* PVDER_TERM_* in the order of env_goal_spec[goal]['reward']['my_spec']
                                 (PVDER_env.py:249), -1 terminated: the reward is their sum.  Default: the goal's
                                 required term (:451) */
  double ev_v_min, ev_v_max, ev_s_min, ev_s_max;
  double delQ_pu, delVdc_pu; /* per-step reference increments (PVDER_env.py:617-618, :225, :229) */
  double max_sim_time;       /* PVDER_env.py:561-575 */
  double substeps_per_sec;   /* 120 */
  uint64_t seed;
  double Q_ref0, Vdc_ref0;
  double y0[PVDER_MAX_STATES]; /* reset state (steady-state init, A.6), delta form */
  double vg_ratio_b, vg_ratio_c; /* grid magnitude of phases b, c relative to phase a (pvder
                                    Grid(unbalance_ratio_b/c); the env builds Grid(events=...) with 1.0,
                                    PVDER_env.py:372).  != 1 needs PVDER_3PH_GENERAL or PVDER_3PH_SPLIT */
} pvder_env_config;

int pvder_abi_version(void);
const char* pvder_error_string(int code);
size_t pvder_sd_fields(int phases);
size_t pvder_si_fields(void);
size_t pvder_config_size(void);   /* sizeof(pvder_env_config): lets a binding verify its struct mirror */

/* Steady-state initialisation (replaces DERModel(..., steadyStateInitialization=True),
 * PVDER_env.py:374-378).  Host-only Newton solve; y0 gets ns doubles. */
int pvder_steady_state(const pvder_params* par, int phases, double Vdc, double Vgrid, double Sinsol,
                       double Q_ref, double wte0, double* y0, double* ma0, double* ia0);

/* reset(): PVDER_env.py:316-334 + 366-398 + 400-411.  mask (device, nullable): reset only envs
 * with mask[i] != 0.  init != 0: first reset after allocation (episode := 0). */
int pvder_reset(const pvder_env_config* cfg, double* sd, int32_t* si, int64_t ld, const uint8_t* mask,
                int32_t init, float* obs_f32, double* obs_f64, int64_t n_envs, int64_t env_offset,
                void* stream);

/* step(action): PVDER_env.py:138-196 (action_calc 198-229, run_simulation 166, reward_calc
 * 231-301, state 531-542, done 183) for n_envs environments in ONE launch.
 * action: device int32[n_envs]; vgrid_tab/sinsol_tab: device double[ev_count][ld] when
 * event_mode == PVDER_EVENTS_TABLE (value in force from event instant j on), else NULL.
 * Outputs (device, each nullable): obs_f32[n_envs][11], obs_f64[n_envs][11], reward_f64[n_envs],
 * reward_i32[n_envs] (discrete mode), done[n_envs]. */
int pvder_step(const pvder_env_config* cfg, double* sd, int32_t* si, int64_t ld, const int32_t* action,
               const double* vgrid_tab, const double* sinsol_tab, float* obs_f32, double* obs_f64,
               double* reward_f64, int32_t* reward_i32, uint8_t* done, int64_t n_envs, int64_t env_offset,
               void* stream);

/* step() that also records the trajectory inside the env step (what the reference's SimulationResults
 * plots after run_simulation(), PVDER_env.py:358-364): for the envs e = j * traj_stride, j < traj_envs,
 * traj[s][r][j] (device double[n_sub_per_step][6*phases + 5 + 2][traj_envs]) receives the state (stored
 * layout, PLL angle as delta) after half-cycle sub-step s, then Vgrid and Sinsol in force during it. */
int pvder_step_record(const pvder_env_config* cfg, double* sd, int32_t* si, int64_t ld, const int32_t* action,
                      const double* vgrid_tab, const double* sinsol_tab, float* obs_f32, double* obs_f64,
                      double* reward_f64, int32_t* reward_i32, uint8_t* done, int64_t n_envs, int64_t env_offset,
                      double* traj, int64_t traj_envs, int64_t traj_stride, void* stream);

/* Event generator (replaces SimulationEvents.create_random_events as called at
 * PVDER_env.py:408-411): materialises the per-env tables the step kernel draws on the fly in
 * PVDER_EVENTS_PHILOX mode.  episode: device int32[n_envs] or NULL (= 0). */
int pvder_generate_events(const pvder_env_config* cfg, const int32_t* episode, double* vgrid_tab,
                          double* sinsol_tab, int64_t ld, int64_t n_envs, int64_t env_offset, void* stream);

/* action_space.sample() for every env (examples/gym_PVDER_environment_import_test.py:23). */
int pvder_sample_actions(uint64_t seed, int64_t step_index, int32_t* action, int64_t n_envs,
                         int64_t env_offset, void* stream);

/* Policy in the loop (the collect step of examples/gym_PVDER_environment_tf_agents_DQN_demo.ipynb:
 * QNetwork(fc_layer_params=(hidden,)) + epsilon-greedy collect policy): obs_f32[n][11] ->
 * Q = W2 relu(W1 obs + b1) + b2 -> argmax, or with probability epsilon a uniform action (Philox stream 2
 * keyed by global env index and step) -> action[n].  step = step_index + *step_index_dev (device counter,
 * nullable: lets a captured CUDA graph draw fresh numbers on every replay).  w1[hidden][11], b1[hidden],
 * w2[5][hidden], b2[5]: device float32, row-major (torch Linear layout); hidden <= 256.  q_out (nullable):
 * Q values [n][5]. */
int pvder_qnet_policy(const float* obs_f32, const float* w1, const float* b1, const float* w2, const float* b2,
                      int hidden, float epsilon, uint64_t seed, int64_t step_index, const int64_t* step_index_dev,
                      int32_t* action, float* q_out, int64_t n_envs, int64_t env_offset, void* stream);

/* Collect step of the same demo with the replay-buffer writer fused in (SoA ring [ring_slots][n][..]): with
 * t = *transitions_dev (device counter of completed transitions; the caller increments it after every env
 * step), the transition the preceding env step completed (finish != 0) is closed in slot (t-1) mod ring_slots
 * -- rb_next = obs, rb_rew = reward, rb_done = done -- and the new obs plus the action chosen for it open slot
 * t mod ring_slots (rb_obs, rb_act).  The exploration stream is keyed by (global env, t). */
int pvder_qnet_collect(const float* obs_f32, const float* w1, const float* b1, const float* w2, const float* b2,
                       int hidden, float epsilon, uint64_t seed, const int64_t* transitions_dev, int32_t* action,
                       float* rb_obs, float* rb_next, int32_t* rb_act, float* rb_rew, uint8_t* rb_done,
                       int64_t ring_slots, const double* reward_f64, const int32_t* reward_i32, const uint8_t* done,
                       int32_t finish, int64_t n_envs, int64_t env_offset, void* stream);

/* Episode statistics (env_utilities.py:12-46) reduced over envs into 16 device doubles:
 * [0] sum return, [1] sum steps, [2] n done, [3] n failed, [4..8] action histogram,
 * [9] windup sub-steps, [10] n_envs, [11] sub-steps redone with library transcendentals. */
int pvder_stats_reduce(const double* sd, const int32_t* si, int64_t ld, int phases, int64_t n_envs,
                       double* out16, void* stream);

/* FP64 FMA peak micro-benchmark (roofline denominator; synchronises). */
int pvder_fp64_peak(int iters, double* tflops, double* ms);

/* ---- host-buffer handle API: the call a Gym user makes, numpy in / numpy out ---- */
typedef struct pvder_env pvder_env;
int pvder_env_create(const pvder_env_config* cfg, int64_t n_envs, int64_t env_offset, pvder_env** out);
int pvder_env_destroy(pvder_env* env);
/* Swap the configuration of a handle (seed, goal, reward terms, event ranges, n_sim ...) keeping its streams and device
   buffers: what PVDER.reset() -- which builds a new simulator per episode, PVDER_env.py:316-334, :366-398 -- does here
   instead of re-creating the handle.  phases / three-phase mode / event-table shape must not change. */
int pvder_env_reconfigure(pvder_env* env, const pvder_env_config* cfg);
int pvder_env_set_event_tables(pvder_env* env, const double* vgrid_tab, const double* sinsol_tab);
int pvder_env_reset_host(pvder_env* env, float* obs_out, double* obs64_out);
/* Copies action host->device, launches pvder_step, copies obs/reward/done device->host, waits. */
int pvder_env_step_host(pvder_env* env, const int32_t* action, float* obs_out, double* obs64_out,
                        double* reward_out, uint8_t* done_out);
/* Same step with COMPACT result formats (opt-in, for consumers bound by the host's copy bandwidth: several ranks per box):
 * observations as IEEE half [n][11] (the Box is [-10, 10]: 1e-3 relative), reward as float32, done as one bit per env
 * (env 32 w + b in bit b of done_bits[w]).  Each output is nullable ("obs on demand").  53 -> 26.1 bytes per env step. */
int pvder_env_step_host_compact(pvder_env* env, const int32_t* action, uint16_t* obs_f16_out, float* reward_f32_out,
                                uint32_t* done_bits_out);
int pvder_env_state_host(pvder_env* env, double* sd_out, int32_t* si_out);
int pvder_env_set_refs_host(pvder_env* env, const double* sd_in);
/* Device pointers of the handle's state (for zero-copy interop). */
int pvder_env_device_ptrs(pvder_env* env, double** sd, int32_t** si, int64_t* ld);
/* Average device time (ms) of the step kernel launches since the last call (CUDA events). */
int pvder_env_kernel_ms(pvder_env* env, double* ms_total, int64_t* launches);
/* Host-buffer pipeline of pvder_env_step_host: chunks of the last call and the running estimate of
   (device->host copy time) / (kernel time) per env that sizes them. */
int pvder_env_pipeline_info(pvder_env* env, int32_t* chunks, double* copy_ratio);
/* The chunk plan pvder_env_step_host uses for a batch of `units` quarter waves of resident CTAs and a copy/kernel
   time ratio q: writes up to 12 chunk sizes (in units), returns their number.  Pure host arithmetic. */
int pvder_plan_chunks(int64_t units, double q, int64_t* sizes);
void* pvder_host_alloc(size_t bytes);   /* pinned host memory */
void pvder_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* PVDER_B200_H_ */
