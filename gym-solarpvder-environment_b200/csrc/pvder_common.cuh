// Shared device-side definitions for the PVDER-v0 step kernels (sm_100a).
//
// Replaces, per environment thread, what the reference obtains from the un-vendored `pvder`
// simulator objects built at reference gym_PVDER/envs/PVDER_env.py:371-391 (DER model, grid
// model, simulation events) -- equations in SURVEY.md Appendix A.
#pragma once
#include <cmath>
#include <cstdint>

#include "../../include/pvder_b200.h"

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define PVDER_DEV __device__ __forceinline__
#define PVDER_NOINLINE __device__ __noinline__
#define PVDER_HD __host__ __device__ __forceinline__
#else
// Plain C++ build of the per-env logic (tests/host_emul only): round-to-nearest intrinsics map
// to the plain IEEE operations (build with -ffp-contract=off).
#define PVDER_DEV inline
#define PVDER_NOINLINE static
#define PVDER_HD inline
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
using std::isfinite;
#endif

namespace pvder {

// Reciprocal of a well-scaled pivot (normal, positive, far from the ends of the exponent range): hardware
// seed (2^-23) + one third-order step (or two Newton steps) = full double accuracy without the IEEE division's special-case
// branches, which split the straight-line stepper code into scheduling regions.
#ifndef PVDER_RCP_CUBIC
#define PVDER_RCP_CUBIC 1
#endif
PVDER_DEV double pvder_rcp(double x) {
#ifdef __CUDACC__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#if PVDER_RCP_CUBIC
  // one third-order step: r (1 + e + e^2), e = 1 - x r ~ 2^-23 -> relative error e^3 ~ 2^-69 before the final rounding:
  // the accuracy of two Newton steps with three dependent FMAs instead of four
  const double e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);
#else
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
#endif
#else
  return 1.0 / x;
#endif
}

using Params = ::pvder_params;

// Exogenous inputs held constant over one half-cycle sub-step (events frozen per sub-step, A.8).
struct Inputs {
  double vg;       // LV-side grid phasor magnitude  = Vgrid_event * par.vgs  (phase a)
  double Qref;     // external reactive power reference (pu), changed only by actions
  double Vdcref;   // external DC-link reference (pu), changed only by actions
  double np_iph;   // Np * Iph(Sinsol)  (A)
  double vgb, vgc; // phase b / c magnitudes = vg * grid unbalance ratio (explicit three-phase models only)
  double pvA;      // (np_iph + Np Irs) * pv_scale: array current per pu of Vdc-power at E = 0 (see ppv_over_v_from_exp)
};

// Inputs in force for one sub-step.  Individually rounded ops: vg feeds the discrete reward.
PVDER_DEV Inputs make_inputs(const pvder_env_config& cfg, double Vgrid, double Qref, double Vdcref, double Sinsol) {
  Inputs in;
  in.vg = __dmul_rn(Vgrid, cfg.par.vgs);
  in.Qref = Qref;
  in.Vdcref = Vdcref;
  in.np_iph = __dmul_rn(cfg.par.np_iph100, __ddiv_rn(Sinsol, 100.0));
  in.vgb = __dmul_rn(in.vg, cfg.vg_ratio_b);
  in.vgc = __dmul_rn(in.vg, cfg.vg_ratio_c);
  in.pvA = (in.np_iph + cfg.par.np_irs) * cfg.par.pv_scale;
  return in;
}

PVDER_DEV double vg_of_phase(const Inputs& in, int phases, int k) {
  return (phases == 1 || k == 0) ? in.vg : ((k == 1) ? in.vgb : in.vgc);
}

// Transcendental side-inputs of the model at one state: sin/cos of the PLL angle delta and the PV array
// current per pu (which needs E = exp(kappa*Vdc)).  E is kept so the record can be advanced incrementally
// (pvder_env_step.cuh: aux_advance).
struct Aux {
  double sn, cs, E;
  double PoV;    // Ppv / Vdc: what the DC-link equation reads at every stage
  double dPoV;   // d(Ppv / Vdc) / dVdc: what its Jacobian reads (once per step)
};

// PV array: Ipv = np_iph - Np Irs (E - 1), Ppv = max(0, Ipv Vdc) pv_scale, written around the array
// current per pu
//   PoV = Ppv / Vdc = max(0, A - B E),  A = (np_iph + Np Irs) pv_scale (changes with the insolation events only),
//   B = Np Irs pv_scale:  one FMA per stage.  The DC-link equation (Ppv - P_inverter) / (C Vdc) =
//   (PoV - P_inverter / Vdc) / C reads PoV directly and P_inverter / Vdc = Ps / 4 has no Vdc in it, so the
//   equation's Vdc-derivative is dPoV / C with dPoV = -B kappa E: neither 1 / Vdc nor Ppv nor dPpv/dVdc is needed.
PVDER_DEV double ppv_over_v_from_exp(const Params& par, const Inputs& in, double e) {
  const double raw = fma(-(par.np_irs * par.pv_scale), e, in.pvA);
  return raw > 0.0 ? raw : 0.0;
}

PVDER_DEV void pov_from_exp(const Params& par, const Inputs& in, double e, double& PoV, double& dPoV) {
  const double B = par.np_irs * par.pv_scale;
  const double raw = fma(-B, e, in.pvA);
  const bool pos = raw > 0.0;
  PoV = pos ? raw : 0.0;
  dPoV = pos ? -(B * par.kappa) * e : 0.0;
}

PVDER_DEV void ppv_eval(const Params& par, const Inputs& in, double Vdc, double& P,
                                         double& dP) {
  const double e = exp(par.kappa * Vdc);
  const double Ipv = in.np_iph - par.np_irs * (e - 1.0);
  const double Pr = Ipv * Vdc * par.pv_scale;
  const double dPr = par.pv_scale * (Ipv - Vdc * (par.np_irs * par.kappa * e));
  const bool pos = Pr > 0.0;
  P = pos ? Pr : 0.0;
  dP = pos ? dPr : 0.0;
}

// ---- Philox4x32-10 (counter-based RNG; bit-exact numpy twin in oracle/philox_twin.py) ----------
PVDER_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0;
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform in [0,1) from two 32-bit words (same construction as numpy's random double).
PVDER_HD double u53(uint32_t hi, uint32_t lo) {
  const uint64_t a = hi >> 5, b = lo >> 6;
  return (double)(a * 67108864ull + b) * (1.0 / 9007199254740992.0);
}

enum : uint32_t { STREAM_EVENTS = 0u, STREAM_ACTIONS = 1u, STREAM_POLICY = 2u };

}  // namespace pvder
