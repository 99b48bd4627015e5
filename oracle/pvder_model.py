"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Restatement of the dynamic-phasor PV-DER model that the reference environment
integrates through the third-party ``pvder`` package (imported at reference
gym_PVDER/envs/PVDER_env.py:26-35; constructed at :371-391).  ``pvder`` is NOT
vendored in /root/reference, is not installed in this image and is not pinned
to a version by the reference (setup.py:5 does not even list it), so the model
below follows the published pvder equations as written down in SURVEY.md
Appendix A, anchored on the constants the reference itself ships
(config_der.json:2-21 derId 50, :66-84 derId 10; PVDER_env.py:56-75).

PARITY UNPINNED: the reference's tests hold no numeric known-answer for this
path (gym_PVDER/tests/test_gym_PVDER.py checks only types/shapes/step counts),
and the reference cannot run here.  What *is* pinned (tests/test_oracle_*.py):
per-unit constants and steady-state operating points recorded in SURVEY.md
Appendix B, P_ref = 45.4 kW (PVDER_env.py:69), Vrms_ref = 177/500.

State order (names as config_der.json:17-18):
  per phase p: iR, iI, xR, xI, uR, uI ; then Vdc, xDC, xQ, xPLL, wte
  single-phase n = 11, three-phase n = 23.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# PV module constants
ISCR, KV, T0, IRS, QE, KB, ADIODE = 8.03, 0.0017, 298.15, 1.2e-7, 1.602e-19, 1.38e-23, 1.92

TWO_PI_3 = 2.0 * math.pi / 3.0


@dataclass
class DERParams:
    """Per-unit parameters of one DER (SURVEY.md A.0, Appendix B known answers)."""

    phases: int
    # bases
    Vbase: float = 500.0
    Sbase: float = 50e3
    wbase: float = 2.0 * math.pi * 60.0
    # per-unit circuit
    Rf: float = 0.0
    Lf: float = 0.0
    C: float = 0.0
    Rt: float = 0.0  # Re(Z1 + Z2/a^2)
    Xt: float = 0.0  # Im(Z1 + Z2/a^2)
    vgs: float = 0.0  # grid phasor magnitude referred to LV side at Vgrid = 1.0
    # gains
    Kp_GCC: float = 0.0
    Ki_GCC: float = 0.0
    Kp_DC: float = 0.0
    Ki_DC: float = 0.0
    Kp_Q: float = 0.0
    Ki_Q: float = 0.0
    wp: float = 0.0
    Kp_PLL: float = 180.0
    Ki_PLL: float = 320.0
    # PV module
    Np: float = 0.0
    Ns: float = 0.0
    Tactual: float = 298.15
    # references / limits
    Vdc_ref0: float = 0.0
    Vrms_ref: float = 0.0
    iref_limit: float = 0.0
    m_limit: float = 1.0
    wte0: float = 6.28
    ss_guess: tuple = (0.0, 0.0, 0.0, 0.0)
    # grid unbalance (pvder Grid(unbalance_ratio_b, unbalance_ratio_c); the env builds
    # Grid(events=...) with the defaults 1.0, reference PVDER_env.py:372): per-phase magnitude factors
    vg_ratio: tuple = (1.0, 1.0, 1.0)
    # three-phase PLL input (A.4): "abc_dq0" = pvder's abc->dq0 of the time-domain voltages (carries a
    # 2w ripple on unbalanced sets), "posseq" = its cycle average, the positive-sequence projection the
    # half-cycle kernel integrates (identical on balanced sets)
    pll_mode: str = "abc_dq0"
    # derived extras kept for known-answer tests
    extras: dict = field(default_factory=dict)

    @property
    def n_state(self) -> int:
        return 6 * self.phases + 5

    @property
    def Vdcbase(self) -> float:
        return self.Vbase

    @property
    def Ibase(self) -> float:
        return self.Sbase / self.Vbase

    @property
    def kappa(self) -> float:
        """exp() argument per pu of Vdc in the diode term."""
        return QE * self.Vdcbase / (KB * self.Tactual * ADIODE * self.Ns)


def load_der_params(der_id: str, config_file: str | None = None) -> DERParams:
    """Per-unit conversion of one derId (values of reference
    config_der.json:2-21 / :66-84 carried in this repo's own der_config.json)."""
    config_file = config_file or os.path.join(_HERE, "der_config.json")
    with open(config_file) as fh:
        raw = json.load(fh)[str(der_id)]
    Vbase, Sbase = 500.0, 50e3
    wbase = 2.0 * math.pi * 60.0
    Zbase = Vbase * Vbase / Sbase
    Lbase = Zbase / wbase
    Cbase = 1.0 / (Zbase * wbase)
    Vgridrated = 20415.0
    Z2 = complex(1.61, 5.54) / Zbase
    a = Vgridrated / (raw["Vrmsrated"] * math.sqrt(2.0))
    Z1 = complex(raw["R1_actual"], raw["X1_actual"]) / Zbase
    Zt = Z1 + Z2 / (a * a)
    vag = Vgridrated / Vbase
    phases = int(raw["phases"])
    Varated = raw["Vrmsrated"] * math.sqrt(2.0)
    # rated peak phase current (iref_limit = Ioverload*Irated/Ibase)
    Irated = (raw["Srated"] / (phases * (Varated / math.sqrt(2.0)))) * math.sqrt(2.0)
    p = DERParams(
        phases=phases, Vbase=Vbase, Sbase=Sbase, wbase=wbase,
        Rf=raw["Rf_actual"] / Zbase, Lf=raw["Lf_actual"] / Lbase, C=raw["C_actual"] / Cbase,
        Rt=Zt.real, Xt=Zt.imag, vgs=vag / a,
        Kp_GCC=raw["Kp_GCC"], Ki_GCC=raw["Ki_GCC"], Kp_DC=raw["Kp_DC"], Ki_DC=raw["Ki_DC"],
        Kp_Q=raw["Kp_Q"], Ki_Q=raw["Ki_Q"], wp=raw["wp"],
        Np=float(raw["Np"]), Ns=float(raw["Ns"]),
        Vdc_ref0=raw["Vdcmpp0"] / Vbase, Vrms_ref=raw["Vrmsrated"] / Vbase,
        iref_limit=raw["Ioverload"] * Irated / (Sbase / Vbase),
        wte0=raw["wte0"],
        ss_guess=(raw["maR0"], raw["maI0"], raw["iaR0"], raw["iaI0"]),
    )
    p.extras = dict(Zbase=Zbase, Lbase=Lbase, Cbase=Cbase, Z2=Z2, Z1=Z1, a=a, vag=vag, Irated=Irated)
    return p


def phase_rot(phases: int):
    """Phasor rotation of phase p (a, b, c) and the abc->dq0 angle offsets (A.0, A.4)."""
    if phases == 1:
        return [complex(1.0, 0.0)], [0.0]
    rot = [complex(1.0, 0.0),
           complex(math.cos(-TWO_PI_3), math.sin(-TWO_PI_3)),
           complex(math.cos(TWO_PI_3), math.sin(TWO_PI_3))]
    alpha = [0.0, TWO_PI_3, -TWO_PI_3]
    return rot, alpha


def ppv_and_slope(p: DERParams, Vdc: float, Sinsol: float):
    """PV array power (pu) and dPpv/dVdc."""
    Iph = (ISCR + KV * (p.Tactual - T0)) * (Sinsol / 100.0)
    e = math.exp(p.kappa * Vdc)
    Ipv = p.Np * Iph - p.Np * IRS * (e - 1.0)
    scale = p.Vdcbase / p.Sbase
    P = Ipv * Vdc * scale
    if P <= 0.0:
        return 0.0, 0.0
    dIpv = -p.Np * IRS * p.kappa * e
    return P, scale * (Ipv + Vdc * dIpv)


@dataclass
class Inputs:
    """Exogenous quantities held constant over one integration piece."""

    Vgrid: float = 1.0     # grid voltage magnitude event value (pu of rated), A.8
    Sinsol: float = 100.0  # insolation event value
    Q_ref: float = 0.0
    Vdc_ref: float = 0.0
    wgrid: float = 2.0 * math.pi * 60.0
    # anti-windup clamp (A.3): None = evaluate inside the RHS from the current state (pvder
    # behaviour); a tuple of bools = hold this per-row freeze mask (sampled at a half-cycle
    # boundary; order: per phase xR,xI,uR,uI ; then xDC, xQ)
    freeze: tuple | None = None


class PVDERModel:
    """ODE right-hand side, analytic Jacobian and algebraic outputs (A.2-A.5)."""

    def __init__(self, params: DERParams):
        self.p = params
        self.rot, self.alpha = phase_rot(params.phases)
        self.n = params.n_state
        self.windup_hits = 0

    # ---- algebraic helpers -------------------------------------------------
    def _vd_and_grads(self, vR, vI, t, wte, wgrid):
        """d-axis PCC voltage seen by the PLL and its partials (A.4).
        Returns vd, [d vd/d iR_p], [d vd/d iI_p], d vd/d wte."""
        p = self.p
        if p.phases == 1:
            psi = wgrid * t - wte
            c, s = math.cos(psi), math.sin(psi)
            vd = vR[0] * c - vI[0] * s
            return vd, [p.Rt * c - p.Xt * s], [-p.Xt * c - p.Rt * s], vR[0] * s + vI[0] * c
        if p.pll_mode == "posseq":
            dl = wte - wgrid * t
            vd, dR, dI, dw = 0.0, [], [], 0.0
            for k in range(3):
                ck = math.cos(dl - self.alpha[k])
                sk = math.sin(dl - self.alpha[k])
                vd += (vR[k] * ck + vI[k] * sk) / 3.0
                dR.append((p.Rt * ck + p.Xt * sk) / 3.0)
                dI.append((-p.Xt * ck + p.Rt * sk) / 3.0)
                dw += (-vR[k] * sk + vI[k] * ck) / 3.0
            return vd, dR, dI, dw
        cwt, swt = math.cos(wgrid * t), math.sin(wgrid * t)
        vd = 0.0
        dR, dI = [], []
        dw = 0.0
        for k in range(3):
            ck = math.cos(wte - self.alpha[k])
            sk = math.sin(wte - self.alpha[k])
            vt_k = vR[k] * cwt - vI[k] * swt
            vd += (2.0 / 3.0) * vt_k * ck
            dR.append((2.0 / 3.0) * (p.Rt * cwt - p.Xt * swt) * ck)
            dI.append((2.0 / 3.0) * (-p.Xt * cwt - p.Rt * swt) * ck)
            dw += -(2.0 / 3.0) * vt_k * sk
        return vd, dR, dI, dw

    def freeze_mask(self, y, inp: Inputs):
        """Per-row clamp decisions at state y: when |m_p| > 10*m_limit the x/u
        integrators of every phase, and when |i_ref| > iref_limit the xDC/xQ integrators, are
        frozen if their derivative has the sign of the state (np.sign(d) == np.sign(x))."""
        p = self.p
        P_ = p.phases
        base = 6 * P_
        m_over = False
        for k in range(P_):
            o = 6 * k
            mR = p.Kp_GCC * y[o + 4] + y[o + 2]
            mI = p.Kp_GCC * y[o + 5] + y[o + 3]
            if mR * mR + mI * mI > (10.0 * p.m_limit) ** 2:
                m_over = True
        Q = self.q_pcc(y, inp)
        Vdc, xDC, xQ = y[base], y[base + 1], y[base + 2]
        irefR = xDC + p.Kp_DC * (inp.Vdc_ref - Vdc)
        irefI = xQ - p.Kp_Q * (inp.Q_ref - Q)
        i_over = irefR * irefR + irefI * irefI > p.iref_limit ** 2
        mask = [False] * (4 * P_ + 2)
        if m_over:
            for k in range(P_):
                o = 6 * k
                r = self.rot[k]
                duR = p.wp * (-y[o + 4] + (r.real * irefR - r.imag * irefI) - y[o])
                duI = p.wp * (-y[o + 5] + (r.imag * irefR + r.real * irefI) - y[o + 1])
                mask[4 * k] = _same_sign(p.Ki_GCC * y[o + 4], y[o + 2])
                mask[4 * k + 1] = _same_sign(p.Ki_GCC * y[o + 5], y[o + 3])
                mask[4 * k + 2] = _same_sign(duR, y[o + 4])
                mask[4 * k + 3] = _same_sign(duI, y[o + 5])
        if i_over:
            mask[4 * P_] = _same_sign(p.Ki_DC * (inp.Vdc_ref - Vdc), xDC)
            mask[4 * P_ + 1] = _same_sign(-p.Ki_Q * (inp.Q_ref - Q), xQ)
        return tuple(mask)

    def q_pcc(self, y, inp: Inputs):
        p = self.p
        Q = 0.0
        for k in range(p.phases):
            o = 6 * k
            iR, iI = y[o], y[o + 1]
            vg = inp.Vgrid * p.vgs * p.vg_ratio[k] * self.rot[k]
            Q += 0.5 * (vg.imag * iR - vg.real * iI + p.Xt * (iR * iR + iI * iI))
        return Q

    # ---- RHS ----------------------------------------------------------------
    def rhs(self, y, t, inp: Inputs):
        p = self.p
        P_ = p.phases
        base = 6 * P_
        Vdc, xDC, xQ, xPLL, wte = y[base], y[base + 1], y[base + 2], y[base + 3], y[base + 4]
        vR, vI = [0.0] * P_, [0.0] * P_
        Q = 0.0
        Pinv = 0.0
        for k in range(P_):
            o = 6 * k
            iR, iI = y[o], y[o + 1]
            vg = inp.Vgrid * p.vgs * p.vg_ratio[k] * self.rot[k]
            vR[k] = vg.real + p.Rt * iR - p.Xt * iI
            vI[k] = vg.imag + p.Xt * iR + p.Rt * iI
            mR = p.Kp_GCC * y[o + 4] + y[o + 2]
            mI = p.Kp_GCC * y[o + 5] + y[o + 3]
            Q += 0.5 * (vI[k] * iR - vR[k] * iI)
            Pinv += 0.25 * Vdc * (mR * iR + mI * iI)
        vd, _, _, _ = self._vd_and_grads(vR, vI, t, wte, inp.wgrid)
        we = p.Kp_PLL * vd + xPLL + 2.0 * math.pi * 60.0
        irefR = xDC + p.Kp_DC * (inp.Vdc_ref - Vdc)
        irefI = xQ - p.Kp_Q * (inp.Q_ref - Q)
        frz = self.freeze_mask(y, inp) if inp.freeze is None else inp.freeze
        if any(frz):
            self.windup_hits += 1
        dy = [0.0] * self.n
        inv_Lf = 1.0 / p.Lf
        wr = we / p.wbase
        for k in range(P_):
            o = 6 * k
            iR, iI = y[o], y[o + 1]
            mR = p.Kp_GCC * y[o + 4] + y[o + 2]
            mI = p.Kp_GCC * y[o + 5] + y[o + 3]
            dy[o] = inv_Lf * (-p.Rf * iR - vR[k] + 0.5 * mR * Vdc) + wr * iI
            dy[o + 1] = inv_Lf * (-p.Rf * iI - vI[k] + 0.5 * mI * Vdc) - wr * iR
            r = self.rot[k]
            rrefR = r.real * irefR - r.imag * irefI
            rrefI = r.imag * irefR + r.real * irefI
            dxR = p.Ki_GCC * y[o + 4]
            dxI = p.Ki_GCC * y[o + 5]
            duR = p.wp * (-y[o + 4] + rrefR - iR)
            duI = p.wp * (-y[o + 5] + rrefI - iI)
            if frz[4 * k]:
                dxR = 0.0
            if frz[4 * k + 1]:
                dxI = 0.0
            if frz[4 * k + 2]:
                duR = 0.0
            if frz[4 * k + 3]:
                duI = 0.0
            dy[o + 2], dy[o + 3], dy[o + 4], dy[o + 5] = dxR, dxI, duR, duI
        Ppv, _ = ppv_and_slope(p, Vdc, inp.Sinsol)
        dy[base] = (Ppv - Pinv) / (Vdc * p.C)
        dxDC = p.Ki_DC * (inp.Vdc_ref - Vdc)
        dxQ = -p.Ki_Q * (inp.Q_ref - Q)
        if frz[4 * P_]:
            dxDC = 0.0
        if frz[4 * P_ + 1]:
            dxQ = 0.0
        dy[base + 1] = dxDC
        dy[base + 2] = dxQ
        dy[base + 3] = p.Ki_PLL * vd
        dy[base + 4] = we
        return dy

    # ---- Jacobian -------------------------------------------------------------
    def jac(self, y, t, inp: Inputs):
        p = self.p
        P_ = p.phases
        n = self.n
        base = 6 * P_
        Vdc, xDC, xQ, xPLL, wte = y[base], y[base + 1], y[base + 2], y[base + 3], y[base + 4]
        J = np.zeros((n, n))
        vR, vI = [0.0] * P_, [0.0] * P_
        dQ_R, dQ_I = [0.0] * P_, [0.0] * P_
        Pinv = 0.0
        Q = 0.0
        for k in range(P_):
            o = 6 * k
            iR, iI = y[o], y[o + 1]
            vg = inp.Vgrid * p.vgs * p.vg_ratio[k] * self.rot[k]
            vR[k] = vg.real + p.Rt * iR - p.Xt * iI
            vI[k] = vg.imag + p.Xt * iR + p.Rt * iI
            dQ_R[k] = 0.5 * (vg.imag + 2.0 * p.Xt * iR)
            dQ_I[k] = 0.5 * (-vg.real + 2.0 * p.Xt * iI)
            mR = p.Kp_GCC * y[o + 4] + y[o + 2]
            mI = p.Kp_GCC * y[o + 5] + y[o + 3]
            Pinv += 0.25 * Vdc * (mR * iR + mI * iI)
            Q += 0.5 * (vI[k] * iR - vR[k] * iI)
        vd, dvd_R, dvd_I, dvd_w = self._vd_and_grads(vR, vI, t, wte, inp.wgrid)
        we = p.Kp_PLL * vd + xPLL + 2.0 * math.pi * 60.0
        inv_Lf = 1.0 / p.Lf
        wr = we / p.wbase
        kw = p.Kp_PLL / p.wbase
        frz = self.freeze_mask(y, inp) if inp.freeze is None else inp.freeze
        for k in range(P_):
            o = 6 * k
            iR, iI = y[o], y[o + 1]
            mR = p.Kp_GCC * y[o + 4] + y[o + 2]
            mI = p.Kp_GCC * y[o + 5] + y[o + 3]
            # current rows
            J[o, o] += inv_Lf * (-p.Rf - p.Rt)
            J[o, o + 1] += inv_Lf * p.Xt + wr
            J[o + 1, o] += -inv_Lf * p.Xt - wr
            J[o + 1, o + 1] += inv_Lf * (-p.Rf - p.Rt)
            for q in range(P_):
                oq = 6 * q
                J[o, oq] += iI * kw * dvd_R[q]
                J[o, oq + 1] += iI * kw * dvd_I[q]
                J[o + 1, oq] += -iR * kw * dvd_R[q]
                J[o + 1, oq + 1] += -iR * kw * dvd_I[q]
            J[o, o + 2] = 0.5 * Vdc * inv_Lf
            J[o + 1, o + 3] = 0.5 * Vdc * inv_Lf
            J[o, o + 4] = 0.5 * p.Kp_GCC * Vdc * inv_Lf
            J[o + 1, o + 5] = 0.5 * p.Kp_GCC * Vdc * inv_Lf
            J[o, base] = 0.5 * mR * inv_Lf
            J[o + 1, base] = 0.5 * mI * inv_Lf
            J[o, base + 3] = iI / p.wbase
            J[o + 1, base + 3] = -iR / p.wbase
            J[o, base + 4] = iI * kw * dvd_w
            J[o + 1, base + 4] = -iR * kw * dvd_w
            # x rows
            J[o + 2, o + 4] = p.Ki_GCC
            J[o + 3, o + 5] = p.Ki_GCC
            # u rows
            r = self.rot[k]
            J[o + 4, o + 4] = -p.wp
            J[o + 5, o + 5] = -p.wp
            J[o + 4, o] += -p.wp
            J[o + 5, o + 1] += -p.wp
            J[o + 4, base + 1] = p.wp * r.real
            J[o + 5, base + 1] = p.wp * r.imag
            J[o + 4, base] = -p.wp * r.real * p.Kp_DC
            J[o + 5, base] = -p.wp * r.imag * p.Kp_DC
            J[o + 4, base + 2] = -p.wp * r.imag
            J[o + 5, base + 2] = p.wp * r.real
            for q in range(P_):
                oq = 6 * q
                J[o + 4, oq] += -p.wp * r.imag * p.Kp_Q * dQ_R[q]
                J[o + 4, oq + 1] += -p.wp * r.imag * p.Kp_Q * dQ_I[q]
                J[o + 5, oq] += p.wp * r.real * p.Kp_Q * dQ_R[q]
                J[o + 5, oq + 1] += p.wp * r.real * p.Kp_Q * dQ_I[q]
            # DC link row
            inv_VC = 1.0 / (Vdc * p.C)
            J[base, o] = -0.25 * Vdc * mR * inv_VC
            J[base, o + 1] = -0.25 * Vdc * mI * inv_VC
            J[base, o + 2] = -0.25 * Vdc * iR * inv_VC
            J[base, o + 3] = -0.25 * Vdc * iI * inv_VC
            J[base, o + 4] = -0.25 * Vdc * p.Kp_GCC * iR * inv_VC
            J[base, o + 5] = -0.25 * Vdc * p.Kp_GCC * iI * inv_VC
            # xQ, PLL rows
            J[base + 2, o] = p.Ki_Q * dQ_R[k]
            J[base + 2, o + 1] = p.Ki_Q * dQ_I[k]
            J[base + 3, o] = p.Ki_PLL * dvd_R[k]
            J[base + 3, o + 1] = p.Ki_PLL * dvd_I[k]
            J[base + 4, o] = p.Kp_PLL * dvd_R[k]
            J[base + 4, o + 1] = p.Kp_PLL * dvd_I[k]
        Ppv, dPpv = ppv_and_slope(p, Vdc, inp.Sinsol)
        inv_VC = 1.0 / (Vdc * p.C)
        J[base, base] = (dPpv - Pinv / Vdc) * inv_VC - (Ppv - Pinv) * inv_VC / Vdc
        J[base + 1, base] = -p.Ki_DC
        J[base + 3, base + 4] = p.Ki_PLL * dvd_w
        J[base + 4, base + 3] = 1.0
        J[base + 4, base + 4] = p.Kp_PLL * dvd_w
        for k in range(P_):
            for j in range(4):
                if frz[4 * k + j]:
                    J[6 * k + 2 + j, :] = 0.0
        if frz[4 * P_]:
            J[base + 1, :] = 0.0
        if frz[4 * P_ + 1]:
            J[base + 2, :] = 0.0
        return J

    # ---- outputs the env reads after a solver call (PVDER_env.py:535-540, 238-299) ----
    def outputs(self, y, inp: Inputs):
        """Algebraic functions of the state with the inputs in force (A.7 last para).
        Expression order here is the contract for the bit-exact device twin."""
        p = self.p
        base = 6 * p.phases
        iR, iI = y[0], y[1]
        vgR = inp.Vgrid * p.vgs
        vaR = vgR + (p.Rt * iR - p.Xt * iI)
        vaI = p.Xt * iR + p.Rt * iI
        Vdc = y[base]
        P = 0.0
        Q = 0.0
        v2 = 0.0
        for k in range(p.phases):
            o = 6 * k
            jR, jI = y[o], y[o + 1]
            vg = ((inp.Vgrid * p.vgs) * p.vg_ratio[k]) * self.rot[k]
            vkR = vg.real + (p.Rt * jR - p.Xt * jI)
            vkI = vg.imag + (p.Xt * jR + p.Rt * jI)
            P = P + 0.5 * (vkR * jR + vkI * jI)
            Q = Q + 0.5 * (vkI * jR - vkR * jI)
            v2 = v2 + (vkR * vkR + vkI * vkI)
        if p.phases == 1:
            Vrms = math.sqrt(v2) / math.sqrt(2.0)
        else:
            Vrms = math.sqrt(v2 / 3.0) / math.sqrt(2.0)
        Ppv, _ = ppv_and_slope(p, Vdc, inp.Sinsol)
        return dict(iaR=iR, iaI=iI, vaR=vaR, vaI=vaI, P_PCC=P, Q_PCC=Q, Vdc=Vdc, Ppv=Ppv,
                    Vrms=Vrms)

    # ---- steady state (A.6) ---------------------------------------------------
    def steady_state(self, Vgrid=1.0, Sinsol=100.0, Q_ref=0.0):
        """Newton solve of di/dt = 0 (we = wbase), Re S = Ppv, Im S_PCC = Q_ref at
        Vdc = Vdc_ref0 for phase a; other phases are rotated copies."""
        p = self.p
        Vdc = p.Vdc_ref0
        Ppv, _ = ppv_and_slope(p, Vdc, Sinsol)
        vg = Vgrid * p.vgs

        def resid(z):
            iR, iI = z
            i = complex(iR, iI)
            v = vg + complex(p.Rt, p.Xt) * i
            vt = v + p.Rf * i + 1j * p.Lf * i
            S = 0.5 * p.phases * vt * i.conjugate()
            Spcc = 0.5 * p.phases * v * i.conjugate()
            return np.array([S.real - Ppv, Spcc.imag - Q_ref]), vt

        z = np.array([p.ss_guess[2], p.ss_guess[3]], dtype=float)
        for _ in range(50):
            r, _ = resid(z)
            Jm = np.zeros((2, 2))
            for j in range(2):
                dz = np.zeros(2)
                dz[j] = 1e-7
                Jm[:, j] = (resid(z + dz)[0] - resid(z - dz)[0]) / 2e-7
            step = np.linalg.solve(Jm, r)
            z = z - step
            if np.max(np.abs(step)) < 1e-15:
                break
        _, vt = resid(z)
        ma0 = 2.0 * vt / Vdc
        ia0 = complex(z[0], z[1])
        y0 = np.zeros(self.n)
        for k in range(p.phases):
            o = 6 * k
            ik = ia0 * self.rot[k]
            mk = ma0 * self.rot[k]
            y0[o], y0[o + 1] = ik.real, ik.imag
            y0[o + 2], y0[o + 3] = mk.real, mk.imag
        base = 6 * p.phases
        y0[base] = Vdc
        y0[base + 1] = ia0.real
        y0[base + 2] = ia0.imag
        y0[base + 3] = 0.0
        y0[base + 4] = p.wte0
        return y0, ma0, ia0


def _same_sign(a: float, b: float) -> bool:
    """np.sign(a) == np.sign(b) (pvder's clamping test; zero equals only zero)."""
    sa = int(a > 0.0) - int(a < 0.0)
    sb = int(b > 0.0) - int(b < 0.0)
    return sa == sb
