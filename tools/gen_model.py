#!/usr/bin/env python
"""Code generator for the register-resident PV-DER model kernels.

Emits gym-solarpvder-environment_b200/csrc/pvder_model_{1ph,3ph}.cuh, each holding, for one
model (single-phase n=11, three-phase n=23; state order SURVEY.md A.1 with the PLL angle kept
as delta = wte - w_grid*t, A.4):

  * rhs():    the autonomous ODE right-hand side,
  * factor(): W = I/(h*gamma) - df/dy built from the analytic Jacobian (differentiated here with
              sympy) and LU-factorised *symbolically*: the sparsity pattern and the pivot order
              are fixed at generation time, so the factorisation is straight-line scalar code
              with no indexing, no pivot search and no zero work (41 of 121 / 165 of 529 entries
              are structurally non-zero) -- everything lives in registers,
  * solve():  the matching forward/backward substitution.

sin/cos(delta), Ppv(Vdc) and 1/Vdc enter through an ``Aux`` record the stepper maintains
incrementally (pvder_env_step.cuh), so the generated code is transcendental-free.

The anti-windup clamp (A.3) is sampled by the caller on the half-cycle grid and enters as per-row
effective gains (0 while clamped): a frozen row has f_r = 0 and W_r = e_r/(h*gamma) at no cost.

Run:  python tools/gen_model.py     (sympy needed only here, never at run time)
"""
import os
import sys

import sympy as sp

# Elimination order and the treatment of the stiff core (measured on B200, profiles/r2b_*_sweep.txt):
#   * tiers: Vdc (1.4) and the PLL angle (1.5) are pivoted BEFORE the current pair -- their two pivots are independent
#     of each other, so their reciprocals run in parallel (the round-1 order Vdc, iaR, iaI, dl had four sequential
#     reciprocal levels: 1.433 -> 1.409 ms per 1 Mi-env step);
# Study switches kept for the record (none of them is faster): CORE_INVERSE = 2 with CRAMER2 (the last 2x2 pivot block -- the
# current pair -- inverted by Cramer's rule: one reciprocal of the determinant instead of two sequential pivots, the block's
# four solve operations a 2x2 mat-vec of depth 2: 1.411 vs 1.419 ms in the 20-step window, 1.470 vs 1.461 over a full
# episode -- inside the box-to-box noise, so the plain LU stays), CORE_INVERSE = 1 | 3 (explicit inverse of the whole 4x4
# core / its last 3 pivots through the LU: more flops than the shorter chains buy back), CONST_PIVOTS = reg | bank,
# SCALED_U (U rows pre-multiplied by the reciprocal pivot: 1.430 vs 1.409 ms, the +9 multiplies cost more than the shorter
# back-substitution gains).
CORE_INVERSE = os.environ.get("PVDER_GEN_CORE_INVERSE", "0") != "0"   # "1": the whole core (<= 4 unknowns); "2"/"3": its last 2/3 pivots
CORE_N = int(os.environ.get("PVDER_GEN_CORE_INVERSE", "0"))
CRAMER2 = os.environ.get("PVDER_GEN_CRAMER2", "1") != "0" and CORE_N == 2
CONST_PIVOTS = os.environ.get("PVDER_GEN_CONST_PIVOTS", "off")   # off | reg | bank
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("PVDER_GEN_OUT") or os.path.join(ROOT, "gym-solarpvder-environment_b200", "csrc")
DL_TIER = float(os.environ.get("PVDER_GEN_DL_TIER", "1.5"))     # 2 = with the currents (round-1 order)
VDC_TIER = float(os.environ.get("PVDER_GEN_VDC_TIER", "1.4"))   # 2 = with the currents (round-1 order)
SCALED_U = os.environ.get("PVDER_GEN_SCALED_U", "0") != "0"

PAR = ["Rf", "Rt", "Xt", "inv_Lf", "inv_wb", "Kp_GCC", "Ki_GCC", "Kp_DC", "Ki_DC", "Kp_Q", "Ki_Q",
       "wp", "Kp_PLL", "Ki_PLL", "inv_C", "w0", "dw"]
INP = ["vg", "vgb", "vgc", "Qref", "Vdcref", "PoV", "dPoV"]


def build(P, mult=1):
    """P phases integrated explicitly; mult > 1: balanced three-phase set represented by phase a
    (phases b, c are rotated copies, so every sum over phases is mult * the phase-a term)."""
    par = {n: sp.Symbol("p_" + n) for n in PAR}
    inp = {n: sp.Symbol("in_" + n) for n in INP}
    names = []
    for k in range(P):
        ph = "abc"[k]
        names += [f"i{ph}R", f"i{ph}I", f"x{ph}R", f"x{ph}I", f"u{ph}R", f"u{ph}I"]
    names += ["Vdc", "xDC", "xQ", "xPLL", "dl"]
    y = [sp.Symbol("y_" + n) for n in names]
    n = len(y)
    base = 6 * P
    Vdc, xDC, xQ, xPLL, dl = y[base:base + 5]
    sn, cs = sp.Symbol("sn"), sp.Symbol("cs")          # sin(dl), cos(dl), supplied by sincos()
    if P == 1:
        rot = [(sp.Integer(1), sp.Integer(0))]
        alpha = [(sp.Integer(1), sp.Integer(0))]       # (cos a_k, sin a_k)
    else:
        h3 = sp.Symbol("SQ3") / 2     # sqrt(3), emitted as a literal
        rot = [(sp.Integer(1), sp.Integer(0)), (-sp.Rational(1, 2), -h3), (-sp.Rational(1, 2), h3)]
        alpha = [(sp.Integer(1), sp.Integer(0)), (-sp.Rational(1, 2), h3), (-sp.Rational(1, 2), -h3)]
    # grid phasor magnitude per phase: an explicit three-phase model takes vg, vgb, vgc (grid
    # unbalance ratios); the single-phase / balanced models use vg only
    vgk = [inp["vg"], inp["vgb"], inp["vgc"]] if (P == 3 and mult == 1) else [inp["vg"]] * P
    p = par
    # Freezable rows (anti-windup, A.3) are "gain x expression": their gains come in as per-row
    # effective values g_b (= the parameter, or 0 while the row is clamped), so clamping costs
    # nothing inside the RHS/Jacobian and a clamped row of W is automatically e_r/(h*gamma).
    nfrz = 4 * P + 2
    # gs[nfrz] is not a freezable row: it carries Ki_PLL pre-scaled by h*gamma (unit-pivot rows, see generate())
    # gs[nfrz + 1 + 2k + {0,1}]: reciprocal pivots of the u rows of phase k (1/(1/(h g) + wp), or h g while clamped)
    gs = [sp.Symbol(f"g_{b}") for b in range(nfrz + 1 + 2 * P)]
    vR, vI, mR, mI = [], [], [], []
    Q = 0
    Pinv = 0
    vd = 0
    for k in range(P):
        o = 6 * k
        iR, iI, xR, xI, uR, uI = y[o:o + 6]
        rr, ri = rot[k]
        vR.append(vgk[k] * rr + p["Rt"] * iR - p["Xt"] * iI)
        vI.append(vgk[k] * ri + p["Xt"] * iR + p["Rt"] * iI)
        mR.append(p["Kp_GCC"] * uR + xR)
        mI.append(p["Kp_GCC"] * uI + xI)
        Q += mult * sp.Rational(1, 2) * (vI[k] * iR - vR[k] * iI)
        Pinv += mult * sp.Rational(1, 4) * (mR[k] * iR + mI[k] * iI)          # P_inverter / Vdc
        ca, sa = alpha[k]
        ck = cs * ca + sn * sa       # cos(dl - a_k)
        sk = sn * ca - cs * sa       # sin(dl - a_k)
        vd += (vR[k] * ck + vI[k] * sk) / P
    we = p["Kp_PLL"] * vd + xPLL + p["w0"]
    wr = we * p["inv_wb"]
    irefR = xDC + p["Kp_DC"] * (inp["Vdcref"] - Vdc)
    irefI = xQ - p["Kp_Q"] * (inp["Qref"] - Q)
    f = [None] * n
    for k in range(P):
        o = 6 * k
        iR, iI, xR, xI, uR, uI = y[o:o + 6]
        rr, ri = rot[k]
        f[o] = p["inv_Lf"] * (-p["Rf"] * iR - vR[k] + sp.Rational(1, 2) * mR[k] * Vdc) + wr * iI
        f[o + 1] = p["inv_Lf"] * (-p["Rf"] * iI - vI[k] + sp.Rational(1, 2) * mI[k] * Vdc) - wr * iR
        f[o + 2] = gs[4 * k] * uR                                                    # Ki_GCC
        f[o + 3] = gs[4 * k + 1] * uI                                                # Ki_GCC
        f[o + 4] = gs[4 * k + 2] * (-uR + (rr * irefR - ri * irefI) - iR)           # wp
        f[o + 5] = gs[4 * k + 3] * (-uI + (ri * irefR + rr * irefI) - iI)           # wp
    # DC link: (Ppv - P_inverter) / (C Vdc) = (Ppv / Vdc - P_inverter / Vdc) / C.  The array current per pu
    # PoV = Ppv / Vdc = max(0, A - B exp(kappa Vdc)) and its slope dPoV = -B kappa exp(kappa Vdc) come from the
    # Aux record: Vdc cancels in the second term, so neither the equation nor its Jacobian needs 1 / Vdc.
    f[base] = (inp["PoV"] - Pinv) * p["inv_C"]
    f[base + 1] = gs[4 * P] * (inp["Vdcref"] - Vdc)                                  # Ki_DC
    f[base + 2] = -gs[4 * P + 1] * (inp["Qref"] - Q)                                 # Ki_Q
    f[base + 3] = gs[nfrz] * vd
    f[base + 4] = p["Kp_PLL"] * vd + xPLL + p["dw"]
    # Jacobian: chain rule through the helper symbols
    helpers = {sn: sp.sin(dl), cs: sp.cos(dl)}
    Ppv_f = sp.Function("PoVf")(Vdc)
    J = {}
    for r in range(n):
        fr = f[r].subs(inp["PoV"], Ppv_f)
        fr_full = fr.subs(helpers)
        for c in range(n):
            d = sp.diff(fr_full, y[c])
            if d == 0:
                continue
            d = d.subs(sp.Derivative(Ppv_f, Vdc), inp["dPoV"]).subs(Ppv_f, inp["PoV"])
            d = d.subs({sp.sin(dl): sn, sp.cos(dl): cs})
            d = sp.simplify(d) if P == 1 else d
            if d != 0:
                J[(r, c)] = d
    frozen_rows = []
    for k in range(P):
        frozen_rows += [6 * k + 2, 6 * k + 3, 6 * k + 4, 6 * k + 5]
    frozen_rows += [base + 1, base + 2]
    # Unit-pivot rows: pure integrators (x, xDC, xQ, xPLL) whose row of W is [.. -J_rc .., 1/(h g), ..] with
    # J_rr = 0.  Their gains arrive PRE-SCALED by h*gamma (so f_r and J_rc come out scaled at no cost) and the
    # stepper scales their c_ij/h sums by the same constant: the pivot is exactly 1 and needs no multiply.
    unit_rows = []
    for k in range(P):
        unit_rows += [6 * k + 2, 6 * k + 3]
    unit_rows += [base + 1, base + 2, base + 3]
    for r in unit_rows:
        assert (r, r) not in J, r
    return dict(P=P, mult=mult, n=n, names=names, gs=gs, y=y, f=f, J=J, par=par, inp=inp, frozen=frozen_rows,
                helpers=(sn, cs), unit_rows=unit_rows)


def emit_rhs_structured(m, acc=False):
    """Hand-structured emission of the right-hand side (same function as m['f'], which the Jacobian
    is differentiated from; equality is asserted numerically in gen-time self-check and by
    tests/test_emul_parity.py).  About 30 % fewer FP64 instructions than CSE of the expanded
    sympy expressions: shared phasor products are formed once and every a*b+c is a single FMA."""
    P, mult = m["P"], m["mult"]
    base = 6 * P
    nm = m["names"]
    L = []
    A = L.append
    # acc: f[] arrives pre-loaded (the stage's sum of c_ij/h K_j) and the right-hand side is ADDED to it; every
    # row but the last ends in a multiply that becomes an FMA with that addend, so the sum costs no extra adds
    T = (lambda i: f"f[{i}]") if acc else (lambda i: None)
    c3 = "0.86602540378443864676"
    rots = [("1", "0")] if P == 1 else [("1", "0"), ("-0.5", "-" + c3), ("-0.5", c3)]
    alph = [("1", "0")] if P == 1 else [("1", "0"), ("-0.5", c3), ("-0.5", "-" + c3)]
    for k in range(P):
        iR, iI, xR, xI, uR, uI = ["y_" + nm[6 * k + j] for j in range(6)]
        rr, ri = rots[k]
        vgn = ["in_vg", "in_vgb", "in_vgc"][k] if (P == 3 and mult == 1) else "in_vg"
        vgR = vgn if rr == "1" else f"({rr} * {vgn})"
        vgI = "0.0" if ri == "0" else f"({ri} * {vgn})"
        A(f"    const double vR{k} = fma(p_Rt, {iR}, fma(-p_Xt, {iI}, {vgR}));")
        if ri == "0":
            A(f"    const double vI{k} = fma(p_Xt, {iR}, p_Rt * {iI});")
        else:
            A(f"    const double vI{k} = fma(p_Xt, {iR}, fma(p_Rt, {iI}, {vgI}));")
        A(f"    const double mR{k} = fma(p_Kp_GCC, {uR}, {xR});")
        A(f"    const double mI{k} = fma(p_Kp_GCC, {uI}, {xI});")
        A(f"    const double qs{k} = fma(vI{k}, {iR}, -(vR{k} * {iI}));")
        A(f"    const double ps{k} = fma(mR{k}, {iR}, mI{k} * {iI});")
        ca, sa = alph[k]
        if k == 0:
            A(f"    const double vdk{k} = fma(cs, vR{k}, sn * vI{k});")
        else:
            A(f"    const double ck{k} = fma(cs, {ca}, sn * {sa}), sk{k} = fma(sn, {ca}, -(cs * {sa}));")
            A(f"    const double vdk{k} = fma(vR{k}, ck{k}, vI{k} * sk{k});")
    A("    const double Qs = " + " + ".join(f"qs{k}" for k in range(P)) + ";")
    A("    const double Ps = " + " + ".join(f"ps{k}" for k in range(P)) + ";")
    if P == 1:
        A("    const double vd = vdk0;")
    else:
        A("    const double vd = (1.0 / 3.0) * (" + " + ".join(f"vdk{k}" for k in range(P)) + ");")
    A("    const double wex = fma(p_Kp_PLL, vd, y_xPLL);")
    A("    // (wex + w0) / wb; the product w0 / wb is launch-invariant")
    A("    const double wr = fma(wex, p_inv_wb, p_w0 * p_inv_wb);")
    A("    const double hV = 0.5 * y_Vdc;")
    A("    const double dV = in_Vdcref - y_Vdc;")
    A(f"    // Qref - Q, Q = {0.5 * mult} Qs")
    A(f"    const double dQ = fma({-0.5 * mult}, Qs, in_Qref);")
    A("    const double irefR = fma(p_Kp_DC, dV, y_xDC);")
    A("    const double irefI = fma(-p_Kp_Q, dQ, y_xQ);")
    for k in range(P):
        o = 6 * k
        iR, iI, xR, xI, uR, uI = ["y_" + nm[o + j] for j in range(6)]
        rr, ri = rots[k]
        def mul_or_fma(i, a, b):
            return f"    f[{i}] = fma({a}, {b}, f[{i}]);" if acc else f"    f[{i}] = {a} * {b};"

        if acc:
            A(f"    f[{o}] = fma(wr, {iI}, fma(p_inv_Lf, fma(mR{k}, hV, fma(-p_Rf, {iR}, -vR{k})), f[{o}]));")
            A(f"    f[{o + 1}] = fma(-wr, {iR}, fma(p_inv_Lf, fma(mI{k}, hV, fma(-p_Rf, {iI}, -vI{k})), f[{o + 1}]));")
        else:
            A(f"    f[{o}] = fma(wr, {iI}, p_inv_Lf * fma(mR{k}, hV, fma(-p_Rf, {iR}, -vR{k})));")
            A(f"    f[{o + 1}] = fma(-wr, {iR}, p_inv_Lf * fma(mI{k}, hV, fma(-p_Rf, {iI}, -vI{k})));")
        A(mul_or_fma(o + 2, f"g_{4 * k}", uR))
        A(mul_or_fma(o + 3, f"g_{4 * k + 1}", uI))
        if k == 0:
            A(mul_or_fma(o + 4, f"g_{4 * k + 2}", f"((irefR - {uR}) - {iR})"))
            A(mul_or_fma(o + 5, f"g_{4 * k + 3}", f"((irefI - {uI}) - {iI})"))
        else:
            A(f"    const double rfR{k} = fma({rr}, irefR, -({ri} * irefI)), rfI{k} = fma({ri}, irefR, {rr} * irefI);")
            A(mul_or_fma(o + 4, f"g_{4 * k + 2}", f"((rfR{k} - {uR}) - {iR})"))
            A(mul_or_fma(o + 5, f"g_{4 * k + 3}", f"((rfI{k} - {uI}) - {iI})"))
    # (Ppv - Vdc Ps / 4) / (C Vdc) = (Ppv / Vdc - Ps / 4) / C: Vdc cancels in the second term
    A(mul_or_fma(base, "p_inv_C", f"fma({-0.25 * mult}, Ps, in_PoV)"))      # in_PoV = Ppv / Vdc
    A(mul_or_fma(base + 1, f"g_{4 * P}", "dV"))
    A(mul_or_fma(base + 2, f"(-g_{4 * P + 1})", "dQ") if acc else f"    f[{base + 2}] = -(g_{4 * P + 1} * dQ);")
    A(mul_or_fma(base + 3, f"g_{4 * P + 2}", "vd"))
    A(f"    f[{base + 4}] = (wex + p_dw) + f[{base + 4}];" if acc else f"    f[{base + 4}] = wex + p_dw;")
    return L


def check_structured_rhs(m, lines, acc=False):
    """Gen-time self-check: evaluate the emitted C (as Python) against the symbolic f (acc: against f + preload)."""
    import math
    import random

    rnd = random.Random(0)
    env = {"fma": lambda a, b, c: a * b + c}
    syms = {}
    for s_ in list(m["par"].values()) + list(m["inp"].values()) + m["y"] + m["gs"] + list(m["helpers"]):
        v = rnd.uniform(0.5, 1.5)
        env[str(s_)] = v
        syms[s_] = v
    syms[sp.Symbol("SQ3")] = math.sqrt(3.0)
    pre = [rnd.uniform(-1.0, 1.0) if acc else 0.0 for _ in range(m["n"])]
    f = list(pre)
    env["f"] = f
    for ln in lines:
        if ln.strip().startswith("//"):
            continue
        stmt = ln.strip().rstrip(";").replace("const double ", "")
        for part in _split_decl(stmt):
            exec(part, env)
    for r in range(m["n"]):
        ref = float(m["f"][r].subs(syms)) + pre[r]
        assert abs(f[r] - ref) <= 1e-12 * max(1.0, abs(ref)), (r, f[r], ref)


def _split_decl(stmt):
    """'a = x, b = y' (two declarators) -> ['a = x', 'b = y']; commas inside parentheses are kept."""
    out, depth, cur = [], 0, ""
    for ch in stmt:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def ccode(e):
    return sp.ccode(e, standard="c99")


def emit_block(assignments, tmp_prefix, indent="    "):
    """CSE + C emission of [(lhs, expr)] ; returns list of lines."""
    exprs = [e for _, e in assignments]
    repl, red = sp.cse(exprs, symbols=sp.numbered_symbols(tmp_prefix), optimizations="basic")
    lines = []
    for s, e in repl:
        lines.append(f"{indent}const double {s} = {ccode(e)};")
    for (lhs, _), e in zip(assignments, red):
        lines.append(f"{indent}{lhs} = {ccode(e)};")
    return lines


def elimination_order(n, pattern, P):
    """Diagonal pivots only.  Integrator-like states (x, xDC, xQ, xPLL: rows with one or two
    entries and a 1/(h*gamma) pivot) go first, then the controller outputs u (pivot
    1/(h*gamma)+wp), then the stiff core (currents, Vdc, delta) -- i.e. block elimination down
    to a dense (2P+2)x(2P+2) Schur complement, found here by a greedy minimum-fill search
    within each tier."""
    base = 6 * P
    tier = {}
    for k in range(P):
        o = 6 * k
        tier[o + 2] = tier[o + 3] = 0
        tier[o + 4] = tier[o + 5] = 1
        tier[o] = tier[o + 1] = 2
    tier[base + 1] = tier[base + 2] = tier[base + 3] = 0
    tier[base] = VDC_TIER
    tier[base + 4] = DL_TIER
    pat = set(pattern)
    remaining = set(range(n))
    order = []
    while remaining:
        best = None
        for k in sorted(remaining):
            rows = [r for r in remaining if r != k and (r, k) in pat]
            cols = [c for c in remaining if c != k and (k, c) in pat]
            fill = sum(1 for r in rows for c in cols if (r, c) not in pat)
            cost = (tier[k], fill, len(rows) * len(cols), k)
            if best is None or cost < best[0]:
                best = (cost, k, rows, cols)
        _, k, rows, cols = best
        for r in rows:
            for c in cols:
                pat.add((r, c))
        remaining.remove(k)
        order.append(k)
    return order


def generate(P, mult=1):
    m = build(P, mult)
    n, y, f, J = m["n"], m["y"], m["f"], m["J"]
    tag = f"{P}ph" if mult == 1 else f"{mult}ph_bal"
    cls = f"Model{P}ph" if mult == 1 else f"Model{mult}phBal"
    L = []
    A = L.append
    A("// GENERATED by tools/gen_model.py -- do not edit by hand.")
    A(f"// PV-DER model, {P * mult} phase(s){' (balanced set carried by phase a)' if mult > 1 else ''}, {n} states: " + " ".join(m["names"]))
    A("// Equations: SURVEY.md Appendix A.2-A.4 (restated from the un-vendored pvder package that")
    A("// reference gym_PVDER/envs/PVDER_env.py:26-35 imports); PLL angle stored as delta = wte - w*t.")
    A("#pragma once")
    A('#include "pvder_common.cuh"')
    A("")
    A("namespace pvder {")
    A("")
    A(f"struct {cls} {{")
    A(f"  static constexpr int NS = {n};")
    A(f"  static constexpr int PHASES = {P};            // phases integrated explicitly")
    A(f"  static constexpr int PHASES_OUT = {P * mult};        // phases of the physical model / stored state")
    A(f"  static constexpr int NS_STORE = {6 * P * mult + 5};       // rows of the stored state")
    A(f"  static constexpr bool BALANCED3 = {'true' if mult > 1 else 'false'};")
    A(f"  static constexpr double PMULT = {float(mult)};         // sum over phases = PMULT * explicit phases")
    A(f"  static constexpr int NNZ_J = {len(J)};")
    nf = len(m["frozen"])
    A(f"  static constexpr int NFRZ = {nf};   // freezable rows / freeze-mask bits (rows: " +
      ",".join(m["names"][r] for r in m["frozen"]) + ")")
    unit = set(m["unit_rows"])
    A(f"  static constexpr int NGAIN = {nf + 1 + 2 * P};  // gn[]: the NFRZ effective gains, Ki_PLL, then the reciprocal pivots of the")
    A("                                    // u rows; the gains of the unit-pivot rows and Ki_PLL are pre-scaled by h*gamma (make_gains)")
    A("  // rows whose equation is scaled by h*gamma so that their pivot is exactly 1 (pure integrators)")
    A(f"  static constexpr unsigned UNIT_MASK = {hex(sum(1 << r for r in unit))}u;   // rows: " + ",".join(m["names"][r] for r in sorted(unit)))
    A("  static constexpr PVDER_HD bool unit_row(int i) { return ((UNIT_MASK >> i) & 1u) != 0u; }")
    A(f"  static constexpr int IDX_VDC = {6 * P};")
    A(f"  static constexpr int IDX_DL = {6 * P + 4};")
    A("")
    # ---------- unpack macros
    def unpack(indent="    "):
        out = []
        for i, s in enumerate(y):
            out.append(f"{indent}const double {s} = y[{i}];")
        return out

    par_unpack = [f"    const double p_{nme} = par.{nme};" for nme in PAR]
    gain_unpack = [f"    const double g_{b} = gn[{b}];" for b in range(len(m["frozen"]) + 1)]
    # u rows: pivot 1/(h g) + g_u is untouched by the earlier eliminations, its reciprocal comes with the gains
    u_piv = {}
    for k_ in range(P):
        u_piv[6 * k_ + 4] = nf + 1 + 2 * k_
        u_piv[6 * k_ + 5] = nf + 1 + 2 * k_ + 1
    # ---------- rhs
    A("  // Autonomous right-hand side f(y).  gn[b]: effective gain of freezable row b (0 while clamped).")
    A("  static PVDER_DEV void rhs(const double (&y)[NS], const Params& par, const Inputs& in, const Aux& aux,")
    A("                            const double (&gn)[NGAIN], double (&f)[NS]) {")
    L.extend(par_unpack)
    L.extend(unpack())
    A("    const double in_vg = in.vg, in_vgb = in.vgb, in_vgc = in.vgc, in_Qref = in.Qref, in_Vdcref = in.Vdcref;")
    A("    (void)in_vgb; (void)in_vgc;")
    A("    constexpr double SQ3 = 1.7320508075688772; (void)SQ3;")
    A("    const double sn = aux.sn, cs = aux.cs, in_PoV = aux.PoV;")
    L.extend(gain_unpack)
    rhs_lines = emit_rhs_structured(m)
    check_structured_rhs(m, rhs_lines)
    L.extend(rhs_lines)
    A("  }")
    A("")
    A("  // f += right-hand side (f pre-loaded with a stage's sum of c_ij/h K_j): the same expressions with the")
    A("  // addend folded into each row's last multiply.")
    A("  static PVDER_DEV void rhs_acc(const double (&y)[NS], const Params& par, const Inputs& in, const Aux& aux,")
    A("                                const double (&gn)[NGAIN], double (&f)[NS]) {")
    L.extend(par_unpack)
    L.extend(unpack())
    A("    const double in_vg = in.vg, in_vgb = in.vgb, in_vgc = in.vgc, in_Qref = in.Qref, in_Vdcref = in.Vdcref;")
    A("    (void)in_vgb; (void)in_vgc;")
    A("    constexpr double SQ3 = 1.7320508075688772; (void)SQ3;")
    A("    const double sn = aux.sn, cs = aux.cs, in_PoV = aux.PoV;")
    L.extend(gain_unpack)
    acc_lines = emit_rhs_structured(m, acc=True)
    check_structured_rhs(m, acc_lines, acc=True)
    L.extend(acc_lines)
    A("  }")
    A("")
    # ---------- factor
    pattern = set(J.keys()) | {(i, i) for i in range(n)}
    order = elimination_order(n, pattern, P)
    wname = lambda r, c: f"w_{r}_{c}"
    A("  // LU factors of W = I/(h*gamma) - J(y) with the UNIT_ROW rows scaled by h*gamma (their gains and their")
    A("  // right-hand sides arrive pre-scaled: pivot 1), fixed pattern, fixed pivot order:")
    A("  //   " + " ".join(m["names"][k] for k in order))
    # symbolic elimination to learn the final pattern
    pat = set(pattern)
    pos = {k: i for i, k in enumerate(order)}
    ops = []   # ('inv', k) ('mul', r, k) ('fma', r, c, k) ('new', r, c, k)
    for k in order:
        ops.append(("inv", k))
        rows = sorted(r for r in range(n) if pos[r] > pos[k] and (r, k) in pat)
        cols = sorted(c for c in range(n) if pos[c] > pos[k] and (k, c) in pat)
        for r in rows:
            ops.append(("mul", r, k))
            for c in cols:
                if (r, c) in pat:
                    ops.append(("fma", r, c, k))
                else:
                    pat.add((r, c))
                    ops.append(("new", r, c, k))
    members = sorted(pat)
    nc = 2 * P + 2
    core = (order[-nc:] if CORE_N == 1 else order[-CORE_N:]) if (CORE_INVERSE and nc <= 4) else []
    coreset = set(core)
    # Launch-constant pivots: w_kk is still its initial value ghinv - J_kk when row k is pivoted and
    # J_kk depends on parameters only.  Their reciprocals come from a host-filled table (and a
    # select when the row can be frozen) instead of a division per sub-step.
    par_syms = set(m["par"].values())
    touched = set()
    const_piv = {}          # k -> index into luc[]
    luc_exprs = []          # sympy expressions of 1/(ghinv - J_kk), ghinv symbol = GH
    GH = sp.Symbol("ghinv")
    for op in ops:
        if op[0] == "inv":
            k = op[1]
            jkk = J.get((k, k), sp.Integer(0))
            if CONST_PIVOTS != "off" and (k, k) not in touched and jkk.free_symbols <= par_syms:
                const_piv[k] = len(luc_exprs)
                luc_exprs.append(1 / (GH - jkk))
        elif op[0] in ("fma", "new"):
            touched.add((op[1], op[2]))
    touched_all = {(op[1], op[2]) for op in ops if op[0] in ("fma", "new")}
    A(f"  static constexpr int N_LUC = {len(luc_exprs) + 1};   // launch-constant reciprocal pivots (+ 1/ghinv)")
    A("  // luc[0] = 1/ghinv (pivot of a frozen row); luc[1 + i] = reciprocal of constant pivot i")
    A("  static PVDER_HD void lu_consts(const Params& par, double ghinv, double* luc) {")
    for nme in PAR:
        A(f"    const double p_{nme} = par.{nme}; (void)p_{nme};")
    A("    luc[0] = 1.0 / ghinv;")
    for i, e in enumerate(luc_exprs):
        A(f"    luc[{i + 1}] = {ccode(e)};")
    A("  }")
    A("")
    A("  struct LU {")
    A("    // strictly-lower entries hold multipliers, upper entries hold U, d_k = 1/pivot")
    for (r, c) in members:
        if r != c and not (r in coreset and c in coreset):
            A(f"    double {wname(r, c)};")
    for k in range(n):
        if not (CONST_PIVOTS == "bank" and k in const_piv) and k not in coreset and k not in unit:
            A(f"    double d_{k};")
    for r in core:
        for c in core:
            A(f"    double ci_{r}_{c};   // explicit inverse of the dense core")
    A("  };")
    in_block = lambda o: CRAMER2 and bool(core) and all(i in coreset for i in o[1:])
    nflop_f = sum(2 if o[0] in ("fma",) else 1 for o in ops if not in_block(o)) + (7 if CRAMER2 and core else 0)
    A(f"  static constexpr int LU_ENTRIES = {len(members)};   // incl. {len(members) - len(pattern)} fill-ins")
    A("")
    A("  static PVDER_DEV void factor(const double (&y)[NS], const Params& par, const Inputs& in, const Aux& aux,")
    A("                               const double (&gn)[NGAIN], double ghinv, const double* luc, LU& lu) {")
    L.extend(par_unpack)
    L.extend(unpack())
    A("    const double in_vg = in.vg, in_vgb = in.vgb, in_vgc = in.vgc, in_Qref = in.Qref, in_Vdcref = in.Vdcref;")
    A("    (void)in_vgb; (void)in_vgc;")
    A("    (void)in_Qref; (void)in_Vdcref;")
    A("    constexpr double SQ3 = 1.7320508075688772; (void)SQ3;")
    A("    const double sn = aux.sn, cs = aux.cs, in_dPoV = aux.dPoV;")
    keys = sorted(J.keys())
    L.extend(gain_unpack)
    for (r, c) in keys:
        A(f"    double j_{r}_{c};")
    L.extend(emit_block([(f"j_{r}_{c}", J[(r, c)]) for (r, c) in keys], "q"))
    # W entries
    for i in range(n):
        if i in unit:
            continue                      # pivot exactly 1 (scaled row, J_ii = 0)
        if (i, i) in J:
            A(f"    double w_{i}_{i} = ghinv - j_{i}_{i};")
        else:
            A(f"    double w_{i}_{i} = ghinv;")
    for (r, c) in keys:
        if r != c:
            A(f"    double {wname(r, c)} = -j_{r}_{c};")
    for op in ops:
        if CRAMER2 and core and all(i in coreset for i in op[1:]):
            continue                      # the last 2x2 block is inverted below, not factored
        if op[0] == "inv":
            k = op[1]
            if k in unit:
                pass
            elif k in u_piv and (k, k) not in touched_all and J[(k, k)] == -m["gs"][4 * (k // 6) + 2 + (k % 6 - 4)]:
                A(f"    const double d_{k} = gn[{u_piv[k]}];")
            elif k in const_piv:
                A(f"    const double d_{k} = luc[{const_piv[k] + 1}];")
            else:
                A(f"    const double d_{k} = pvder_rcp(w_{k}_{k});")
        elif op[0] == "mul":
            _, r, k = op
            if k not in unit:
                A(f"    {wname(r, k)} *= d_{k};")
        elif op[0] == "fma":
            _, r, c, k = op
            A(f"    {wname(r, c)} = fma(-{wname(r, k)}, {wname(k, c)}, {wname(r, c)});")
        else:
            _, r, c, k = op
            A(f"    double {wname(r, c)} = -{wname(r, k)} * {wname(k, c)};")
    if core and CRAMER2:
        a, b = core
        A(f"    const double rdet = pvder_rcp(fma({wname(a, a)}, {wname(b, b)}, -({wname(a, b)} * {wname(b, a)})));")
        A(f"    lu.ci_{a}_{a} = {wname(b, b)} * rdet;")
        A(f"    lu.ci_{a}_{b} = -({wname(a, b)} * rdet);")
        A(f"    lu.ci_{b}_{a} = -({wname(b, a)} * rdet);")
        A(f"    lu.ci_{b}_{b} = {wname(a, a)} * rdet;")
    elif core:
        # inverse of the core from its LU (unit-lower multipliers w_r_k, upper w_k_c, d_k = 1/u_kk)
        cpos = {k: i for i, k in enumerate(core)}
        has = lambda r, c: (r, c) in pat
        # Linv (unit lower)
        for ci_, c in enumerate(core):
            for r in core[ci_ + 1:]:
                terms = []
                if has(r, c):
                    terms.append(f"{wname(r, c)}")
                for k in core[ci_ + 1:cpos[r]]:
                    if has(r, k):
                        terms.append(f"{wname(r, k)} * li_{k}_{c}")
                A(f"    const double li_{r}_{c} = -(" + (" + ".join(terms) if terms else "0.0") + ");")
        # Uinv (upper)
        for ci_ in range(len(core) - 1, -1, -1):
            k = core[ci_]
            A(f"    const double ui_{k}_{k} = d_{k};")
        for cj in range(len(core)):
            c = core[cj]
            for ci_ in range(cj - 1, -1, -1):
                k = core[ci_]
                terms = []
                for j in core[ci_ + 1:cj + 1]:
                    if has(k, j):
                        terms.append(f"{wname(k, j)} * ui_{j}_{c}")
                A(f"    const double ui_{k}_{c} = -d_{k} * (" + (" + ".join(terms) if terms else "0.0") + ");")
        for r in core:
            for c in core:
                terms = []
                for j in core[max(cpos[r], cpos[c]):]:
                    li = "1.0" if j == c else f"li_{j}_{c}"
                    terms.append(f"ui_{r}_{j}" if li == "1.0" else f"ui_{r}_{j} * {li}")
                A(f"    lu.ci_{r}_{c} = " + " + ".join(terms) + ";")
    def scaled_u(r, c):
        return SCALED_U and r not in unit and not (CONST_PIVOTS == "bank" and r in const_piv) and pos[c] > pos[r] and r not in coreset
    for (r, c) in members:
        if r != c and not (r in coreset and c in coreset):
            if scaled_u(r, c):
                A(f"    lu.{wname(r, c)} = {wname(r, c)} * d_{r};")
            else:
                A(f"    lu.{wname(r, c)} = {wname(r, c)};")
    for k in range(n):
        if not (CONST_PIVOTS == "bank" and k in const_piv) and k not in coreset and k not in unit:
            A(f"    lu.d_{k} = d_{k};")
    A("  }")
    A("")
    # ---------- solve
    A("  static PVDER_DEV void solve(const LU& lu, const double* luc, double (&b)[NS]) {")
    A("    (void)luc;")
    nflop_s = 0
    for k in order:
        if k in coreset:
            continue
        for r in sorted(r for r in range(n) if pos[r] > pos[k] and (r, k) in pat):
            A(f"    b[{r}] = fma(-lu.{wname(r, k)}, b[{k}], b[{r}]);")
            nflop_s += 2
    if core:
        # core unknowns by a dense mat-vec with the explicit inverse (pairwise sums: depth 3)
        for r in core:
            t = [f"lu.ci_{r}_{c} * b[{c}]" for c in core]
            if len(t) == 4:
                A(f"    const double cx_{r} = fma(lu.ci_{r}_{core[0]}, b[{core[0]}], {t[1]}) + fma(lu.ci_{r}_{core[2]}, b[{core[2]}], {t[3]});")
            else:
                A(f"    const double cx_{r} = " + " + ".join(t) + ";")
            nflop_s += 2 * len(core) - 1
        for r in core:
            A(f"    b[{r}] = cx_{r};")
    for k in reversed(order):
        if k in coreset:
            continue
        if SCALED_U and k not in unit and not (CONST_PIVOTS == "bank" and k in const_piv):
            A(f"    b[{k}] *= lu.d_{k};")
            nflop_s += 1
            # most recently solved unknown last: everything else is off the chain
            for c in sorted((c for c in range(n) if pos[c] > pos[k] and (k, c) in pat), key=lambda c: -pos[c]):
                A(f"    b[{k}] = fma(-lu.{wname(k, c)}, b[{c}], b[{k}]);")
                nflop_s += 2
            continue
        for c in sorted(c for c in range(n) if pos[c] > pos[k] and (k, c) in pat):
            A(f"    b[{k}] = fma(-lu.{wname(k, c)}, b[{c}], b[{k}]);")
            nflop_s += 2
        if k in unit:
            continue                      # unit pivot
        if CONST_PIVOTS == "bank" and k in const_piv:
            A(f"    b[{k}] *= luc[{const_piv[k] + 1}];")
        else:
            A(f"    b[{k}] *= lu.d_{k};")
        nflop_s += 1
    A("  }")
    A("")
    A(f"  static constexpr int FLOPS_LU = {nflop_f};      // executed flops of the symbolic LU")
    A(f"  static constexpr int FLOPS_SOLVE = {nflop_s};   // executed flops of one substitution pair")
    A("};")
    A("")
    A("}  // namespace pvder")
    path = os.path.join(OUT, f"pvder_model_{tag}.cuh")
    with open(path, "w") as fh:
        fh.write("\n".join(L) + "\n")
    print(f"{path}: const pivots {sorted(const_piv)}")
    print(f"{path}: n={n} nnz(J)={len(J)} LU entries={len(members)} order={[m['names'][k] for k in order]} "
          f"lu_flops={nflop_f} solve_flops={nflop_s}")
    return m, order


if __name__ == "__main__":
    generate(1)
    generate(3)
    generate(1, mult=3)      # balanced three-phase set on phase a (balanced reduction)
