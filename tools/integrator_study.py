"""Design study (not product, not test): accuracy of fixed half-cycle (h = 1/120 s) L-stable
one-step schemes against the tight LSODA oracle, on the restated PVDER model.  Used once to
choose the in-kernel integrator; results are quoted in DESIGN.md.

    python tools/integrator_study.py [model_1|model_2] [n_env_steps]
"""
import math
import random
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle.env_oracle import DEFAULT_EVENTS_SPEC, OraclePVDEREnv, create_random_events  # noqa: E402
from oracle.pvder_model import Inputs  # noqa: E402

W = 2.0 * math.pi * 60.0
H = 1.0 / 120.0


def make_aut(model, inp):
    """Autonomous form: last state is delta = wte - w*t."""
    n = model.n

    def f(y):
        d = np.array(model.rhs(list(y), 0.0, inp))
        d[n - 1] -= W
        return d

    def J(y):
        return model.jac(list(y), 0.0, inp)

    return f, J


# ---- Rosenbrock schemes in the (A, C, gamma, m) "KPP" form ----------------------------------
RODAS3 = dict(
    g=0.5,
    A=[[0, 0, 0, 0], [0, 0, 0, 0], [2, 0, 0, 0], [2, 0, 1, 0]],
    C=[[0, 0, 0, 0], [4, 0, 0, 0], [1, -1, 0, 0], [1, -1, -8.0 / 3.0, 0]],
    m=[2, 0, 1, 1],
)
RODAS4 = dict(
    g=0.25,
    A=[[0] * 6,
       [0.1544000000000000e+01, 0, 0, 0, 0, 0],
       [0.9466785280815826, 0.2557011698983284, 0, 0, 0, 0],
       [0.3314825187068521e+01, 0.2896124015972201e+01, 0.9986419139977817, 0, 0, 0],
       [0.1221224509226641e+01, 0.6019134481288629e+01, 0.1253708332932087e+02, -0.6878860361058950, 0, 0],
       [0.1221224509226641e+01, 0.6019134481288629e+01, 0.1253708332932087e+02, -0.6878860361058950, 1, 0]],
    C=[[0] * 6,
       [-0.5668800000000000e+01, 0, 0, 0, 0, 0],
       [-0.2430093356833875e+01, -0.2063599157091915, 0, 0, 0, 0],
       [-0.1073529058151375, -0.9594562251023355e+01, -0.2047028614809616e+02, 0, 0, 0],
       [0.7496443313967647e+01, -0.1024680431464352e+02, -0.3399990352819905e+02, 0.1170890893206160e+02, 0, 0],
       [0.8083246795921522e+01, -0.7981132988064893e+01, -0.3152159432874371e+02, 0.1631930543123136e+02,
        -0.6058818238834054e+01, 0]],
    m=[0.1221224509226641e+01, 0.6019134481288629e+01, 0.1253708332932087e+02, -0.6878860361058950, 1, 1],
)


# Hairer & Wanner's ROS4 code, "L-stable" coefficient set (IV.7, table 7.2): 4 stages, order 4, gamma = 0.57282,
# stage 4 evaluated at Y3 (3 right-hand sides per step).  The kernel's scheme since round 1 (PVDER_SCHEME 4).
ROS4L = dict(
    g=0.57282,
    A=[[0] * 4, [2.0, 0, 0, 0],
       [0.1867943637803922e+01, 0.2344449711399156, 0, 0],
       [0.1867943637803922e+01, 0.2344449711399156, 0, 0]],
    C=[[0] * 4, [-0.7137615036412310e+01, 0, 0, 0],
       [0.2580708087951457e+01, 0.6515950076447975, 0, 0],
       [-0.2137148994382534e+01, -0.3214669691237626, -0.6949742501781779, 0]],
    m=[0.2255570073418735e+01, 0.2870493262186792, 0.4353179431840180, 0.1093502252409163e+01],
)
# Kaps-Rentrop GRK4T: same cost, A(89.3 deg)-stable with |R(inf)| = 0.454 (fast transients decay slowly)
GRK4T = dict(
    g=0.231,
    A=[[0] * 4, [2.0, 0, 0, 0],
       [0.4524708207373116e+01, 0.4163528788597648e+01, 0, 0],
       [0.4524708207373116e+01, 0.4163528788597648e+01, 0, 0]],
    C=[[0] * 4, [-0.5071675338776316e+01, 0, 0, 0],
       [0.6020152728650786e+01, 0.1597506846727117, 0, 0],
       [-0.1856343618686113e+01, -0.8505380858179826e+01, -0.2084075136023187e+01, 0]],
    m=[0.3957503746640777e+01, 0.4624892388363313e+01, 0.6174772638750108, 0.1282612945269037e+01],
)


def order_conditions(tab):
    """Residuals of the eight order conditions up to order 4 (Hairer & Wanner IV.7, table 7.1) and R(inf), from the
    transformed coefficients: Gamma^-1 = diag(1/gamma) - C, alpha = A Gamma, b = m Gamma."""
    g = tab["g"]
    A, Cm, m = np.array(tab["A"], float), np.array(tab["C"], float), np.array(tab["m"], float)
    s = len(m)
    Gam = np.linalg.inv(np.eye(s) / g - Cm)
    al = A @ Gam
    b = m @ Gam
    be = al + Gam - g * np.eye(s)          # beta_ij = alpha_ij + gamma_ij, strictly lower part
    ai, bi = al.sum(axis=1), be.sum(axis=1)
    res = [b.sum() - 1.0,
           b @ bi - (0.5 - g),
           b @ ai ** 2 - 1.0 / 3.0,
           b @ be @ bi - (1.0 / 6.0 - g + g * g),
           b @ ai ** 3 - 0.25,
           b @ (ai * (al @ bi)) - (0.125 - g / 3.0),
           b @ be @ ai ** 2 - (1.0 / 12.0 - g / 3.0),
           b @ be @ be @ bi - (1.0 / 24.0 - g / 2.0 + 1.5 * g * g - g ** 3)]
    r_inf = 1.0 - b @ np.linalg.inv(al + Gam) @ np.ones(s)     # stability function R(z) = 1 + z b (I - z B)^-1 1 at infinity
    return np.array(res), r_inf


def rosenbrock_step(f, J, y, h, tab):
    g, A, C, m = tab["g"], tab["A"], tab["C"], tab["m"]
    s = len(m)
    n = len(y)
    Wm = np.eye(n) / (h * g) - J(y)
    lu = np.linalg.inv(Wm)
    K = []
    for i in range(s):
        Yi = y.copy()
        for j in range(i):
            if A[i][j]:
                Yi = Yi + A[i][j] * K[j]
        rhs = f(Yi)
        for j in range(i):
            if C[i][j]:
                rhs = rhs + (C[i][j] / h) * K[j]
        K.append(lu @ rhs)
    out = y.copy()
    for i in range(s):
        if m[i]:
            out = out + m[i] * K[i]
    return out


def radau5_step(f, J, y, h, newton=8):
    s6 = math.sqrt(6.0)
    A = np.array([[(88 - 7 * s6) / 360, (296 - 169 * s6) / 1800, (-2 + 3 * s6) / 225],
                  [(296 + 169 * s6) / 1800, (88 + 7 * s6) / 360, (-2 - 3 * s6) / 225],
                  [(16 - s6) / 36, (16 + s6) / 36, 1.0 / 9.0]])
    n = len(y)
    Jm = J(y)
    M = np.eye(3 * n) - h * np.kron(A, Jm)
    Minv = np.linalg.inv(M)
    Z = np.zeros(3 * n)
    for _ in range(newton):
        F = np.concatenate([f(y + Z[i * n:(i + 1) * n]) for i in range(3)])
        R = Z - h * (np.kron(A, np.eye(n)) @ F)
        dZ = -Minv @ R
        Z = Z + dZ
        if np.max(np.abs(dZ)) < 1e-14:
            break
    return y + Z[2 * n:]


def sdirk4_step(f, J, y, h, newton=6):
    g = 0.25
    A = [[g], [0.5, g], [17 / 50, -1 / 25, g], [371 / 1360, -137 / 2720, 15 / 544, g],
         [25 / 24, -49 / 48, 125 / 16, -85 / 12, g]]
    n = len(y)
    Minv = np.linalg.inv(np.eye(n) - h * g * J(y))
    Ks = []
    for i in range(5):
        base = y.copy()
        for j in range(i):
            base = base + h * A[i][j] * Ks[j]
        k = Ks[-1].copy() if Ks else f(y)
        for _ in range(newton):
            r = k - f(base + h * g * k)
            dk = -Minv @ r
            k = k + dk
            if np.max(np.abs(dk)) < 1e-13:
                break
        Ks.append(k)
    out = y.copy()
    for j in range(5):
        out = out + h * A[4][j] * Ks[j]
    return out


SCHEMES = {
    "rodas3": lambda f, J, y, h: rosenbrock_step(f, J, y, h, RODAS3),
    "rodas4": lambda f, J, y, h: rosenbrock_step(f, J, y, h, RODAS4),
    "ros4l": lambda f, J, y, h: rosenbrock_step(f, J, y, h, ROS4L),
    "grk4t": lambda f, J, y, h: rosenbrock_step(f, J, y, h, GRK4T),
    "rodas3x2": lambda f, J, y, h: rosenbrock_step(f, J, rosenbrock_step(f, J, y, h / 2, RODAS3), h / 2, RODAS3),
    "rodas4x2": lambda f, J, y, h: rosenbrock_step(f, J, rosenbrock_step(f, J, y, h / 2, RODAS4), h / 2, RODAS4),
    "sdirk4": sdirk4_step,
    "radau5": radau5_step,
}


def order_check():
    """Convergence-order check of the tableaux on a stiff scalar-ish problem."""
    lam = -50.0

    def f(y):
        return np.array([lam * (y[0] - math.cos(y[1])) - math.sin(y[1]), 1.0])

    def J(y):
        return np.array([[lam, lam * math.sin(y[1]) - math.cos(y[1])], [0.0, 0.0]])

    for name, tab in (("rodas3", RODAS3), ("rodas4", RODAS4), ("ros4l", ROS4L), ("grk4t", GRK4T)):
        res, r_inf = order_conditions(tab)
        print(f"order conditions {name}: max |residual| orders 1-3 {np.abs(res[:4]).max():.1e}, order 4 "
              f"{np.abs(res[4:]).max():.1e}; R(inf) = {r_inf:+.3f}")
    for name in ("rodas3", "rodas4", "ros4l", "grk4t", "sdirk4", "radau5"):
        errs = []
        for N in (10, 20, 40, 80):
            y = np.array([1.0, 0.0])
            for _ in range(N):
                y = SCHEMES[name](f, J, y, 1.0 / N)
            errs.append(abs(y[0] - math.cos(1.0)))
        orders = [math.log2(errs[i] / errs[i + 1]) for i in range(3)]
        print(f"order check {name}: errs {errs[0]:.2e}..{errs[-1]:.2e} observed orders {np.round(orders, 2)}")


def main():
    model_type = sys.argv[1] if len(sys.argv) > 1 else "model_1"
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    order_check()
    spec = {k: dict(v) for k, v in DEFAULT_EVENTS_SPEC.items()}
    spec["voltage"].update(min=0.90, max=1.02)
    spec["insolation"].update(ENABLE=True)
    events = create_random_events(spec, random.Random(7))
    arng = random.Random(3)
    actions = [arng.randrange(5) for _ in range(nsteps)]
    env = OraclePVDEREnv(model_type=model_type, solver="tight", events=events)
    env.reset()
    n = env.model.n
    # truth at every sub-step
    truth = [env.y.copy()]
    inputs = []
    t0 = time.time()
    for a in actions:
        dQ = env.delQ_pu if a == 1 else -env.delQ_pu if a == 2 else 0.0
        dV = env.delVdc_pu if a == 3 else -env.delVdc_pu if a == 4 else 0.0
        env.Q_ref += dQ
        env.Vdc_ref += dV
        for _ in range(2 * env.n):
            inputs.append((env.events.vgrid(env.t()), env.events.sinsol(env.t()), env.Q_ref, env.Vdc_ref))
            env._integrate(1)
            truth.append(env.y.copy())
    truth = np.array(truth)
    tt = np.arange(len(truth)) * H
    truth_aut = truth.copy()
    truth_aut[:, n - 1] -= W * tt
    print(f"truth: {len(truth) - 1} sub-steps in {time.time() - t0:.1f}s")
    names = ["iR", "iI", "xR", "xI", "uR", "uI"] * env.params.phases + ["Vdc", "xDC", "xQ", "xPLL", "delta"]
    for name, stepper in SCHEMES.items():
        y = truth_aut[0].copy()
        errs = np.zeros((len(truth), n))
        t0 = time.time()
        for k, (vg, si, q, vd) in enumerate(inputs):
            inp = Inputs(Vgrid=vg, Sinsol=si, Q_ref=q, Vdc_ref=vd, freeze=(False,) * (4 * env.params.phases + 2))
            f, J = make_aut(env.model, inp)
            y = stepper(f, J, y, H)
            errs[k + 1] = np.abs(y - truth_aut[k + 1])
        scale = 1e-8 / 1e-5 + np.abs(truth_aut)
        rel = errs / scale
        es = 2 * env.n  # env-step boundaries
        print(f"{name:9s} {time.time() - t0:5.1f}s  max abs err all sub-steps: "
              f"{errs.max():.2e}; after 0.25 s: {errs[30:].max():.2e}; "
              f"env-step pts rel(max over states) {rel[es::es].max():.2e}; "
              f"after 1st env step {rel[2 * es::es].max():.2e}")
        worst = errs[es::es].max(axis=0)
        print("           per-state max abs err at env-step pts:",
              " ".join(f"{nm}:{e:.1e}" for nm, e in zip(names[-11:], worst[-11:])))


if __name__ == "__main__":
    main()
