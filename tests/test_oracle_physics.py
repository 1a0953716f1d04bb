"""First-principles pins of the CPU oracle.

The reference ships no golden vectors and OpenSim cannot run here (parity
unpinned), so the oracle's physics is checked against independent
derivations: a numerical Lagrangian built from finite differences of a
separate numpy forward-kinematics routine, finite differences of path
lengths, closed-form contact / limit-force values, energy conservation.
"""
import numpy as np
import pytest

from bioimitation_gym_b200 import refmotion

G = 9.80665


def _rand_state(cm, rng, in_air=True):
    t = cm.tables
    names = cm.dof_names
    q = np.zeros(t.n_dof)
    for i, n in enumerate(names):
        if n in ("pelvis_tx", "pelvis_ty", "pelvis_tz"):
            q[i] = {"pelvis_tx": 0.3, "pelvis_ty": 1.3 if in_air else 0.93, "pelvis_tz": -0.1}[n]
        elif n.startswith("knee"):
            q[i] = rng.uniform(-1.2, -0.1)
        elif n.startswith("hip_flexion"):
            q[i] = rng.uniform(-0.3, 0.9)
        else:
            q[i] = rng.uniform(-0.3, 0.3)
    u = rng.uniform(-2.0, 2.0, t.n_dof)
    return q, u


def _skew_vec(S):
    return np.array([S[2, 1] - S[1, 2], S[0, 2] - S[2, 0], S[1, 0] - S[0, 1]]) * 0.5


def _jacobians(cm, q, eps=1e-6):
    t = cm.tables
    nb, nd = t.n_bodies, t.n_dof
    R0, p0 = refmotion.host_fk(cm, q)
    Jv = np.zeros((nb, 3, nd))
    Jw = np.zeros((nb, 3, nd))
    for k in range(nd):
        dq = np.zeros(nd)
        dq[k] = eps
        Rp, pp = refmotion.host_fk(cm, q + dq)
        Rm, pm = refmotion.host_fk(cm, q - dq)
        for b in range(nb):
            com = np.asarray(t.body_com[b][:])
            Jv[b, :, k] = ((pp[b] + Rp[b] @ com) - (pm[b] + Rm[b] @ com)) / (2 * eps)
            Jw[b, :, k] = _skew_vec((Rp[b] - Rm[b]) / (2 * eps) @ R0[b].T)
    return R0, p0, Jv, Jw


def _mass_matrix(cm, q):
    t = cm.tables
    R0, p0, Jv, Jw = _jacobians(cm, q)
    M = np.zeros((t.n_dof, t.n_dof))
    for b in range(t.n_bodies):
        i6 = t.body_inertia[b][:]
        Ic = np.array([[i6[0], i6[3], i6[4]], [i6[3], i6[1], i6[5]], [i6[4], i6[5], i6[2]]])
        Iw = R0[b] @ Ic @ R0[b].T
        M += t.body_mass[b] * Jv[b].T @ Jv[b] + Jw[b].T @ Iw @ Jw[b]
    return M


def _potential(cm, q):
    t = cm.tables
    R, p = refmotion.host_fk(cm, q)
    V = 0.0
    for b in range(t.n_bodies):
        V += t.body_mass[b] * G * (p[b] + R[b] @ np.asarray(t.body_com[b][:]))[1]
    return V


@pytest.mark.parametrize("key", ["2d_torque", "3d_torque", "3d_torque_prosthetic"])
def test_mass_matrix_and_bias_match_numerical_lagrangian(oracle_lib, models, rng, key):
    cm = models[key]
    t = cm.tables
    for _ in range(3):
        q, u = _rand_state(cm, rng)
        ev = oracle_lib.eval_dynamics(t, q, u)
        M = _mass_matrix(cm, q)
        assert np.allclose(ev["mass_matrix"], ev["mass_matrix"].T, atol=1e-12)
        assert np.linalg.eigvalsh(ev["mass_matrix"]).min() > 0
        np.testing.assert_allclose(ev["mass_matrix"], M, rtol=1e-6, atol=1e-7)
        # bias = Mdot u - dT/dq + dV/dq  (Lagrange), all by central differences
        eps = 1e-4
        nd = t.n_dof
        Mdot_u = (_mass_matrix(cm, q + eps * u) - _mass_matrix(cm, q - eps * u)) / (2 * eps) @ u
        dT = np.zeros(nd)
        dV = np.zeros(nd)
        for k in range(nd):
            dq = np.zeros(nd)
            dq[k] = eps
            dT[k] = 0.5 * u @ ((_mass_matrix(cm, q + dq) - _mass_matrix(cm, q - dq)) / (2 * eps)) @ u
            dV[k] = (_potential(cm, q + dq) - _potential(cm, q - dq)) / (2 * eps)
        bias = Mdot_u - dT + dV
        scale = np.abs(bias).max()
        np.testing.assert_allclose(ev["bias"], bias, atol=2e-5 * scale)
        # forward dynamics solves M udot = -bias
        np.testing.assert_allclose(ev["mass_matrix"] @ ev["udot"], -ev["bias"], atol=1e-9 * scale)


def test_free_fall_is_gravity(oracle_lib, models):
    cm = models["2d_torque"]
    q = np.zeros(9)
    q[2] = 1.5
    ev = oracle_lib.eval_dynamics(cm.tables, q, np.zeros(9))
    expect = np.zeros(9)
    expect[2] = -G
    np.testing.assert_allclose(ev["udot"], expect, atol=1e-9)


def _numpy_path_length(cm, q, i):
    t = cm.tables
    R, p = refmotion.host_fk(cm, q)
    pts = []
    for k in range(t.mus_pt_begin[i], t.mus_pt_begin[i] + t.mus_pt_count[i]):
        kind, b, d = t.pt_kind[k], t.pt_body[k], t.pt_dof[k]
        if kind == 1 and not (t.pt_range[k][0] - 1e-5 <= q[d] <= t.pt_range[k][1] + 1e-5):
            continue
        if kind == 2:
            loc = np.array([refmotion._func(t, t.pt_func[k][c], q[d]) for c in range(3)])
        else:
            loc = np.asarray(t.pt_loc[k][:])
        pts.append(p[b] + R[b] @ loc)
    return sum(np.linalg.norm(pts[j + 1] - pts[j]) for j in range(len(pts) - 1))


@pytest.mark.parametrize("key", ["2d_muscle", "3d_muscle"])
def test_path_length_speed_and_moment_arms(oracle_lib, models, rng, key):
    cm = models[key]
    t = cm.tables
    nm, nd = t.n_muscles, t.n_dof
    q, u = _rand_state(cm, rng)
    L, Ld = oracle_lib.path_lengths(t, q, u)
    for i in range(nm):
        assert abs(L[i] - _numpy_path_length(cm, q, i)) < 1e-12
    eps = 1e-6
    Lp, _ = oracle_lib.path_lengths(t, q + eps * u, u)
    Lm, _ = oracle_lib.path_lengths(t, q - eps * u, u)
    np.testing.assert_allclose(Ld, (Lp - Lm) / (2 * eps), atol=1e-7)
    # generalized muscle force = -sum_i T_i dL_i/dq  (moment arm = -dL/dq)
    act = np.full(nm, 0.3)
    lm = np.array([oracle_lib.equilibrium_lm(t, i, L[i], 0.3) for i in range(nm)])
    ev = oracle_lib.eval_dynamics(t, q, u, act=act, lm=lm, ctrl=act)
    tq = models[key.replace("muscle", "torque")].tables
    ev0 = oracle_lib.eval_dynamics(tq, q, u)
    Qmus = ev0["bias"] - ev["bias"]
    dLdq = np.zeros((nm, nd))
    for k in range(nd):
        dq = np.zeros(nd)
        dq[k] = eps
        a, _ = oracle_lib.path_lengths(t, q + dq, u)
        b, _ = oracle_lib.path_lengths(t, q - dq, u)
        dLdq[:, k] = (a - b) / (2 * eps)
    expect = -(ev["tendon_force"][:, None] * dLdq).sum(axis=0)
    np.testing.assert_allclose(Qmus, expect, atol=1e-5 * max(1.0, np.abs(expect).max()))


def test_muscle_curves_known_answers(oracle_lib, models):
    t = models["2d_muscle"].tables
    ce = lambda c, x: oracle_lib.curve_eval(t, c, x)[0]
    assert abs(ce(0, 1.0) - 1.0) < 1e-9          # f_L(1) = 1
    assert ce(0, 0.4441) == pytest.approx(0.0, abs=1e-12)
    assert ce(0, 1.8123) == pytest.approx(0.0, abs=1e-12)
    assert abs(ce(1, 0.0) - 1.0) < 1e-9          # f_V(0) = 1
    assert abs(ce(1, -1.0)) < 1e-12              # f_V(-1) = 0
    assert abs(ce(1, 1.0) - 1.4) < 1e-12         # f_V(1) = 1.4
    assert abs(oracle_lib.curve_eval(t, 1, 0.0)[1] - 5.0) < 1e-6   # isometric slope
    assert abs(ce(2, 1.7) - 1.0) < 1e-12         # f_PE(1.7) = 1
    assert ce(2, 1.0) == pytest.approx(0.0, abs=1e-12)
    assert abs(ce(3, 1.049) - 1.0) < 1e-12       # f_T at 4.9 % strain
    assert abs(oracle_lib.curve_eval(t, 3, 1.049)[1] - 1.375 / 0.049) < 1e-6
    assert ce(3, 0.99) == 0.0                    # slack tendon carries no force
    # monotone where they must be
    xs = np.linspace(-1.2, 1.2, 400)
    fv = np.array([ce(1, x) for x in xs])
    assert np.all(np.diff(fv) >= -1e-12)
    xs = np.linspace(0.9, 1.1, 400)
    ft = np.array([ce(3, x) for x in xs])
    assert np.all(np.diff(ft) >= -1e-12)


def test_table_error_against_exact_bezier(models):
    """The 512-interval cubic-Hermite tables (used by BOTH the oracle and the
    CUDA kernels) stay within 1e-6 (normalised force) of the exact
    quintic-Bezier curves they tabulate."""
    from bioimitation_gym_b200 import curves
    t = models["2d_muscle"].tables
    for ci, cv in enumerate(curves.default_curves()):
        xs = np.linspace(cv.x0 - 0.05, cv.x1 + 0.05, 4001)
        exact, _ = cv.eval(xs)
        tab = np.ctypeslib.as_array(t.curve_tab)[ci]
        approx, _ = curves.hermite_eval(t.curve_x0[ci], t.curve_x1[ci], tab, xs)
        assert np.abs(exact - approx).max() < 1e-6, ci


def test_equilibrium_and_fibre_velocity_solve(oracle_lib, models, rng):
    cm = models["2d_muscle"]
    t = cm.tables
    nm = t.n_muscles
    q, u = _rand_state(cm, rng)
    L, _ = oracle_lib.path_lengths(t, q, u)
    for a in (0.05, 0.5, 1.0):
        lm = np.array([oracle_lib.equilibrium_lm(t, i, L[i], a) for i in range(nm)])
        ev = oracle_lib.eval_dynamics(t, q, u, act=np.full(nm, a), lm=lm, ctrl=np.full(nm, a))
        # static equilibrium: fibre velocity from the damped-equilibrium solve vanishes
        clamped = lm <= np.ctypeslib.as_array(t.mus_lm_min)[:nm] * (1 + 1e-12)
        assert np.all(np.abs(ev["lmdot"][~clamped]) < 1e-7)
        # force balance along the tendon: F_fibre * cos(alpha) = F_tendon
        h = np.ctypeslib.as_array(t.mus_height)[:nm]
        cosa = np.sqrt(lm ** 2 - h ** 2) / lm
        np.testing.assert_allclose(ev["fiber_force"][~clamped] * cosa[~clamped],
                                   ev["tendon_force"][~clamped], rtol=1e-8, atol=1e-6)
    # away from equilibrium the Newton solve still balances the forces
    lm = np.ctypeslib.as_array(t.mus_lopt)[:nm] * rng.uniform(0.6, 1.3, nm)
    ev = oracle_lib.eval_dynamics(t, q, u, act=np.full(nm, 0.4), lm=lm, ctrl=np.full(nm, 0.4),
                                  newton_iters=50)
    h = np.ctypeslib.as_array(t.mus_height)[:nm]
    cosa = np.sqrt(lm ** 2 - h ** 2) / lm
    np.testing.assert_allclose(ev["fiber_force"] * cosa, ev["tendon_force"], rtol=1e-8, atol=1e-6)


def test_activation_dynamics_closed_form(oracle_lib, models):
    t = models["2d_muscle"].tables
    nm = t.n_muscles
    q = np.zeros(9)
    q[2] = 1.3
    lm = np.ctypeslib.as_array(t.mus_lopt)[:nm].copy()
    a = np.full(nm, 0.2)
    up = oracle_lib.eval_dynamics(t, q, np.zeros(9), act=a, lm=lm, ctrl=np.full(nm, 1.0))
    dn = oracle_lib.eval_dynamics(t, q, np.zeros(9), act=a, lm=lm, ctrl=np.full(nm, 0.0))
    np.testing.assert_allclose(up["adot"], (1.0 - 0.2) / (0.01 * (0.5 + 1.5 * 0.2)), rtol=1e-12)
    # excitation is clamped to minimum_activation = 0.01
    np.testing.assert_allclose(dn["adot"], (0.01 - 0.2) / (0.04 / (0.5 + 1.5 * 0.2)), rtol=1e-12)


def test_hunt_crossley_closed_form(oracle_lib, models):
    cm = models["2d_torque"]
    t = cm.tables
    q = np.zeros(9)
    q[2] = 2.0
    u = np.zeros(9)
    ev = oracle_lib.eval_dynamics(t, q, u)
    assert np.all(ev["contact"] == 0)            # separated: no force
    # lower the model until the right heel sphere alone penetrates by d
    bottoms = refmotion.sphere_bottoms(cm, q)
    heel = int(np.argmin(bottoms))
    d = 0.004
    q[2] = 2.0 - bottoms[heel] - d
    vy, vx = -0.3, 0.05
    u[2], u[1] = vy, vx
    ev = oracle_lib.eval_dynamics(t, q, u)
    k = 0.5 * 2.0e6 ** (2.0 / 3.0)
    vrel = abs(vx) / 0.1
    f = ff = 0.0
    for s_ in range(t.n_spheres):                 # pure translation: same velocity everywhere
        ds = bottoms[heel] + d - bottoms[s_]
        if ds <= 0:
            continue
        R = t.sph_radius[s_]
        fH = 4.0 / 3.0 * k * ds * np.sqrt(R * k * ds)
        fs = fH * (1 + 1.5 * 1.0 * (-vy))
        f += fs
        ff += fs * (min(vrel, 1.0) * (0.8 + 2 * (0.8 - 0.8) / (1 + vrel ** 2)) + 0.6 * abs(vx))
    n = 1
    F = ev["contact"][:, :3].sum(axis=0)
    np.testing.assert_allclose(F[1], n * f, rtol=1e-10)
    np.testing.assert_allclose(F[0], -n * ff, rtol=1e-10)
    # whole-model check: total vertical acceleration of the COM = (F - m g) / m
    com_acc_y = (F[1] - t.total_mass * G) / t.total_mass
    M = ev["mass_matrix"]
    # generalized momentum along pelvis_ty is the total linear momentum in y
    assert abs((M[2] @ ev["udot"]) / t.total_mass - com_acc_y) < 1e-6 * abs(com_acc_y) + \
        abs(ev["bias"][2] + F[1] - t.total_mass * G) / t.total_mass + 1e-9


def test_coordinate_limit_force(oracle_lib, models):
    cm = models["2d_torque"]
    t = cm.tables
    q = np.zeros(9)
    q[2] = 1.5
    u = np.zeros(9)
    knee = cm.dof_names.index("knee_angle_r")
    li = cm.limit_names.index("knee_limit_r")
    q[knee] = -0.5                                # inside [-140, 0] deg
    assert oracle_lib.eval_dynamics(t, q, u)["limit_force"][li] == 0.0
    K = 20.0 * 180.0 / np.pi                      # N m / rad
    D = 0.25 * 180.0 / np.pi
    w = np.deg2rad(10.0)
    q[knee] = 0.3                                 # beyond upper limit + transition
    u[knee] = 0.7
    f = oracle_lib.eval_dynamics(t, q, u)["limit_force"][li]
    assert f == pytest.approx(-K * 0.3 - D * 0.7, rel=1e-12)
    q[knee] = 0.5 * w                             # half-way through the transition
    s = 10 * 0.5 ** 3 - 15 * 0.5 ** 4 + 6 * 0.5 ** 5
    f = oracle_lib.eval_dynamics(t, q, u)["limit_force"][li]
    assert f == pytest.approx(-K * s * 0.5 * w - D * s * 0.7, rel=1e-12)
    q[knee] = np.deg2rad(-140.0) - 0.3
    u[knee] = 0.0
    f = oracle_lib.eval_dynamics(t, q, u)["limit_force"][li]
    assert f == pytest.approx(K * 0.3, rel=1e-12)


def test_energy_conservation_in_flight(oracle_lib, models, rng):
    """Conservative system (no contact, no limits hit, zero torques): RK4 on
    the oracle RHS keeps T + V constant."""
    cm = models["3d_torque"]
    t = cm.tables
    q, u = _rand_state(cm, rng)
    u *= 0.5

    def energy(q, u):
        return 0.5 * u @ _mass_matrix(cm, q) @ u + _potential(cm, q)

    def f(q, u):
        return u, oracle_lib.eval_dynamics(t, q, u)["udot"]

    E0 = energy(q, u)
    h = 1e-3
    for _ in range(100):
        k1 = f(q, u)
        k2 = f(q + 0.5 * h * k1[0], u + 0.5 * h * k1[1])
        k3 = f(q + 0.5 * h * k2[0], u + 0.5 * h * k2[1])
        k4 = f(q + h * k3[0], u + h * k3[1])
        q = q + h / 6 * (k1[0] + 2 * k2[0] + 2 * k3[0] + k4[0])
        u = u + h / 6 * (k1[1] + 2 * k2[1] + 2 * k3[1] + k4[1])
    ev = oracle_lib.eval_dynamics(t, q, u)
    assert np.all(ev["limit_force"] == 0) and np.all(ev["contact"] == 0)
    assert abs(energy(q, u) - E0) < 1e-6 * abs(E0)


def test_simm_spline_interpolates_knots(oracle_lib, models):
    t = models["2d_muscle"].tables
    M = oracle_lib.M
    for f in range(t.n_funcs):
        if t.func_kind[f] != M["BIO_FUNC_SPLINE"]:
            continue
        kb, n = t.func_knot_begin[f], t.func_knot_count[f]
        for i in range(n):
            y, d1, d2 = oracle_lib.func_eval(t, f, t.knot_x[kb + i])
            assert abs(y - t.knot_c[kb + i][0]) < 1e-12
        # C1/C2 continuity at interior knots and derivative consistency
        for i in range(1, n - 1):
            x = t.knot_x[kb + i]
            l = oracle_lib.func_eval(t, f, x - 1e-9)
            r = oracle_lib.func_eval(t, f, x + 1e-9)
            assert abs(l[1] - r[1]) < 1e-6 and abs(l[2] - r[2]) < 1e-4 * max(1.0, abs(l[2]))
        x = 0.5 * (t.knot_x[kb] + t.knot_x[kb + n - 1])
        e = 1e-6
        y0, d1, d2 = oracle_lib.func_eval(t, f, x)
        yp = oracle_lib.func_eval(t, f, x + e)[0]
        ym = oracle_lib.func_eval(t, f, x - e)[0]
        assert abs((yp - ym) / (2 * e) - d1) < 1e-7
        # linear extrapolation with the end slope
        x0 = t.knot_x[kb]
        a = oracle_lib.func_eval(t, f, x0 - 0.5)
        assert abs(a[0] - (t.knot_c[kb][0] - 0.5 * t.knot_c[kb][1])) < 1e-12 and a[2] == 0.0


def test_adaptive_baseline_integrator_tracks_fine_fixed_step(oracle_lib):
    """The CPU-baseline scheme (error-controlled Runge-Kutta-Merson, accuracy 1e-3: stand-in for the
    reference's default opensim.Manager integrator) against RK4 with h = 50 us over 10 control steps."""
    from bioimitation_gym_b200 import registry, tasks
    orc = oracle_lib
    out = {}
    for name, cfg in (("adaptive", dict(integrator="adaptive_rkm")), ("fine", dict(integrator="rk4", substeps=200))):
        full = tasks.merged_config(dict(num_envs=4, seed=0, **cfg))
        spec, cm, ref, task = registry.build_env_tables("MuscleWalkingImitation2D-v0", full, None, None)
        rt = orc.RefTables(ref["q"], ref["u"], ref["body_pos"], ref["com_pos"])
        env = orc.OracleVecEnv(cm.tables, task, rt, 4, seed=0)
        env.reset()
        rng = np.random.default_rng(0)
        for _ in range(10):
            env.step(rng.uniform(0, 1, (4, 14)))
        out[name] = env.get_state()
    assert np.max(np.abs(out["adaptive"]["q"] - out["fine"]["q"])) < 5e-3
    assert np.max(np.abs(out["adaptive"]["act"] - out["fine"]["act"])) < 5e-3
