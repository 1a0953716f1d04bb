"""Batched inverse-dynamics operator set (bioimitation_gym_b200/inverse_dynamics.py) on the GPU,
against the CPU oracle's mass matrix / bias and against the forward dynamics of the same library.
Mirrors the reference's InverseDynamics (inverse_dynamics.cpp:44-200).  fp64, tolerance 1e-8."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _states(idm, n, seed=3):
    from bioimitation_gym_b200 import assets
    ref = idm.full.ref
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, ref["q"].shape[0], n)
    t = idm.full.cm.tables
    cd = np.ctypeslib.as_array(t.coord_dof)[:t.n_coords]
    q = np.zeros((n, idm.n_dof))
    u = np.zeros((n, idm.n_dof))
    for k, d in enumerate(cd):
        if d >= 0:
            q[:, d] = ref["q"][rows, k]
            u[:, d] = ref["u"][rows, k]
    q += rng.normal(0, 0.02, q.shape)
    return q, u, rng


@pytest.mark.parametrize("model", ["2D", "3D"])
def test_operators_against_oracle_and_forward_dynamics(model):
    import torch
    from bioimitation_gym_b200.inverse_dynamics import InverseDynamics
    from oracle import oracle as orc
    n = 64
    idm = InverseDynamics(model, num_envs=n)
    q, u, rng = _states(idm, n)
    t = idm.full.cm.tables
    nd, na = idm.n_dof, t.n_act
    M = idm.mass_matrix(q).cpu().numpy()
    f = idm.calculateTotalForces(0.0, q, u).cpu().numpy()
    g = idm.calculateGravity(0.0, q).cpu().numpy()
    c = idm.calculateCoriolis(0.0, q, u).cpu().numpy()
    c0 = idm.calculateCoriolis(0.0, q, np.zeros_like(u)).cpu().numpy()
    worst_M = worst_f = 0.0
    for i in range(0, n, 4):
        o = orc.eval_dynamics(t, q[i], u[i], ctrl=np.zeros(na))
        worst_M = max(worst_M, np.max(np.abs(M[i] - o["mass_matrix"])) / np.max(np.abs(o["mass_matrix"])))
        worst_f = max(worst_f, np.max(np.abs(f[i] - o["bias"]) / np.maximum(np.abs(o["bias"]), 1.0)))
    assert worst_M < 1e-8 and worst_f < 1e-8
    assert np.allclose(M, np.swapaxes(M, 1, 2), atol=1e-10) and np.all(np.linalg.eigvalsh(M) > 0)
    assert np.max(np.abs(c0)) < 1e-9                       # no velocity, no Coriolis
    # gravity on the vertical pelvis translation = -m g
    names = idm.dof_names
    ty = names.index("pelvis_ty")
    assert np.allclose(g[:, ty], t.total_mass * t.gravity[1], rtol=1e-9)
    # M qddot + c = g + tau: forward dynamics of the force-free tree with random torques
    tau_act = rng.uniform(-30, 30, (n, na))
    act_dof = np.ctypeslib.as_array(t.act_dof)[:na]
    tau = np.zeros((n, nd))
    tau[:, act_dof] = tau_act
    idm.bare.set_state(dict(q=q, u=u))
    ev = idm.bare.eval_debug(torch.as_tensor(tau_act))
    udot = ev["udot"].cpu().numpy()
    lhs = np.einsum("nij,nj->ni", M, udot) + c
    assert np.max(np.abs(lhs - (g + tau)) / np.maximum(np.abs(g + tau), 1.0)) < 1e-8
    # residual forces of the full model reproduce the torques that produced the motion
    idm.full.set_state(dict(q=q, u=u))
    udot_full = idm.full.eval_debug(torch.as_tensor(tau_act))["udot"].clone()
    res = idm.calculateResidualForces(0.0, q, u, udot_full).cpu().numpy()
    assert np.max(np.abs(res - tau) / np.maximum(np.abs(tau), 1.0)) < 1e-7
    # the in-library operators (bio_id_multiply_m / _minv: composite inertias + the step path's sparse L^T D L)
    # against dense algebra on the oracle-checked mass matrix, and M^-1 M a = a
    a = rng.normal(0, 1, (n, nd))
    Ma = idm.multiplyByM(0.0, q, a).cpu().numpy()
    assert np.max(np.abs(Ma - np.einsum("nij,nj->ni", M, a))) < 1e-9 * np.max(np.abs(Ma))
    Mia = idm.multiplyByMInv(0.0, q, a).cpu().numpy()
    want = np.stack([np.linalg.solve(M[i], a[i]) for i in range(n)])
    assert np.max(np.abs(Mia - want) / np.maximum(np.abs(want), 1.0)) < 1e-8
    back = idm.multiplyByMInv(0.0, q, idm.multiplyByM(0.0, q, a)).cpu().numpy()
    assert np.max(np.abs(back - a)) < 1e-8
    # residual = M qddot + total forces; stable PD against its closed form (example_position_control.py:143-190)
    res2 = idm.calculateResidualForces(0.0, q, u, a).cpu().numpy()
    assert np.max(np.abs(res2 - (np.einsum("nij,nj->ni", M, a) + f)) / np.maximum(np.abs(res2), 1.0)) < 1e-8
    tau_pd = rng.normal(0, 20, (n, nd))
    kd, h = 35.0, 0.01
    got = idm.stable_pd(0.0, q, u, tau_pd, kd, h).cpu().numpy()
    res0 = idm.calculateResidualForces(0.0, q, u, np.zeros_like(q)).cpu().numpy()
    want = np.stack([tau_pd[i] - kd * h * np.linalg.solve(M[i] + kd * h * np.eye(nd), tau_pd[i] - res0[i]) for i in range(n)])
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)) < 1e-8
    launches = idm.bare.launch_count
    idm.multiplyByMInv(0.0, q, a)
    assert idm.bare.launch_count - launches <= 3           # state transposes + one operator kernel, no library solver
    idm.close()


def test_operators_in_fp32_and_through_the_raw_abi():
    """The operator kernels in the production precision, called through the named C entry points."""
    import ctypes
    import torch
    from bioimitation_gym_b200.inverse_dynamics import InverseDynamics
    n = 32
    idm = InverseDynamics("3D", num_envs=n, dtype="float32")
    idd = InverseDynamics("3D", num_envs=n, dtype="float64")
    q, u, rng = _states(idm, n)
    a = rng.normal(0, 1, (n, idm.n_dof))
    m32 = idm.multiplyByM(0.0, q, a).double().cpu().numpy()
    m64 = idd.multiplyByM(0.0, q, a).cpu().numpy()
    assert np.max(np.abs(m32 - m64)) < 2e-5 * np.max(np.abs(m64))
    i32 = idm.multiplyByMInv(0.0, q, m64).double().cpu().numpy()
    assert np.max(np.abs(i32 - a)) < 5e-3                  # cond(M) ~ 1e4: foot vs trunk inertia
    env = idd.bare
    x = torch.as_tensor(a, dtype=torch.float64, device=env.device).contiguous()
    out = torch.empty_like(x)
    s = ctypes.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)
    assert env.lib.bio_id_multiply_m(env.handle, x.data_ptr(), out.data_ptr(), s) == 0
    assert np.allclose(out.cpu().numpy(), m64, rtol=1e-12, atol=1e-12)
    assert env.lib.bio_id_apply(env.handle, 99, x.data_ptr(), None, None, out.data_ptr(), s) != 0
    assert b"unknown operator" in env.lib.bio_last_error()
    idm.close()
    idd.close()


def test_controllers_and_list_interface():
    from bioimitation_gym_b200.inverse_dynamics import InverseDynamics
    idm = InverseDynamics("2D", num_envs=1)
    q, u, rng = _states(idm, 1)
    ql, ul = q[0].tolist(), u[0].tolist()
    f = idm.calculateTotalForces(0.0, ql, ul)
    assert isinstance(f, list) and len(f) == idm.n_dof
    tau_pd = rng.normal(0, 5, idm.n_dof).tolist()
    # Kd = 0: the stable PD law is the PD torque itself
    assert np.allclose(idm.stable_pd(0.0, ql, ul, tau_pd, 0.0, 0.01), tau_pd)
    tau = idm.stable_pd(0.0, ql, ul, tau_pd, 50.0, 0.01)
    assert np.all(np.isfinite(tau)) and not np.allclose(tau, tau_pd)
    # computed torque with zero PD part is plain inverse dynamics
    ad = rng.normal(0, 1, idm.n_dof).tolist()
    ct = idm.computed_torque(0.0, ql, ul, ad, [0.0] * idm.n_dof)
    assert np.allclose(ct, idm.calculateResidualForces(0.0, ql, ul, ad))
    idm.close()
