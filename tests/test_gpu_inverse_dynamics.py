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
    # M^-1 M a = a
    a = rng.normal(0, 1, (n, nd))
    back = idm.multiplyByMInv(0.0, q, idm.multiplyByM(0.0, q, a)).cpu().numpy()
    assert np.max(np.abs(back - a)) < 1e-8
    idm.close()


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
