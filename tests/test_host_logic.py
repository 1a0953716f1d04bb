"""CPU tests of the host side: table ABI, model compiler, task tables, oracle
env semantics (obs layout / reward / done / reset), storage I/O."""
import ctypes
import re
import math
import os

import numpy as np
import pytest

from bioimitation_gym_b200 import assets, ctables as ct, model_compiler as mc, refmotion, registry, tasks
from bioimitation_gym_b200.osim_parser import parse_osim

from conftest import REF_DATA


# ----------------------------------------------------------------- ABI / library
def test_header_structs_match_oracle_build(oracle_lib):
    L = oracle_lib.lib()
    assert L.orc_sizeof_model_tables() == ctypes.sizeof(ct.BioModelTables)
    assert L.orc_sizeof_task_config() == ctypes.sizeof(ct.BioTaskConfig)


def test_cuda_library_loads_and_exports_every_symbol():
    """No GPU needed: dlopen + symbol lookup + ABI size check only."""
    import re
    from bioimitation_gym_b200 import backend
    lib = backend.load_library()
    hdr = open(ct.HEADER).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(bio_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no entry points parsed from the header"
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
        assert name in backend.C_API, "python binding misses %s" % name
    assert lib.bio_abi_version() == ct.MACROS["BIO_ABI_VERSION"]


def test_no_cpu_fallback():
    import torch
    from bioimitation_gym_b200 import backend
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(backend.BioError):
        backend.VecEnv("MuscleWalkingImitation2D-v0", dict(num_envs=2))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.dirname(assets.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(root, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle\b", r"libbio_oracle", r"bio_oracle\.c", r"orc_[a-z_]+\s*\(",
                            r"#include\s+\"[^\"]*oracle"):
                    assert not re.search(pat, txt, re.M), "%s uses the oracle (%s)" % (f, pat)


# ----------------------------------------------------------------- model compiler
def test_model_facts(models):
    m2, m3 = models["2d_muscle"].tables, models["3d_muscle"].tables
    assert (m2.n_bodies, m2.n_dof, m2.n_muscles, m2.n_pathpts, m2.n_spheres, m2.n_limits) == (7, 9, 14, 48, 6, 6)
    assert (m3.n_bodies, m3.n_dof, m3.n_muscles, m3.n_pathpts, m3.n_coords) == (7, 14, 22, 68, 17)
    assert abs(m2.total_mass - 75.1646) < 1e-3 and abs(m3.total_mass - 41.5) < 1e-6
    assert models["2d_torque"].tables.n_act == 7 and models["3d_torque"].tables.n_act == 11
    assert models["3d_muscle_prosthetic"].tables.n_muscles == 19
    assert models["3d_muscle_prosthetic"].tables.n_dof == 12
    assert models["2d_muscle"].dof_names[:3] == ["pelvis_tilt", "pelvis_tx", "pelvis_ty"]
    assert models["2d_muscle"].muscle_names[:7] == ["hamstrings_r", "glut_max_r", "iliopsoas_r", "vasti_r",
                                                    "gastroc_r", "soleus_r", "tib_ant_r"]
    # contact stiffness: two equal materials, k* = 0.5 * k^(2/3)
    assert m2.sph_k[0] == pytest.approx(0.5 * 2.0e6 ** (2.0 / 3.0))
    # limit forces are stored in radians
    assert m2.lim_kup[0] == pytest.approx(20.0 * 180.0 / math.pi)
    assert m2.lim_qup[0] == pytest.approx(math.radians(120.0))
    # pelvis_ty default of the predictive model (opensim_utils.py:215)
    assert m2.dof_default_q[2] == pytest.approx(1.02)


def test_merged_body_mass_properties(models):
    """pelvis+torso (2D) merged rigidly: mass, COM and inertia follow the parallel-axis theorem."""
    cm = models["2d_muscle"]
    t = cm.tables
    assert t.body_mass[0] == pytest.approx(11.777 + 34.2366)
    com = (11.777 * np.array([-0.0707, 0, 0]) + 34.2366 * (np.array([-0.1007, 0.0815, 0]) + np.array([-0.03, 0.32, 0]))) \
        / (11.777 + 34.2366)
    np.testing.assert_allclose(np.asarray(t.body_com[0][:]), com, atol=1e-12)
    # foot = talus + calcn + toes
    assert t.body_mass[3] == pytest.approx(0.1 + 1.25 + 0.2166)


@pytest.mark.needs_reference
@pytest.mark.parametrize("key", list(assets.MODEL_SPECS))
def test_fixtures_match_a_fresh_compile_of_the_reference_xml(key):
    sub, surgery = assets.MODEL_SPECS[key]
    fresh = assets.model_from_osim(os.path.join(REF_DATA, sub, "scale", "model_scaled.osim"), surgery)
    stored = assets.load_model(key)
    a, b = ct.struct_to_dict(fresh.tables), ct.struct_to_dict(stored.tables)
    for k in a:
        np.testing.assert_array_equal(np.asarray(a[k]), np.asarray(b[k]), err_msg=k)
    assert fresh.dof_names == stored.dof_names and fresh.muscle_names == stored.muscle_names


@pytest.mark.needs_reference
def test_in_memory_surgery_matches_the_committed_model_predictive_osim():
    """construct_predictive_model (opensim_utils.py:204-222) restated in memory vs
    the one artefact the reference ships: data/02905/02905_PRE/scale/model_predictive.osim."""
    base = os.path.join(REF_DATA, "02905", "02905_PRE", "scale")
    ours = mc.compile_model(mc.construct_predictive_model(parse_osim(os.path.join(base, "model_scaled.osim"))))
    theirs = mc.compile_model(parse_osim(os.path.join(base, "model_predictive.osim")))
    a, b = ct.struct_to_dict(ours.tables), ct.struct_to_dict(theirs.tables)
    for k in a:
        if np.asarray(a[k]).dtype.kind == "f":
            np.testing.assert_allclose(np.asarray(a[k]), np.asarray(b[k]), rtol=1e-12, atol=1e-12, err_msg=k)
        else:
            np.testing.assert_array_equal(np.asarray(a[k]), np.asarray(b[k]), err_msg=k)


def test_simm_spline_is_c2_interpolant():
    x = np.array([-2.0944, -1.74533, -1.39626, -1.0472, -0.698132, -0.349066, -0.174533, 0.197344, 0.337395,
                  0.490178, 1.52146, 2.0944])
    y = np.array([-0.0032, 0.00179, 0.00411, 0.0041, 0.00212, -0.001, -0.0031, -0.005227, -0.005435, -0.005574,
                  -0.005435, -0.00525])
    b, c, d = mc.simm_spline_coefficients(x, y)
    for i in range(len(x) - 1):
        dx = x[i + 1] - x[i]
        assert y[i] + dx * (b[i] + dx * (c[i] + dx * d[i])) == pytest.approx(y[i + 1], abs=1e-14)
        assert b[i] + dx * (2 * c[i] + 3 * dx * d[i]) == pytest.approx(b[i + 1], abs=1e-12)
        assert 2 * c[i] + 6 * dx * d[i] == pytest.approx(2 * c[i + 1], abs=1e-10)
    # two knots: straight line
    b2, c2, d2 = mc.simm_spline_coefficients([0.0, 2.0], [1.0, 5.0])
    assert list(b2) == [2.0, 2.0] and not c2.any() and not d2.any()


# ----------------------------------------------------------------- task tables
@pytest.mark.parametrize("env_id,dim,na", [
    ("MuscleWalkingImitation2D-v0", 138, 14), ("MuscleWalkingImitation3D-v0", 201, 22),
    ("TorqueWalkingImitation2D-v0", 96, 7), ("TorqueWalkingImitation3D-v0", 135, 11),
    ("MuscleLockedKneeImitation3D-v0", 192, 19), ("MusclePalsyImitation3D-v0", 201, 22)])
def test_observation_and_action_sizes(env_id, dim, na, oracle_lib):
    spec, cm, ref, task = registry.build_env_tables(env_id, {})
    assert task.obs_dim == dim and cm.tables.n_act == na
    assert oracle_lib.lib().orc_obs_dim(ctypes.byref(cm.tables), ctypes.byref(task)) == dim
    _, _, _, t2 = registry.build_env_tables(env_id, dict(use_target_obs=False, use_GRF=False))
    n_tx = 1
    assert t2.obs_dim == dim - 2 * (cm.tables.n_coords - n_tx) - 12


def test_all_17_env_ids_are_known():
    assert len(tasks.ENV_SPECS) == 17
    for env_id, spec in tasks.ENV_SPECS.items():
        assert spec.model in assets.MODEL_SPECS


def test_task_constants_follow_the_reference_classes():
    _, _, ref, t = registry.build_env_tables("MuscleWalkingImitation2D-v0", {})
    assert (t.cycle, t.n_steps, t.reset_max_index, t.feed_mean_action) == (132, 264, 132, 1)
    assert (t.term_height, t.term_limit_force, t.term_acc) == (0.75, 1000.0, 1e4)
    assert (t.w_imitate, t.w_effort, t.w_action) == (0.8, 0.2, 0.1) and t.horizon == 5 and t.dt == 0.01
    _, _, ref, t = registry.build_env_tables("MuscleWalkingImitation2D-v0", dict(mode="test"))
    assert t.n_steps == ref["q"].shape[0] - 2 and t.reset_max_index == 0
    _, _, ref3, t3 = registry.build_env_tables("MuscleWalkingImitation3D-v0", {})
    assert (t3.cycle, t3.reset_max_index, t3.term_feet_cross, t3.reward_use_feet) == (50, 50, 1, 1)
    assert t3.n_steps == ref3["q"].shape[0] - 2 and t3.action_r_scale == 0.5
    _, _, _, tp = registry.build_env_tables("MusclePalsyImitation3D-v0", {})
    assert tp.feed_mean_action == 0 and tp.perturb_negative_only == 1
    _, cmq, _, tq = registry.build_env_tables("TorqueWalkingImitation3D-v0", {})
    # PD index quirk of torque_walking_imitation_env3D.py:131-132 is reproduced verbatim
    assert list(tq.pd_v_coord[:11]) == [0, 4, 5, 6, 8, 8, 9, 10, 11, 12, 13]
    assert list(tq.pd_x_coord[:11]) == [0, 1, 2, 6, 7, 8, 9, 10, 11, 12, 13]
    assert list(tq.pd_kp[:11]) == [100, 100, 100, 100, 100, 50, 100, 100, 100, 100, 50]
    assert tq.effort_torque == 1 and tq.n_reward_terms == 4


def test_missing_reference_motion_is_reported(monkeypatch):
    from bioimitation_gym_b200 import assets
    monkeypatch.setattr(assets, "DATA_DIR", "/nonexistent")
    with pytest.raises(FileNotFoundError):
        registry.build_env_tables("MuscleJumpingImitation3D-v0", {}, model=assets.load_model("3d_muscle"))


def test_every_env_id_has_tables_and_the_jump_reference_mirrors(models):
    """All 17 registered IDs construct (reference bioimitation/__init__.py:23-133); the jumping tasks
    run their table forwards then mirrored (muscle_jumping_imitation_env2D.py:73-76,290-294: N = 2 (rows - 2),
    cycle = N / 2, effort progress along pelvis_ty :341)."""
    for env_id in tasks.ENV_SPECS:
        spec, cm, ref, task = registry.build_env_tables(env_id, {})
        assert ref["q"].shape[1] == cm.tables.n_coords and task.reset_max_index <= ref["q"].shape[0] - 2
    spec, cm, ref, task = registry.build_env_tables("MuscleJumpingImitation2D-v0", {})
    rows = ref["q"].shape[0]
    assert task.n_steps == 2 * (rows - 2) and task.cycle == rows - 2 and task.ref_mirror == 1 and task.effort_use_dy == 1
    ty = cm.coord_names.index("pelvis_ty")
    # the table is the way UP: the pelvis ends at its apex, above the standing height, feet off the ground
    assert ref["q"][-1, ty] > ref["q"][0, ty] + 0.15 and abs(ref["u"][-1, ty]) < 0.35
    dof_cols = [cm.coord_names.index(n) for n in cm.dof_names]
    assert refmotion.sphere_bottoms(cm, ref["q"][-1, dof_cols]).min() > 0.1
    assert abs(refmotion.sphere_bottoms(cm, ref["q"][0, dof_cols]).min() + 0.008) < 2e-3
    spec, cm3, ref3, task3 = registry.build_env_tables("MuscleRunningImitation3D-v0", {})
    assert task3.cycle == 70 and ref3["q"].shape == (282, 17)
    # locked coordinates keep their compiled value, the right foot stays on the right
    for j, n in enumerate(cm3.coord_names):
        if n not in cm3.dof_names:
            assert np.allclose(ref3["q"][:, j], cm3.tables.coord_const[j])
    bn = list(ref3["body_names"])
    assert (ref3["body_pos"][:, bn.index("calcn_r"), 2] > ref3["body_pos"][:, bn.index("calcn_l"), 2]).all()


# ----------------------------------------------------------------- oracle env semantics
def _oracle_env(oracle_lib, env_id, n=4, seed=3, **cfg):
    spec, cm, ref, task = registry.build_env_tables(env_id, cfg)
    rt = oracle_lib.RefTables(ref["q"], ref["u"], ref["body_pos"], ref["com_pos"])
    return oracle_lib.OracleVecEnv(cm.tables, task, rt, n, seed=seed), spec, cm, ref, task


def test_observation_layout_matches_app_c(oracle_lib):
    env, spec, cm, ref, task = _oracle_env(oracle_lib, "MuscleWalkingImitation2D-v0", n=3)
    obs = env.reset()
    st = env.get_state()
    t = cm.tables
    for i in range(3):
        istep = st["istep"][i]
        q, u = st["q"][i], st["u"][i]
        assert obs[i, 0] == pytest.approx((istep / 132.0) % 1.0)
        np.testing.assert_allclose(obs[i, 1:8], q[[0, 3, 4, 5, 6, 7, 8]])          # pelvis_tx/ty dropped
        np.testing.assert_allclose(obs[i, 8:17], u)
        np.testing.assert_allclose(obs[i, 26:34], np.delete(ref["q"][istep + 1], 1))  # target without pelvis_tx
        np.testing.assert_allclose(obs[i, 34:42], np.delete(ref["u"][istep + 1], 1))
        ev = oracle_lib.eval_dynamics(t, q, u, st["act"][i], st["lm"][i], np.zeros(14), newton_iters=task.newton_iters)
        np.testing.assert_allclose(obs[i, 17:26], ev["udot"], rtol=1e-12)
        rel = ev["obs_pos"] - np.array([q[1], q[2], 0.0])
        np.testing.assert_allclose(obs[i, 42:69], rel.reshape(-1), atol=1e-12)
        np.testing.assert_allclose(obs[i, 69:72], ev["com_pos"] - np.array([q[1], q[2], 0.0]), atol=1e-12)
        np.testing.assert_allclose(obs[i, 72:81], ev["obs_vel"][:3].reshape(-1), atol=1e-12)
        np.testing.assert_allclose(obs[i, 84:126].reshape(14, 3)[:, 0], st["act"][i])
        np.testing.assert_allclose(obs[i, 84:126].reshape(14, 3)[:, 1], st["lm"][i])
        w = t.total_mass * 9.80665
        np.testing.assert_allclose(obs[i, 126:129], ev["contact"][0, :3] / w, atol=1e-12)
        np.testing.assert_allclose(obs[i, 129:132], ev["contact"][0, 3:] / (w * 1.8), atol=1e-12)
    # reset state: reference row, default activation 0.05, static fibre equilibrium
    assert np.all(st["act"] == 0.05) and np.all(st["first"] == 1)
    np.testing.assert_allclose(st["q"], ref["q"][st["istep"]], atol=0)


def test_reward_is_the_reference_formula(oracle_lib):
    """Recompute get_reward / calc_cost_of_transport (env2D.py:267-403) in numpy from
    the oracle's post-step state and compare with the oracle's reward."""
    env, spec, cm, ref, task = _oracle_env(oracle_lib, "MuscleWalkingImitation2D-v0", n=2, auto_reset=False)
    env.reset()
    t = cm.tables
    rng = np.random.default_rng(0)
    last = None
    hist = [None, None]
    old_px = np.zeros(2)
    for k in range(6):
        a = rng.uniform(0, 1, (2, 14))
        for i in range(2):
            if hist[i] is None:
                hist[i] = [a[i].copy() for _ in range(5)]
                if last is None:
                    last = [None, None]
                last[i] = a[i].copy()
            hist[i] = hist[i][1:] + [a[i].copy()]
        outs = [env.step_env_debug(i, a[i]) for i in range(2)]
        st = env.get_state()
        for i in range(2):
            obs, rew, reason, terms, ev = outs[i]
            curr = np.mean(hist[i], axis=0)
            istep = st["istep"][i]
            q = st["q"][i]
            q_err = np.mean((q - ref["q"][istep]) ** 2)
            com = np.array(ev.com_pos[:])
            com_err = np.mean((com - ref["com_pos"][istep]) ** 2)
            position_r, com_r = math.exp(-30 * q_err), math.exp(-20 * com_err)
            act, lm = st["act"][i], st["lm"][i]
            total = 1.51 * t.total_mass
            for mi in range(14):
                mass = t.mus_fiso[mi] / 0.25e6 * 1059.7 * t.mus_lopt[mi]
                lam = mc.SLOW_TWITCH_2D[mi]
                e = min(max(curr[mi], 0.0), 1.0)
                fa = 40 * lam * math.sin(0.5 * math.pi * e) + 133 * (1 - lam) * (1 - math.cos(0.5 * math.pi * e))
                fm = 74 * lam * math.sin(0.5 * math.pi * act[mi]) + 111 * (1 - lam) * (1 - math.cos(0.5 * math.pi * act[mi]))
                ln, v = lm[mi] / t.mus_lopt[mi], ev.lmdot[mi]
                g = 0.5 if ln < 0.5 else (ln if ln < 1.0 else (-2 * ln + 3 if ln < 1.5 else 0.0))
                total += mass * fa + mass * g * fm + max(0.0, 0.25 * ev.fiber_force[mi] * -v) + \
                    max(0.0, ev.active_fiber_force[mi] * -v)
            effort = total / (20 * 14 ** 2)
            effort_r = math.exp(-effort / max(q[1] - old_px[i] + 1, 1))
            action_r = math.exp(-np.linalg.norm(curr - last[i]))
            expect = (0.5 + 0.8) * position_r * com_r + 0.2 * effort_r + 0.1 * action_r
            assert rew == pytest.approx(expect, rel=1e-12)
            assert terms[0] == pytest.approx(position_r) and terms[1] == pytest.approx(com_r)
            assert terms[4] == pytest.approx(math.exp(-2 * np.linalg.norm(act)))
            last[i] = curr
            old_px[i] = q[1]


def test_termination_reasons(oracle_lib):
    M = ct.MACROS
    env, spec, cm, ref, task = _oracle_env(oracle_lib, "TorqueWalkingImitation2D-v0", n=1, auto_reset=False)
    env.reset()
    st = env.get_state()
    # torso below 0.75 m
    st["q"][0][:] = 0
    st["q"][0][2] = 0.55
    st["u"][0][:] = 0
    env.set_state(st)
    _, _, reason, _, _ = env.step_env_debug(0, np.zeros(7))
    assert reason == M["BIO_DONE_HEIGHT"]
    # knee hyper-extended far beyond the limit: |limit force| > 1000
    env.reset()
    st = env.get_state()
    st["q"][0][:] = 0
    st["q"][0][2] = 1.5
    st["q"][0][4] = 2.5       # still > 1 rad past the limit after the 10 ms step
    st["u"][0][:] = 0
    env.set_state(st)
    _, _, reason, _, _ = env.step_env_debug(0, np.zeros(7))
    assert reason == M["BIO_DONE_LIMIT_FORCE"]
    # horizon
    env.reset()
    st = env.get_state()
    st["q"][0][:] = 0
    st["q"][0][2] = 1.6
    st["u"][0][:] = 0
    st["istep"][0] = task.n_steps - 1
    env.set_state(st)
    _, _, reason, _, _ = env.step_env_debug(0, np.zeros(7))
    assert reason == M["BIO_DONE_HORIZON"]
    # NaN action is zeroed, not propagated (opensim_wrapper.py:93-95)
    env.reset()
    obs, rew, reason, _, _ = env.step_env_debug(0, np.full(7, np.nan))
    assert np.isfinite(obs).all() and np.isfinite(rew)


def test_reset_index_distribution_and_determinism(oracle_lib):
    L = oracle_lib.lib()
    draws = np.array([L.orc_rand(7, e, 0, 1) % 133 for e in range(4000)])
    assert draws.min() == 0 and draws.max() == 132
    counts = np.bincount(draws, minlength=133)
    assert counts.min() > 8 and counts.max() < 60           # roughly uniform
    assert L.orc_rand(7, 5, 0, 1) == L.orc_rand(7, 5, 0, 1)
    assert L.orc_rand(7, 5, 0, 1) != L.orc_rand(7, 5, 1, 1)
    env, *_ = _oracle_env(oracle_lib, "MuscleWalkingImitation2D-v0", n=64, mode="test")
    env.reset()
    assert np.all(env.get_state()["istep"] == 0)            # test mode starts at row 0 (env2D.py:141-142)


def test_pd_controller_of_torque_envs(oracle_lib):
    """First step from rest at the reference pose with target = current angles and
    zero speeds gives zero torque (torque env2D.py:134-139)."""
    env, spec, cm, ref, task = _oracle_env(oracle_lib, "TorqueWalkingImitation2D-v0", n=1, auto_reset=False)
    env.reset()
    st = env.get_state()
    st["u"][0][:] = 0
    env.set_state(st)
    x = st["q"][0][[0, 3, 4, 5, 6, 7, 8]]
    env.step_env_debug(0, x)
    assert np.allclose(env.get_state()["last_action"][0], 0.0, atol=1e-12)
    # and a unit error on the hip gives Kp = 100 (mean over a history filled with the same value)
    env.reset()
    st = env.get_state()
    st["u"][0][:] = 0
    env.set_state(st)
    x = st["q"][0][[0, 3, 4, 5, 6, 7, 8]].copy()
    x[1] += 1.0
    env.step_env_debug(0, x)
    assert env.get_state()["last_action"][0][1] == pytest.approx(100.0)


# ----------------------------------------------------------------- reference motion / storage
def test_storage_round_trip(tmp_path):
    t = np.arange(20) * 0.01
    data = np.stack([t, np.sin(t), np.cos(t)], axis=1)
    p = str(tmp_path / "x.sto")
    refmotion.write_storage(p, "Coordinates", ["time", "a", "b"], data, in_degrees=True)
    labels, back, deg = refmotion.read_storage(p)
    assert labels == ["time", "a", "b"] and deg
    np.testing.assert_allclose(back, data, atol=1e-8)


def test_lowpass_and_resample():
    t = np.arange(400) * 0.005
    x = np.sin(2 * np.pi * 1.0 * t)[:, None] + 0.2 * np.sin(2 * np.pi * 40.0 * t)[:, None]
    tt, xr = refmotion.resample_linear(t, x, 0.01)
    assert abs(tt[1] - tt[0] - 0.01) < 1e-12 and xr.shape[0] == 200
    y = refmotion.lowpass_zero_phase(xr, 100.0, 6.0)
    clean = np.sin(2 * np.pi * 1.0 * tt)
    assert np.abs(y[20:-20, 0] - clean[20:-20]).max() < 0.03    # 40 Hz removed, 1 Hz kept, no phase lag


def test_reference_tables_are_consistent(models):
    ref = assets.load_ref("3d_walking")
    cm = models["3d_muscle"]
    assert ref["q"].shape == (364, 17) and ref["coord_names"] == cm.coord_names
    row = 100
    dof_cols = [cm.coord_names.index(n) for n in cm.dof_names]
    bp, com = refmotion.body_kinematics(cm, ref["q"][row, dof_cols], ref["body_names"])
    np.testing.assert_allclose(bp, ref["body_pos"][row], atol=1e-12)
    np.testing.assert_allclose(com, ref["com_pos"][row], atol=1e-12)
    # speeds are the time derivative of the coordinates
    np.testing.assert_allclose(ref["u"][1:-1], (ref["q"][2:] - ref["q"][:-2]) / 0.02, atol=1e-9)
    # the stance foot touches the ground (Hertz penetration of a few mm .. cm)
    low = min(refmotion.sphere_bottoms(cm, ref["q"][r, dof_cols]).min() for r in range(0, 364, 10))
    assert -0.06 < low < 0.005


def test_device_replay_buffer_matches_reference_semantics():
    """Ring buffer with the reference's field names (replay_buffer.py:9-74, dataset.py:46-66),
    batch inserts, mask = 0 on real episode ends, uniform sampling."""
    import torch
    from bioimitation_gym_b200.rollout import Batch, DeviceReplayBuffer
    buf = DeviceReplayBuffer(obs_dim=3, action_dim=2, capacity=10, device="cpu")
    obs = torch.arange(12, dtype=torch.float32).reshape(4, 3)
    act = torch.ones(4, 2)
    rew = torch.tensor([1.0, 2.0, 3.0, 4.0])
    done = torch.tensor([0, 1, 0, 1], dtype=torch.uint8)
    tl = torch.tensor([0, 0, 0, 1], dtype=torch.uint8)
    buf.insert_step(obs, act, rew, done, obs + 100, time_limit_done=tl)
    assert len(buf) == 4 and buf.insert_index == 4
    assert buf.masks[:4].tolist() == [1.0, 0.0, 1.0, 1.0] and buf.dones_float[:4].tolist() == [0.0, 1.0, 0.0, 1.0]
    buf.insert_step(obs, act, rew, done, obs + 100)
    buf.insert_step(obs, act, rew, done, obs + 100)           # wraps around
    assert len(buf) == 10 and buf.insert_index == 2
    assert torch.equal(buf.observations[0], obs[2]) and torch.equal(buf.observations[1], obs[3])
    buf.insert(obs[0].numpy(), act[0].numpy(), 5.0, 1.0, 0.0, obs[1].numpy())   # single transition, reference signature
    assert buf.insert_index == 3 and float(buf.rewards[2]) == 5.0
    b = buf.sample(32, generator=torch.Generator().manual_seed(0))
    assert isinstance(b, Batch) and b.observations.shape == (32, 3) and b.masks.shape == (32,)
    assert torch.allclose(b.next_observations[b.rewards != 5.0] - b.observations[b.rewards != 5.0], torch.tensor(100.0))


def test_state_storage_labels_and_file_round_trip(tmp_path, models):
    from bioimitation_gym_b200 import refmotion
    from bioimitation_gym_b200.rollout import state_labels
    cm = models["2d_muscle"]
    labels = state_labels(cm)
    assert labels[0] == "time" and len(labels) == 1 + 2 * cm.tables.n_dof + 2 * cm.tables.n_muscles
    assert "pelvis_tilt/value" in labels and "pelvis_tilt/speed" in labels
    data = np.random.default_rng(0).normal(size=(5, len(labels)))
    path = str(tmp_path / "simulation_States.sto")
    refmotion.write_storage(path, "simulation_States", labels, data)
    got_labels, got, in_deg = refmotion.read_storage(path)
    assert got_labels == labels and not in_deg and np.allclose(got, data, atol=1e-7)


def test_rllib_vector_env_adapter_interface():
    """vector_reset / reset_at / vector_step over a stand-in backend (no GPU, no ray)."""
    import torch
    from bioimitation_gym_b200.rllib_adapter import BioVectorEnv

    class Fake:
        num_envs, obs_dim, n_act = 3, 4, 2
        action_low, action_high = np.zeros(2), np.ones(2)
        dtype = torch.float32

        def __init__(self):
            self.t = 0
            self.torch = torch

        def reset_np(self):
            return np.zeros((3, 4), dtype=np.float32)

        def step_np(self, a):
            self.t += 1
            assert a.shape == (3, 2) and a.dtype == np.float32
            done = np.array([0, 1, 0], dtype=bool)
            return (np.full((3, 4), float(self.t), dtype=np.float32), np.arange(3.0, dtype=np.float32), done,
                    {"all_rewards": np.ones((3, 5), dtype=np.float32)})

        def close(self):
            pass

    v = BioVectorEnv("MuscleWalkingImitation2D-v0", backend_env=Fake())
    obs = v.vector_reset()
    assert len(obs) == 3 and obs[0].shape == (4,)
    o, r, d, info = v.vector_step([[0.1, 0.2]] * 3)
    assert len(o) == 3 and r == [0.0, 1.0, 2.0] and d == [False, True, False] and len(info[1]["all_rewards"]) == 5
    assert np.array_equal(v.reset_at(1), o[1])          # the finished env already holds its next episode's first obs
    # every array handed out is a copy: later steps (the GPU overwrites its page-locked buffers) do not change it
    keep = o[0].copy()
    v.vector_step([[0.1, 0.2]] * 3)
    v.vector_step([[0.1, 0.2]] * 3)
    assert np.array_equal(o[0], keep) and not np.shares_memory(o[0], v._obs)
    assert v.action_space.shape == (2,) and v.observation_space.shape == (4,) and v.get_sub_environments() == []


def test_bioimitation_shim_serves_the_reference_import_path(monkeypatch):
    """`import gym, bioimitation; gym.make(id, config=cfg)` (reference tests/test_env.py:15-26): the shim
    registers the 17 IDs under the reference's module paths.  gym is not installed here, so a minimal
    stand-in registry plays its part (register / make resolving `module:Class` entry points)."""
    import importlib
    import sys
    import types
    reg = {}
    gym = types.ModuleType("gym")
    gym_envs = types.ModuleType("gym.envs")
    gym_reg = types.ModuleType("gym.envs.registration")
    gym_reg.register = lambda id, entry_point, **kw: reg.__setitem__(id, entry_point)
    gym.envs, gym_envs.registration = gym_envs, gym_reg
    for name, mod in (("gym", gym), ("gym.envs", gym_envs), ("gym.envs.registration", gym_reg)):
        monkeypatch.setitem(sys.modules, name, mod)
    for name in [k for k in sys.modules if k == "bioimitation" or k.startswith("bioimitation.")]:
        monkeypatch.delitem(sys.modules, name)
    bio = importlib.import_module("bioimitation")
    assert bio.REGISTERED["gym"] and len(reg) == 17
    assert reg["MuscleWalkingImitation2D-v0"] == \
        "bioimitation.imitation_envs.envs.muscle.planar.muscle_walking_imitation_env2D:MuscleWalkingImitationEnv2D"
    assert reg["TorqueLockedKneeImitation3D-v0"] == \
        "bioimitation.imitation_envs.envs.torque.spatial.torque_locked_knee_imitation_env3D:TorqueLockedKneeImitationEnv3D"
    for env_id, ep in reg.items():                      # every entry point resolves the way gym resolves it
        mod, cls = ep.split(":")
        assert getattr(importlib.import_module(mod), cls).ENV_ID == env_id
    # the reference's own module paths exist for each registered class (reference bioimitation/__init__.py:11-20,75-86)
    from bioimitation.imitation_envs.envs.muscle.spatial.muscle_palsy_imitation_env3D import MusclePalsyImitationEnv3D
    assert MusclePalsyImitationEnv3D.ENV_ID == "MusclePalsyImitation3D-v0"
