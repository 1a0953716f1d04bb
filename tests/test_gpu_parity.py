"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the
same seeded inputs.  Run on the B200 box: python -m pytest tests -m gpu

Tolerances (stated, SURVEY 8c parity protocol):
  fp64 build   : relative 1e-6 per step with denominators max(|x|, scale),
                 scale = F_iso (muscle forces), 1 rad/s^2 (udot), body weight
                 (contact), 1 (obs, reward)
  fp32 build   : per-step (re-synchronised) relative 2e-3 on forces/udot,
                 1e-3 absolute on reward; measured values are printed
  trajectory   : 100 free-running steps, fp64 drift bound 1e-6 rad on q;
                 fp32 drift reported and bounded by 5e-2 rad while both sides
                 are still in the same episode
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mk(env_id, n, dtype, seed=11, **kw):
    import torch  # noqa: F401
    from bioimitation_gym_b200 import backend
    from oracle import oracle as orc
    cfg = dict(num_envs=n, dtype=dtype, seed=seed)
    cfg.update(kw)
    env = backend.VecEnv(env_id, cfg)
    ref = orc.RefTables(env.ref["q"], env.ref["u"], env.ref["body_pos"], env.ref["com_pos"])
    cpu = orc.OracleVecEnv(env.cm.tables, env.task, ref, n, seed=seed)
    return env, cpu


def _actions(env, rng, n):
    if env.spec.torque:
        return rng.uniform(-1.0, 1.0, (n, env.n_act))
    return rng.uniform(0.0, 1.0, (n, env.n_act))


def _np(t):
    return t.detach().cpu().double().numpy()


def _rel(a, b, scale):
    return np.abs(a - b) / np.maximum(np.abs(b), scale)


ENV_IDS_2D = ["MuscleWalkingImitation2D-v0", "TorqueWalkingImitation2D-v0"]
ENV_IDS_ALL = ENV_IDS_2D + ["MuscleWalkingImitation3D-v0", "MusclePalsyImitation3D-v0",
                            "MuscleLockedKneeImitation3D-v0", "TorqueWalkingImitation3D-v0",
                            "TorqueLockedKneeImitation2D-v0", "MuscleRunningImitation2D-v0",
                            "MuscleLockedKneeImitation2D-v0"]


@pytest.mark.parametrize("env_id", ENV_IDS_ALL)
def test_reset_matches_oracle_fp64(env_id):
    env, cpu = _mk(env_id, 96, "float64")
    og = _np(env.reset())
    oc = cpu.reset()
    sg = env.get_state()
    sc = cpu.get_state()
    assert (sg["istep"].cpu().numpy() == sc["istep"]).all()
    for k in ("q", "u", "act", "lm"):
        np.testing.assert_allclose(_np(sg[k]), sc[k], rtol=1e-9, atol=1e-10, err_msg=k)
    assert np.max(_rel(og, oc, 1.0)) < 1e-6
    assert len(set(sc["istep"].tolist())) > 5      # reference rows are actually sampled
    env.close()


@pytest.mark.parametrize("env_id", ENV_IDS_ALL)
def test_single_evaluation_fp64(env_id):
    """Same (state, controls): muscle forces, udot, contact wrench, M, bias."""
    import torch
    env, cpu = _mk(env_id, 64, "float64", auto_reset=False)
    rng = np.random.default_rng(5)
    env.reset()
    cpu.reset()
    from oracle import oracle as orc
    t = env.cm.tables
    worst = {}
    for k in range(6):           # walk a few steps so contact and limits become active
        a = _actions(env, rng, 64)
        env.step(torch.as_tensor(a, dtype=env.dtype, device=env.device))
        cpu.step(a)
        st = cpu.get_state()
        env.set_state({kk: st[kk] for kk in ("q", "u", "act", "lm")})
        ctrl = np.clip(a if not env.spec.torque else a * 50, np.ctypeslib.as_array(t.act_min)[:env.n_act],
                       np.ctypeslib.as_array(t.act_max)[:env.n_act])
        dbg = env.eval_debug(torch.as_tensor(ctrl, dtype=env.dtype, device=env.device))
        weight = t.total_mass * 9.80665
        for i in range(0, 64, 7):
            ev = orc.eval_dynamics(t, st["q"][i], st["u"][i], st["act"][i], st["lm"][i], ctrl[i],
                                   newton_iters=env.task.newton_iters)
            checks = [("udot", ev["udot"], 1.0), ("bias", ev["bias"], 1.0), ("contact", ev["contact"], weight),
                      ("limit_force", ev["limit_force"], 1.0), ("mass_matrix", ev["mass_matrix"], 1e-2)]
            if t.n_muscles:
                fiso = np.ctypeslib.as_array(t.mus_fiso)[:t.n_muscles]
                checks += [("tendon_force", ev["tendon_force"], fiso), ("fiber_force", ev["fiber_force"], fiso),
                           ("fiber_vel", ev["lmdot"], 1e-2), ("act_dot", ev["adot"], 1.0),
                           ("path_len", ev["path_len"], 1.0), ("path_vel", ev["path_vel"], 1e-2)]
            for name, want, scale in checks:
                got = _np(dbg[name][i])
                e = np.max(_rel(got, want, scale)) if want.size else 0.0
                worst[name] = max(worst.get(name, 0.0), e)
    print(env_id, {k: "%.2e" % v for k, v in worst.items()})
    for name, e in worst.items():
        assert e < 1e-6, (name, e)
    env.close()


@pytest.mark.parametrize("env_id", ENV_IDS_ALL)
def test_step_parity_fp64_100_steps(env_id):
    """Free-running 100 control steps with auto-reset: obs, reward, terms, done
    and the integrated state agree at every step.  Rounding-level differences
    (1e-11 on udot, see test_single_evaluation_fp64) grow through the stiff
    contact dynamics, so the stated 100-step drift bound is 1e-4."""
    import torch
    n = 48
    env, cpu = _mk(env_id, n, "float64")
    rng = np.random.default_rng(9)
    env.reset()
    cpu.reset()
    worst_obs = worst_rew = worst_q = worst_terms = 0.0
    n_done = 0
    for k in range(100):
        a = _actions(env, rng, n)
        if k == 17:
            a[3, :] = np.nan                       # NaN action -> zeros (opensim_wrapper.py:93-95)
        obs, rew, done, info = env.step(torch.as_tensor(a, dtype=env.dtype, device=env.device))
        oc, rc, dc, tc, reasons = cpu.step(a)
        assert (done.cpu().numpy() == dc).all(), "done mismatch at step %d" % k
        n_done += int(dc.sum())
        # accelerations / contact forces are O(1e3): relative with scale 1e2
        worst_obs = max(worst_obs, np.max(_rel(_np(obs), oc, 100.0)))
        worst_rew = max(worst_rew, np.max(np.abs(_np(rew) - rc)))
        worst_terms = max(worst_terms, np.max(np.abs(_np(info["all_rewards"]) - tc)))
        sg, sc = env.get_state(), cpu.get_state()
        worst_q = max(worst_q, np.max(np.abs(_np(sg["q"]) - sc["q"])))
        assert (sg["istep"].cpu().numpy() == sc["istep"]).all()
    print(env_id, "fp64 100 steps: obs %.2e reward %.2e terms %.2e q-drift %.2e episodes finished %d"
          % (worst_obs, worst_rew, worst_terms, worst_q, n_done))
    assert worst_obs < 1e-4 and worst_rew < 1e-4 and worst_q < 1e-4 and worst_terms < 1e-4
    assert n_done > 0                              # auto-reset path exercised
    stats = env.stats().cpu().numpy()
    assert stats[0] == 100 * n and stats[1] == n_done and stats[10] == 1
    env.close()


@pytest.mark.parametrize("env_id", ENV_IDS_2D + ["MuscleWalkingImitation3D-v0"])
def test_step_parity_fp32_resynchronised(env_id):
    """fp32 production build, one control step from identical states (the fp64
    oracle state is copied to the GPU before every step): stated per-step
    tolerances 2e-4 rad on q, 2e-5 m on fibre length, 2e-3 on the observation
    (relative, scale 100 for the O(1e3) accelerations), 2e-3 on the reward."""
    import torch
    n = 64
    env, cpu = _mk(env_id, n, "float32")
    rng = np.random.default_rng(21)
    env.reset()
    cpu.reset()
    worst = dict(obs=0.0, rew=0.0, q=0.0, lm=0.0)
    agree = total = 0
    t = env.cm.tables
    n_pel = sum(1 for i in range(t.n_coords) if t.coord_pelvis_trans[i] != 0)
    acc0 = 1 + (t.n_coords - n_pel) + t.n_coords      # slice of coordinate_acc in the observation
    acc1 = acc0 + t.n_coords
    for k in range(40):
        st = cpu.get_state()
        env.set_state(st)
        a = _actions(env, rng, n).astype(np.float32).astype(np.float64)
        obs, rew, done, info = env.step(torch.as_tensor(a, dtype=env.dtype, device=env.device))
        oc, rc, dc, tc, reasons = cpu.step(a)
        dg = done.cpu().numpy()
        same = (dg == dc)
        agree += int(same.sum())
        total += n
        live = same & (dc == 0)
        sg, sc = env.get_state(), cpu.get_state()
        if live.any():
            worst["q"] = max(worst["q"], np.max(np.abs(_np(sg["q"]) - sc["q"])[live]))
            if env.n_muscles:
                worst["lm"] = max(worst["lm"], np.max(np.abs(_np(sg["lm"]) - sc["lm"])[live]))
            rel = _rel(_np(obs), oc, 100.0)[live]
            worst["acc"] = max(worst.get("acc", 0.0), np.max(rel[:, acc0:acc1]))
            rel[:, acc0:acc1] = 0.0
            worst["obs"] = max(worst["obs"], np.max(rel))
            worst["rew"] = max(worst["rew"], np.max(np.abs(_np(rew) - rc)[live]))
    print(env_id, "fp32 re-synchronised per-step errors:", {k: "%.2e" % v for k, v in worst.items()},
          "done agreement %d/%d" % (agree, total))
    assert worst["q"] < 2e-4 and worst["lm"] < 2e-5
    # coordinate_acc is the solution of an ill-conditioned 9..14-dof solve (foot vs trunk inertia)
    assert worst["obs"] < 2e-3 and worst["acc"] < 3e-2 and worst["rew"] < 2e-3
    assert agree >= 0.995 * total
    env.close()


def test_fp32_trajectory_drift_bound():
    """100 free-running steps in fp32 vs the fp64 oracle: drift while both
    sides are still in their first episode."""
    import torch
    n = 64
    env, cpu = _mk("MuscleWalkingImitation2D-v0", n, "float32", auto_reset=False)
    rng = np.random.default_rng(33)
    env.reset()
    cpu.reset()
    alive = np.ones(n, dtype=bool)
    drift = 0.0
    steps_alive = 0
    for k in range(100):
        a = _actions(env, rng, n).astype(np.float32).astype(np.float64)
        obs, rew, done, info = env.step(torch.as_tensor(a, dtype=env.dtype, device=env.device))
        oc, rc, dc, tc, _ = cpu.step(a)
        alive &= (dc == 0) & (done.cpu().numpy() == 0)
        if not alive.any():
            break
        sg, sc = env.get_state(), cpu.get_state()
        drift = max(drift, np.max(np.abs(_np(sg["q"]) - sc["q"])[alive]))
        steps_alive = k + 1
    print("fp32 free-running drift over %d steps: %.3e rad" % (steps_alive, drift))
    assert steps_alive >= 10
    assert drift < 5e-2
    env.close()


def test_get_set_state_round_trip_and_determinism():
    import torch
    env, _ = _mk("MuscleWalkingImitation2D-v0", 128, "float32")
    env.reset()
    g = torch.Generator(device="cpu").manual_seed(0)
    acts = [torch.rand((128, 14), generator=g) for _ in range(5)]
    for a in acts[:2]:
        env.step(a)
    snap = {k: v.clone() for k, v in env.get_state().items()}
    outs = []
    for rep in range(2):
        env.set_state(snap)
        for a in acts[2:]:
            obs, rew, done, _ = env.step(a)
        outs.append((obs.clone(), rew.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    env.close()


def test_sharding_invariance():
    """Counter-based RNG keyed by the global env index: 2 shards of 64 envs
    equal one batch of 128 (what makes multi-GPU results independent of G)."""
    import torch
    from bioimitation_gym_b200 import backend
    full = backend.VecEnv("MuscleWalkingImitation2D-v0", dict(num_envs=128, seed=4))
    lo = backend.VecEnv("MuscleWalkingImitation2D-v0", dict(num_envs=64, seed=4, env_offset=0))
    hi = backend.VecEnv("MuscleWalkingImitation2D-v0", dict(num_envs=64, seed=4, env_offset=64))
    of = full.reset().clone()
    assert torch.equal(of[:64], lo.reset()) and torch.equal(of[64:], hi.reset())
    g = torch.Generator(device="cpu").manual_seed(1)
    for _ in range(40):
        a = torch.rand((128, 14), generator=g)
        o, r, d, _ = full.step(a)
        o1, r1, d1, _ = lo.step(a[:64])
        o2, r2, d2, _ = hi.step(a[64:])
        assert torch.equal(o[:64], o1) and torch.equal(o[64:], o2)
        assert torch.equal(d[:64], d1) and torch.equal(d[64:], d2)
    for e in (full, lo, hi):
        e.close()


@pytest.mark.parametrize("env_id,n_act", [("MuscleWalkingImitation2D-v0", 14), ("MuscleWalkingImitation3D-v0", 22)])
def test_persistent_launch_walks_every_env(env_id, n_act):
    """The step kernel is a persistent launch (one CTA per SM, warps walk the env items):
    a batch large enough that every warp takes several items, with many auto-resets hitting
    only one of the two envs of a warp, must equal the same envs stepped in small shards
    that need one item per warp."""
    import torch
    from bioimitation_gym_b200 import backend
    n, shard = 12000, 1500
    full = backend.VecEnv(env_id, dict(num_envs=n, seed=9))
    parts = [backend.VecEnv(env_id, dict(num_envs=shard, seed=9, env_offset=k)) for k in range(0, n, shard)]
    of = full.reset().clone()
    for k, pe in enumerate(parts):
        assert torch.equal(of[k * shard:(k + 1) * shard], pe.reset())
    g = torch.Generator(device="cpu").manual_seed(2)
    n_done = 0
    for _ in range(120):
        a = torch.rand((n, n_act), generator=g)
        o, r, d, _ = full.step(a)
        n_done += int(d.sum())
        for k, pe in enumerate(parts):
            sl = slice(k * shard, (k + 1) * shard)
            o1, r1, d1, _ = pe.step(a[sl])
            assert torch.equal(o[sl], o1) and torch.equal(r[sl], r1) and torch.equal(d[sl], d1)
    assert n_done > 20           # auto-reset was exercised
    for e in [full] + parts:
        e.close()


def test_host_buffer_entry_point_matches_device_path():
    import torch
    env, _ = _mk("MuscleWalkingImitation2D-v0", 256, "float32")
    env2, _ = _mk("MuscleWalkingImitation2D-v0", 256, "float32")
    env.reset()
    obs_h = np.zeros((256, env.obs_dim), dtype=np.float32)
    env2.reset_host(obs_h)
    rew_h = np.zeros(256, dtype=np.float32)
    done_h = np.zeros(256, dtype=np.uint8)
    terms_h = np.zeros((256, 5), dtype=np.float32)
    rng = np.random.default_rng(2)
    for _ in range(5):
        a = rng.uniform(0, 1, (256, 14)).astype(np.float32)
        o, r, d, info = env.step(torch.as_tensor(a))
        env2.step_host(a, obs_h, rew_h, done_h, terms_h)
        assert np.array_equal(o.cpu().numpy(), obs_h) and np.array_equal(r.cpu().numpy(), rew_h)
        assert np.array_equal(d.cpu().numpy(), done_h)
    env.close()
    env2.close()


def test_host_entry_point_with_page_locked_buffers_matches_device_path():
    """Page-locked host buffers are read / written by the kernel in place (no staging copies): same results
    as the device path, with a mix of pinned and pageable buffers falling back to the staged copies."""
    import torch
    n = 301                                    # odd tail: the last warp holds one env
    env, _ = _mk("MuscleWalkingImitation2D-v0", n, "float32")
    env2, _ = _mk("MuscleWalkingImitation2D-v0", n, "float32")
    env3, _ = _mk("MuscleWalkingImitation2D-v0", n, "float32")
    env.reset(); env2.reset(); env3.reset()
    pin = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt).pin_memory()
    a_p, o_p, r_p, d_p, t_p = pin(n, 14), pin(n, env.obs_dim), pin(n), pin(n, dt=torch.uint8), pin(n, 5)
    o3 = np.zeros((n, env.obs_dim), dtype=np.float32)      # pageable observation buffer, pinned others
    rng = np.random.default_rng(5)
    n_done = 0
    for _ in range(150):
        a = rng.uniform(0, 1, (n, 14)).astype(np.float32)
        a_p.copy_(torch.as_tensor(a))
        o, r, d, info = env.step(torch.as_tensor(a))
        env2.step_host(a_p.numpy(), o_p.numpy(), r_p.numpy(), d_p.numpy(), t_p.numpy())
        assert np.array_equal(o.cpu().numpy(), o_p.numpy()) and np.array_equal(r.cpu().numpy(), r_p.numpy())
        assert np.array_equal(d.cpu().numpy(), d_p.numpy())
        assert np.array_equal(info["all_rewards"].cpu().numpy(), t_p.numpy())
        r3, d3, t3 = pin(n), pin(n, dt=torch.uint8), pin(n, 5)
        env3.step_host(a_p.numpy(), o3, r3.numpy(), d3.numpy(), t3.numpy())
        assert np.array_equal(o3, o_p.numpy()) and np.array_equal(r3.numpy(), r_p.numpy())
        n_done += int(d_p.sum())
    assert n_done > 0
    for e in (env, env2, env3):
        e.close()


@pytest.mark.parametrize("env_id,n_act", [("MuscleWalkingImitation2D-v0", 14), ("MuscleWalkingImitation3D-v0", 22),
                                          ("MuscleLockedKneeImitation3D-v0", 19)])
def test_compiled_and_streamed_muscle_paths_give_the_same_trajectory(env_id, n_act, monkeypatch):
    """The kernels evaluate the muscle geometry from compiled paths (constant length + live segments per variant
    of the conditional points); BIO_PLANAR_STREAM_PATHS=1 at create time makes them stream over the path points
    instead (the fallback any model may need).  fp64, 30 control steps with auto-reset: same observations."""
    import torch
    n = 96
    monkeypatch.setenv("BIO_PLANAR_STREAM_PATHS", "0")
    a_env, _ = _mk(env_id, n, "float64")
    monkeypatch.setenv("BIO_PLANAR_STREAM_PATHS", "1")
    b_env, _ = _mk(env_id, n, "float64")
    monkeypatch.setenv("BIO_PLANAR_STREAM_PATHS", "0")
    oa, ob = a_env.reset().clone(), b_env.reset().clone()
    assert torch.equal(oa, ob)
    g = torch.Generator().manual_seed(4)
    worst = 0.0
    for _ in range(30):
        a = torch.rand((n, n_act), generator=g, dtype=torch.float64)
        oa, ra, da, _ = a_env.step(a)
        ob, rb, db, _ = b_env.step(a)
        assert torch.equal(da, db)
        worst = max(worst, float(((oa - ob).abs() / oa.abs().clamp(min=1.0)).max()), float((ra - rb).abs().max()))
    print(env_id, "compiled vs streamed paths, worst difference %.2e" % worst)
    assert worst < 1e-6                      # free-running contact dynamics amplify the 1e-16 rounding differences
    a_env.close()
    b_env.close()


def test_numpy_api_matches_tensor_api():
    """step_np / reset_np (page-locked double buffers, what the gym classes and the RLlib adapter call) return
    the same numbers as the tensor API; the arrays of a call stay valid through the next call."""
    import torch
    n = 64
    env, _ = _mk("MuscleWalkingImitation2D-v0", n, "float32")
    env2, _ = _mk("MuscleWalkingImitation2D-v0", n, "float32")
    o0 = env.reset().cpu().numpy()
    assert np.array_equal(o0, env2.reset_np())
    rng = np.random.default_rng(9)
    prev = None
    for _ in range(6):
        a = rng.uniform(0, 1, (n, 14)).astype(np.float32)
        o, r, d, info = env.step(torch.as_tensor(a))
        o2, r2, d2, info2 = env2.step_np(a)
        assert np.array_equal(o.cpu().numpy(), o2) and np.array_equal(r.cpu().numpy(), r2)
        assert np.array_equal(d.cpu().numpy().astype(bool), d2)
        assert np.array_equal(info["all_rewards"].cpu().numpy(), info2["all_rewards"])
        if prev is not None:
            assert np.array_equal(prev[0], prev[1])      # the previous call's array was not overwritten
        prev = (o2, o2.copy())
    env.close()
    env2.close()


def test_errors_are_reported_not_thrown():
    import torch
    from bioimitation_gym_b200 import backend
    env, _ = _mk("MuscleWalkingImitation2D-v0", 8, "float32")
    with pytest.raises(ValueError):
        env.step(torch.zeros((8, 3)))
    with pytest.raises(backend.BioError):
        backend.VecEnv("MuscleWalkingImitation2D-v0", dict(num_envs=0))
    env.close()


def test_single_env_classes_keep_the_reference_call_pattern():
    """tests/test_env.py:15-26 of the reference: make, reset, step with sampled actions."""
    from bioimitation_gym_b200 import envs
    env = envs.make("MuscleWalkingImitation2D-v0", dict(mode="train"))
    assert type(env).__name__ == "MuscleWalkingImitationEnv2D"
    assert env.action_space.shape == (14,) and env.observation_space.shape == (138,)
    assert float(env.action_space.low.min()) == 0.0 and float(env.action_space.high.max()) == 1.0
    obs = env.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (138,)
    for _ in range(30):
        obs, reward, done, info = env.step(env.action_space.sample())
        assert obs.shape == (138,) and isinstance(reward, float) and isinstance(done, bool)
        assert len(info["all_rewards"]) == 5
        if done:
            obs = env.reset()
    d = env.reset(obs_as_dict=True)
    assert list(d)[:4] == ["phase", "coordinate_pos", "coordinate_vel", "coordinate_acc"]
    assert "pelvis_tx" not in d["coordinate_pos"] and len(d["muscles"]) == 14
    env.close()
    tq = envs.make("TorqueWalkingImitation3D-v0", dict(max_actuation=150))
    assert tq.action_space.shape == (11,) and float(tq.action_space.high[0]) == 200.0   # model compiled with 200
    tq.close()


def test_large_batch_and_all_integrators_run():
    import torch
    from bioimitation_gym_b200 import backend
    for integ, sub in (("rk2", 20), ("rk4", 10), ("semi_implicit_euler", 40), ("implicit_damping", 10)):
        env = backend.VecEnv("MuscleWalkingImitation2D-v0", dict(num_envs=1000, integrator=integ, substeps=sub))
        env.reset()
        for _ in range(3):
            obs, rew, done, _ = env.step(torch.rand((1000, 14), device=env.device))
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
        env.close()
    env = backend.VecEnv("MusclePalsyImitation3D-v0", dict(num_envs=3000, apply_perturbations=True))
    env.reset()
    for _ in range(5):
        obs, rew, done, _ = env.step(torch.rand((3000, 22), device=env.device))
    assert torch.isfinite(obs).all()
    env.close()


def test_trajectory_recorder_writes_state_storage(tmp_path):
    """save_simulation (opensim_wrapper.py:334-338): the recorded states of one env as an
    OpenSim states storage that reads back to what get_state returned."""
    import torch
    from bioimitation_gym_b200 import backend, refmotion
    from bioimitation_gym_b200.rollout import TrajectoryRecorder
    env = backend.VecEnv("MuscleWalkingImitation2D-v0", dict(num_envs=8, dtype="float64", seed=2, auto_reset=False))
    rec = TrajectoryRecorder(env, [3])
    env.reset()
    rec.record()
    g = torch.Generator().manual_seed(0)
    for _ in range(5):
        env.step(torch.rand((8, 14), generator=g, dtype=torch.float64))
        rec.record()
    path = rec.save_simulation(str(tmp_path))
    labels, data, _ = refmotion.read_storage(path)
    assert data.shape == (6, 1 + 2 * env.n_dof + 2 * env.n_muscles)
    st = env.get_state()
    assert np.allclose(data[-1, 1:1 + 2 * env.n_dof:2], st["q"][3].cpu().numpy(), atol=1e-7)
    assert np.allclose(np.diff(data[:, 0]), 0.01, atol=1e-9)
    env.close()


@pytest.mark.parametrize("env_id", ["MuscleWalkingImitation2D-v0", "MuscleWalkingImitation3D-v0"])
def test_perturbation_force_parity_fp64(env_id):
    """apply_perturbations=True (piecewise-constant +-50 N on the torso, muscle_walking_imitation_env2D.py:83-100):
    the force sequence is keyed by (seed, env, knot) on both sides; 150 free-running steps in fp64."""
    import torch
    n = 32
    env, cpu = _mk(env_id, n, "float64", seed=5, apply_perturbations=True)
    assert env.task.perturb == 1
    rng = np.random.default_rng(3)
    env.reset()
    cpu.reset()
    worst = 0.0
    for k in range(150):
        a = _actions(env, rng, n)
        obs, rew, done, info = env.step(torch.as_tensor(a, dtype=env.dtype, device=env.device))
        oc, rc, dc, tc, _ = cpu.step(a)
        assert (done.cpu().numpy() == dc).all(), "done mismatch at step %d" % k
        worst = max(worst, np.max(_rel(_np(obs), oc, 100.0)), np.max(np.abs(_np(rew) - rc)))
    print(env_id, "perturbed, fp64 150 steps: worst obs/reward difference %.2e" % worst)
    assert worst < 1e-4
    env.close()


def test_statistical_parity_fp32_batch():
    """over a batch, the fp32 production build and the fp64 oracle end episodes for the
    same reasons at the same rate: episodes finished, mean episode length and done-reason histogram of
    2048 envs x 160 steps (iid uniform excitations, auto-reset) within 2 % of each other."""
    import os
    import torch
    n, steps = 2048, 160
    env, _ = _mk("MuscleWalkingImitation2D-v0", n, "float32", seed=21)
    from oracle import oracle as orc
    ref = orc.RefTables(env.ref["q"], env.ref["u"], env.ref["body_pos"], env.ref["com_pos"])
    cpu = orc.OracleVecEnv(env.cm.tables, env.task, ref, n, seed=21, threads=min(16, os.cpu_count() or 1))
    rng = np.random.default_rng(8)
    env.reset()
    cpu.reset()
    env.stats(reset=True)
    hist = np.zeros(8)
    ep_cpu = 0
    len_sum_cpu = 0
    ep_len = np.zeros(n, dtype=np.int64)
    for k in range(steps):
        a = rng.uniform(0, 1, (n, 14))
        env.step(torch.as_tensor(a, dtype=env.dtype, device=env.device))
        _, _, dc, _, reasons = cpu.step(a)
        ep_len += 1
        fin = dc.astype(bool)
        ep_cpu += int(fin.sum())
        len_sum_cpu += int(ep_len[fin].sum())
        ep_len[fin] = 0
        for bit in range(6):
            hist[bit] += int(((reasons >> bit) & 1)[fin].sum())
    st = env.stats().cpu().numpy()
    ep_gpu, len_gpu = st[1], st[3] / max(st[1], 1)
    len_cpu = len_sum_cpu / max(ep_cpu, 1)
    print("episodes gpu %d cpu %d, mean length gpu %.2f cpu %.2f, done reasons gpu %s cpu %s"
          % (ep_gpu, ep_cpu, len_gpu, len_cpu, st[4:10].astype(int).tolist(), hist[:6].astype(int).tolist()))
    assert ep_cpu > 1000
    assert abs(ep_gpu - ep_cpu) <= 0.02 * ep_cpu
    assert abs(len_gpu - len_cpu) <= 0.02 * len_cpu
    assert np.all(np.abs(st[4:10] - hist[:6]) <= 0.02 * ep_cpu + 2)
    env.close()
