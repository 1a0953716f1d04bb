"""The CUDA path (through the C ABI) against the committed golden trajectories
(tests/golden/*.npz: fp64 oracle dumps, SURVEY 8d Config 1).  No oracle code runs
here: the fixtures are the reference.

Tolerances (stated):
  fp64 free-running trajectory (same seed, same actions, auto-reset):
      obs relative 1e-4 (scale 100 for accelerations / forces), reward 1e-4, q 1e-4
  one control step from each golden pre-step state (all steps batched in one launch):
      fp64: obs 1e-6 relative (scale 100), reward / terms 1e-6, udot / tendon force /
            contact wrench of the post-step state 1e-6 relative
      fp32: GOLDEN_FP32_TOL (2 x the worst value measured on the B200 over all fixtures and both launch shapes)
The post-step record (tendon forces, accelerations, contact wrenches) is read from the COOPERATIVE STEP KERNEL
itself (BioStepExtra), in both precisions and both fp32 launch shapes.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = ["config1_muscle_walking_2d.npz", "torque_walking_2d.npz", "muscle_walking_3d.npz",
         "muscle_locked_knee_3d.npz", "muscle_palsy_3d.npz", "torque_walking_3d.npz", "muscle_running_2d.npz",
         "muscle_locked_knee_2d.npz", "muscle_jumping_2d.npz", "muscle_jumping_3d.npz", "torque_running_3d.npz"]
STATE_KEYS = ("q", "u", "act", "lm", "last_action", "history", "old_px", "istep", "first", "hist_pos", "episode")


def _np(t):
    return t.detach().cpu().double().numpy()


def _rel(a, b, scale):
    return np.abs(a - b) / np.maximum(np.abs(b), scale)


@pytest.mark.parametrize("fname", FILES)
def test_free_running_trajectory_fp64(fname):
    import torch
    from bioimitation_gym_b200 import backend
    g = np.load(os.path.join(GOLDEN, fname))
    env = backend.VecEnv(str(g["env_id"]), dict(num_envs=1, dtype="float64", seed=int(g["seed"])))
    obs0 = _np(env.reset())
    assert np.max(_rel(obs0[0], g["reset_obs"], 1.0)) < 1e-9
    worst = dict(obs=0.0, rew=0.0, q=0.0)
    for k in range(g["action"].shape[0]):
        st = env.get_state()
        worst["q"] = max(worst["q"], np.max(np.abs(_np(st["q"])[0] - g["q"][k])))
        assert int(st["istep"][0]) == g["istep"][k] and int(st["episode"][0]) == g["episode"][k]
        obs, rew, done, info = env.step(torch.as_tensor(g["action"][k][None], dtype=env.dtype, device=env.device))
        assert int(done[0]) == g["done"][k], "done mismatch at step %d" % k
        worst["obs"] = max(worst["obs"], np.max(_rel(_np(obs)[0], g["obs"][k], 100.0)))
        worst["rew"] = max(worst["rew"], abs(float(rew[0]) - g["reward"][k]))
    print(fname, "fp64 free-running vs golden:", worst)
    assert worst["obs"] < 1e-4 and worst["rew"] < 1e-4 and worst["q"] < 1e-4
    env.close()


# fp32, one control step from the golden pre-step states: obs relative with the scale floor 100 (1 rad/s^2 on an
# acceleration), reward / terms absolute, q in rad; the step kernel's own read-outs: tendon force / F_iso, contact /
# body weight, udot relative with the scale floor 100.
# 2 x the worst value measured over the 11 fixtures and both launch shapes: obs 6.8e-3 (its acceleration entries),
# reward / terms 3.0e-6, q 2.4e-6, tendon 9.3e-5, contact 9.2e-5, udot 6.8e-3.
GOLDEN_FP32_TOL = dict(obs=1.4e-2, rew=6e-6, q=5e-6, tendon=2e-4, contact=2e-4, udot=1.4e-2)


@pytest.mark.parametrize("dtype,threads", [("float64", ""), ("float32", "lo"), ("float32", "hi")])
@pytest.mark.parametrize("fname", FILES)
def test_one_step_from_every_golden_state(fname, dtype, threads, monkeypatch):
    """Env i of the batch is loaded with the golden pre-step state of step i; one launch
    steps them all.  Rows that finish an episode are compared on reward / done only (their
    observation is that of a reset keyed by the env index)."""
    import torch
    from bioimitation_gym_b200 import backend
    g = np.load(os.path.join(GOLDEN, fname))
    n = g["action"].shape[0]
    if threads:
        monkeypatch.setenv("BIO_COOP_THREADS", threads)
    env = backend.VecEnv(str(g["env_id"]), dict(num_envs=n, dtype=dtype, seed=int(g["seed"])))
    if threads:
        monkeypatch.delenv("BIO_COOP_THREADS")
        assert env.coop_shape()[1] == {"lo": 512, "hi": 896 if env.spec.spatial else 640}[threads]
    ex = env.enable_step_extra("udot", "tendon_force", "contact")
    env.set_state({k: g[k] for k in STATE_KEYS})
    obs, rew, done, info = env.step(torch.as_tensor(g["action"], dtype=env.dtype, device=env.device))
    live = g["done"] == 0
    assert (done.cpu().numpy() == g["done"]).all()
    e_obs = np.max(_rel(_np(obs)[live], g["obs"][live], 100.0))
    e_rew = np.max(np.abs(_np(rew) - g["reward"]))
    e_terms = np.max(np.abs(_np(info["all_rewards"]) - g["terms"]))
    st = env.get_state()
    nxt = np.flatnonzero(live[:-1])
    e_q = np.max(np.abs(_np(st["q"])[nxt] - g["q"][nxt + 1]))
    # post-step evaluation record of the step kernel (taken before any reset, so finished rows count too)
    weight = abs(env.cm.tables.total_mass * env.cm.tables.gravity[1])
    fin = g["reason"] != 32                      # rows that ended non-finite hold no meaningful forces
    e_c = np.max(np.abs(_np(ex["contact"])[fin] - g["contact"][fin])) / weight
    e_u = np.max(_rel(_np(ex["udot"])[fin], g["udot"][fin], 1.0 if dtype == "float64" else 100.0))
    e_f = 0.0
    if env.n_muscles:
        fiso = np.ctypeslib.as_array(env.cm.tables.mus_fiso)[:env.n_muscles]
        e_f = np.max(np.abs(_np(ex["tendon_force"])[fin] - g["tendon_force"][fin]) / fiso)
    print(fname, dtype, threads, "one step from golden states: obs %.2e reward %.2e terms %.2e q %.2e | step-kernel "
          "read-outs: tendon %.2e contact %.2e udot %.2e" % (e_obs, e_rew, e_terms, e_q, e_f, e_c, e_u))
    if dtype == "float64":
        assert e_obs < 1e-6 and e_rew < 1e-6 and e_terms < 1e-6 and e_q < 1e-8
        assert e_f < 1e-6 and e_c < 1e-6 and e_u < 1e-6
    else:
        t = GOLDEN_FP32_TOL
        assert e_obs < t["obs"] and e_rew < t["rew"] and e_terms < t["rew"] and e_q < t["q"]
        assert e_f < t["tendon"] and e_c < t["contact"] and e_u < t["udot"]
    env.close()
