"""The CUDA path (through the C ABI) against the committed golden trajectories
(tests/golden/*.npz: fp64 oracle dumps, SURVEY 8d Config 1).  No oracle code runs
here: the fixtures are the reference.

Tolerances (stated):
  fp64 free-running trajectory (same seed, same actions, auto-reset):
      obs relative 1e-4 (scale 100 for accelerations / forces), reward 1e-4, q 1e-4
  one control step from each golden pre-step state (all steps batched in one launch):
      fp64: obs 1e-6 relative (scale 100), reward / terms 1e-6, udot / tendon force /
            contact wrench of the post-step state 1e-6 relative
      fp32: obs 1e-2 relative (scale 100: 1 rad/s^2 on an acceleration; measured 2.9e-3 in 2D, 7.2e-3 in 3D), reward 2e-3, q 2e-4 rad
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = ["config1_muscle_walking_2d.npz", "torque_walking_2d.npz", "muscle_walking_3d.npz",
         "muscle_locked_knee_3d.npz"]
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


@pytest.mark.parametrize("dtype,tol_obs,tol_rew,tol_q", [("float64", 1e-6, 1e-6, 1e-8), ("float32", 1e-2, 2e-3, 2e-4)])
@pytest.mark.parametrize("fname", FILES)
def test_one_step_from_every_golden_state(fname, dtype, tol_obs, tol_rew, tol_q):
    """Env i of the batch is loaded with the golden pre-step state of step i; one launch
    steps them all.  Rows that finish an episode are compared on reward / done only (their
    observation is that of a reset keyed by the env index)."""
    import torch
    from bioimitation_gym_b200 import backend
    g = np.load(os.path.join(GOLDEN, fname))
    n = g["action"].shape[0]
    env = backend.VecEnv(str(g["env_id"]), dict(num_envs=n, dtype=dtype, seed=int(g["seed"])))
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
    print(fname, dtype, "one step from golden states: obs %.2e reward %.2e terms %.2e q %.2e" % (e_obs, e_rew, e_terms, e_q))
    assert e_obs < tol_obs and e_rew < tol_rew and e_terms < tol_rew and e_q < tol_q
    if dtype == "float64" and not env.spec.torque:
        # post-step evaluation record: same state, controls = clipped mean of the action history
        ctrl = np.clip(_np(st["history"]).mean(axis=1), 0.0, 1.0)
        ev = env.eval_debug(torch.as_tensor(ctrl, dtype=env.dtype, device=env.device))
        fiso = np.ctypeslib.as_array(env.cm.tables.mus_fiso)[:env.n_muscles]
        weight = abs(env.cm.tables.total_mass * env.cm.tables.gravity[1])
        e_f = np.max(np.abs(_np(ev["tendon_force"])[live] - g["tendon_force"][live]) / fiso)
        e_c = np.max(np.abs(_np(ev["contact"])[live] - g["contact"][live])) / weight
        e_u = np.max(_rel(_np(ev["udot"])[live], g["udot"][live], 1.0))
        print(fname, "post-step evaluation: tendon %.2e contact %.2e udot %.2e" % (e_f, e_c, e_u))
        assert e_f < 1e-6 and e_c < 1e-6 and e_u < 1e-6
    env.close()
