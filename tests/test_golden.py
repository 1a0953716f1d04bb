"""The CPU oracle against the committed golden trajectories (tests/golden/*.npz,
written by tests/golden/make_golden.py).  Pins the oracle: any change of its
arithmetic shows up here before it can silently move the GPU parity target."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = ["config1_muscle_walking_2d.npz", "torque_walking_2d.npz", "muscle_walking_3d.npz",
         "muscle_locked_knee_3d.npz", "muscle_palsy_3d.npz", "torque_walking_3d.npz", "muscle_running_2d.npz",
         "muscle_locked_knee_2d.npz", "muscle_jumping_2d.npz", "muscle_jumping_3d.npz", "torque_running_3d.npz"]


def _env(env_id, seed):
    from bioimitation_gym_b200 import registry, tasks
    from oracle import oracle as orc
    cfg = tasks.merged_config(dict(num_envs=1, seed=seed))
    spec, cm, ref, task = registry.build_env_tables(env_id, cfg, None, None)
    rt = orc.RefTables(ref["q"], ref["u"], ref["body_pos"], ref["com_pos"])
    return orc.OracleVecEnv(cm.tables, task, rt, 1, seed=seed), cm, task


@pytest.mark.parametrize("fname", FILES)
def test_golden_files_are_complete(fname):
    g = np.load(os.path.join(GOLDEN, fname))
    n = g["action"].shape[0]
    for k in ("q", "u", "act", "lm", "obs", "reward", "terms", "done", "reason", "tendon_force", "udot", "contact"):
        assert g[k].shape[0] == n, k
    assert np.isfinite(g["obs"]).all() and np.isfinite(g["reward"]).all()
    assert g["contact"].shape[1:] == (2, 6)


@pytest.mark.parametrize("fname", FILES)
def test_oracle_reproduces_golden_trajectory(fname, oracle_lib):
    g = np.load(os.path.join(GOLDEN, fname))
    env, cm, task = _env(str(g["env_id"]), int(g["seed"]))
    obs0 = env.reset()
    np.testing.assert_allclose(obs0[0], g["reset_obs"], rtol=1e-9, atol=1e-12)
    steps = min(300, g["action"].shape[0])
    for k in range(steps):
        st = env.get_state()
        np.testing.assert_allclose(st["q"][0], g["q"][k], rtol=1e-9, atol=1e-11)
        assert st["istep"][0] == g["istep"][k] and st["episode"][0] == g["episode"][k]
        obs, rew, done, terms, reasons = env.step(g["action"][k][None])
        np.testing.assert_allclose(obs[0], g["obs"][k], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(rew[0], g["reward"][k], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(terms[0], g["terms"][k], rtol=1e-9, atol=1e-12)
        assert done[0] == g["done"][k] and reasons[0] == g["reason"][k]


def test_golden_evaluation_matches_state_based_evaluation(oracle_lib):
    """The per-step evaluation record (tendon force, udot, contact) equals an
    independent evaluation of the next pre-step state where no reset happened."""
    orc = oracle_lib
    g = np.load(os.path.join(GOLDEN, "config1_muscle_walking_2d.npz"))
    env, cm, task = _env(str(g["env_id"]), int(g["seed"]))
    checked = 0
    for k in range(0, 400, 7):
        if g["done"][k]:
            continue
        # the state before step k+1 is the state after step k; controls = mean action of step k
        hist = g["history"][k + 1]
        ctrl = np.clip(hist.mean(axis=0), 0.0, 1.0)
        ev = orc.eval_dynamics(cm.tables, g["q"][k + 1], g["u"][k + 1], g["act"][k + 1], g["lm"][k + 1], ctrl)
        np.testing.assert_allclose(ev["tendon_force"], g["tendon_force"][k], rtol=1e-7, atol=1e-7)
        np.testing.assert_allclose(ev["contact"], g["contact"][k], rtol=1e-7, atol=1e-6)
        checked += 1
    assert checked > 30
