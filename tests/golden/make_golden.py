#!/usr/bin/env python
"""Generate the golden trajectory fixtures (SURVEY 8d "Config 1").

The reference cannot run here (opensim==4.1 is not installable offline), so the
golden vectors are dumps of the CPU oracle (oracle/bio_oracle.c, fp64): they pin
the oracle against accidental change (tests/test_golden.py, CPU) and are the
fixed inputs/outputs the CUDA path is compared with on the GPU box
(tests/test_gpu_golden.py).  PARITY UNPINNED against OpenSim; see DESIGN.md 5.

  python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

Protocol per fixture: env `env_id`, 1 env, seed 0, default config, actions iid
U[lo,hi]^na from numpy default_rng(0) (the equivalent of action_space.sample(),
reference tests/test_env.py:24), `steps` control steps with auto-reset.  Per
step the dump holds the pre-step state, the action, obs / reward / reward terms /
done / done-reason, and the evaluation of the post-step state before any reset
(tendon forces, udot, contact wrenches).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FIXTURES = [
    # (file, env id, steps)
    ("config1_muscle_walking_2d.npz", "MuscleWalkingImitation2D-v0", 1000),
    ("torque_walking_2d.npz", "TorqueWalkingImitation2D-v0", 150),
    ("muscle_walking_3d.npz", "MuscleWalkingImitation3D-v0", 150),
    ("muscle_locked_knee_3d.npz", "MuscleLockedKneeImitation3D-v0", 100),
    ("muscle_palsy_3d.npz", "MusclePalsyImitation3D-v0", 100),
    ("torque_walking_3d.npz", "TorqueWalkingImitation3D-v0", 100),
    ("muscle_running_2d.npz", "MuscleRunningImitation2D-v0", 100),
    ("muscle_locked_knee_2d.npz", "MuscleLockedKneeImitation2D-v0", 100),
    ("muscle_jumping_2d.npz", "MuscleJumpingImitation2D-v0", 220),     # past the mirror point of the reference index
    ("muscle_jumping_3d.npz", "MuscleJumpingImitation3D-v0", 100),
    ("torque_running_3d.npz", "TorqueRunningImitation3D-v0", 100),
]
STATE_KEYS = ("q", "u", "act", "lm", "last_action", "history", "old_px", "istep", "first", "hist_pos", "episode")


def action_bounds(spec):
    return (-1.0, 1.0) if spec.torque else (0.0, 1.0)


def generate(env_id, steps, seed=0):
    from bioimitation_gym_b200 import registry, tasks
    from oracle import oracle as orc
    cfg = tasks.merged_config(dict(num_envs=1, seed=seed))
    spec, cm, ref, task = registry.build_env_tables(env_id, cfg, None, None)
    rt = orc.RefTables(ref["q"], ref["u"], ref["body_pos"], ref["com_pos"])
    env = orc.OracleVecEnv(cm.tables, task, rt, 1, seed=seed)
    shadow = orc.OracleVecEnv(cm.tables, task, rt, 1, seed=seed)
    t = cm.tables
    nd, nm, na = t.n_dof, t.n_muscles, t.n_act
    lo, hi = action_bounds(spec)
    rng = np.random.default_rng(0)
    out = {k: [] for k in STATE_KEYS}
    for k in ("action", "obs", "reward", "terms", "done", "reason", "tendon_force", "udot", "contact"):
        out[k] = []
    obs0 = env.reset()
    A = np.ctypeslib.as_array
    for _ in range(steps):
        st = env.get_state()
        for k in STATE_KEYS:
            out[k].append(st[k][0].copy())
        a = rng.uniform(lo, hi, (1, na))
        shadow.set_state(st)
        _, _, _, _, ev = shadow.step_env_debug(0, a[0])
        out["tendon_force"].append(A(ev.tendon_force)[:nm].copy())
        out["udot"].append(A(ev.udot)[:nd].copy())
        out["contact"].append(A(ev.contact).copy())
        obs, rew, done, terms, reasons = env.step(a)
        out["action"].append(a[0])
        out["obs"].append(obs[0])
        out["reward"].append(rew[0])
        out["terms"].append(terms[0])
        out["done"].append(done[0])
        out["reason"].append(reasons[0])
    res = {k: np.asarray(v) for k, v in out.items()}
    res["reset_obs"] = obs0[0]
    res["env_id"] = np.asarray(env_id)
    res["seed"] = np.asarray(seed)
    return res


def main():
    only = set(sys.argv[1:])              # file names to (re)generate; default: all
    for fname, env_id, steps in FIXTURES:
        if only and fname not in only:
            continue
        res = generate(env_id, steps)
        path = os.path.join(HERE, fname)
        np.savez_compressed(path, **res)
        print("%s: %d steps, %d episodes finished, %.0f kB" % (
            fname, steps, int(res["done"].sum()), os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
