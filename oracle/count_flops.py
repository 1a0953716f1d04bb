#!/usr/bin/env python
"""Measure the algorithmic floating-point work of one env step with the
op-counting build of the oracle (count_flops.cpp) and write
profiles/flop_count.json.

Counting rule: every + - x / sqrt and every transcendental
(sin, cos, exp) counts as ONE operation; comparisons, abs, floor are listed
but not included in `flops`.  The number is the mean over a rollout with the
benchmark's action distribution, auto-reset included, so Newton iteration
counts, active contacts and resets are weighted as they occur.

  python oracle/count_flops.py            # test infrastructure, CPU only
"""
from __future__ import annotations

import ctypes
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != HERE]
sys.path.insert(0, ROOT)
LIB = os.path.join(HERE, "_build", "libbio_oracle_count.so")

ENVS = ["MuscleWalkingImitation2D-v0", "TorqueWalkingImitation2D-v0", "MuscleRunningImitation2D-v0",
        "MuscleJumpingImitation2D-v0", "MuscleLockedKneeImitation2D-v0", "MuscleWalkingImitation3D-v0",
        "MusclePalsyImitation3D-v0", "MuscleLockedKneeImitation3D-v0", "TorqueWalkingImitation3D-v0"]


def build():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["g++", "-O1", "-fPIC", "-std=c++17", "-shared", "-fpermissive", "-w", "-o", LIB,
                           os.path.join(HERE, "count_flops.cpp"), "-lm"])


def main():
    build()
    os.environ["BIO_ORACLE_LIB"] = LIB
    from bioimitation_gym_b200 import registry, tasks
    from oracle import oracle as orc
    L = orc.lib()
    out = {}
    for env_id in ENVS:
        spec, cm, ref, task = registry.build_env_tables(env_id, {})
        rt = orc.RefTables(ref["q"], ref["u"], ref["body_pos"], ref["com_pos"])
        n, steps = 16, 60
        env = orc.OracleVecEnv(cm.tables, task, rt, n, seed=0)
        env.reset()
        rng = np.random.default_rng(0)
        lo, hi = (-1.0, 1.0) if spec.torque else (0.0, 1.0)
        L.orc_count_reset()
        resets = 0
        for _ in range(steps):
            o, r, d, t, reasons = env.step(rng.uniform(lo, hi, (n, cm.tables.n_act)))
            resets += int(d.sum())
        ops = (ctypes.c_longlong * 6)()
        L.orc_count_get(ops)
        ops = [int(x) for x in ops]
        per = n * steps
        flops = sum(ops[:5]) / per
        rhs_per_sub = {0: 1, 1: 2, 2: 4, 3: 1}[task.integrator]
        out[env_id] = dict(
            flops_per_env_step=flops, add=ops[0] / per, mul=ops[1] / per, div=ops[2] / per, sqrt=ops[3] / per,
            transcendental=ops[4] / per, compare_abs_floor_not_counted=ops[5] / per,
            integrator=[k for k, v in tasks.INTEGRATORS.items() if v == task.integrator][0],
            substeps=task.n_substeps, rhs_evals_per_env_step=task.n_substeps * rhs_per_sub + 1,
            flops_per_rhs_eval_incl_overheads=flops / (task.n_substeps * rhs_per_sub + 1),
            sample="%d envs x %d control steps, %d resets, actions U[%g,%g]" % (n, steps, resets, lo, hi))
        print(env_id, "%.0f flops/env-step (%.0f per RHS evaluation)" % (flops, out[env_id]["flops_per_rhs_eval_incl_overheads"]))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "flop_count.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
