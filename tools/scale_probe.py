#!/usr/bin/env python
"""Kernel time of one control step against the number of envs (warps per SM): separates the per-warp
dependency latency (one warp per SM) from issue contention.   python tools/scale_probe.py [env_id]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bioimitation_gym_b200 import backend

env_id = sys.argv[1] if len(sys.argv) > 1 else "MuscleWalkingImitation2D-v0"
for n in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "296,592,1184,2368,3552,4096,4736,8192,16384".split(","))]:
    env = backend.VecEnv(env_id, dict(num_envs=n, seed=1))
    env.reset()
    lo = -1.0 if env.spec.torque else 0.0
    a = torch.rand((n, env.n_act), device=env.device) * (1.0 - lo) + lo
    for _ in range(20):
        env.step(a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(100):
        env.step(a)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 10.0
    print("%6d envs  %8.1f us/step  %7.2f M env-steps/s" % (n, us, n / us))
    env.close()
