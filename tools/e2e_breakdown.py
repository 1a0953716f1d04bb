#!/usr/bin/env python
"""Where the end-to-end step (host buffers) spends its time: the pieces of bio_step_host timed apart.

  python tools/e2e_breakdown.py [--envs 4096] [--iters 200]

Wall-clock per call with a device synchronize after every call (what a trainer sees)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, iters, sync):
    for _ in range(5):
        fn()
    sync()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    sync()
    return (time.perf_counter() - t0) / iters * 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--env-id", default="MuscleWalkingImitation2D-v0")
    args = ap.parse_args()
    import torch
    from bioimitation_gym_b200 import backend
    N = args.envs
    env = backend.VecEnv(args.env_id, dict(num_envs=N, seed=1))
    env.reset()
    dev = env.device
    sync = torch.cuda.synchronize
    a_dev = torch.rand((N, env.n_act), device=dev)
    pin = lambda *s, dt=torch.float32: torch.empty(s, dtype=dt).pin_memory()
    a_pin = pin(N, env.n_act)
    a_pin.copy_(a_dev.cpu())
    o_pin, r_pin, d_pin, t_pin = pin(N, env.obs_dim), pin(N), pin(N, dt=torch.uint8), pin(N, env.n_terms)
    an, on, rn, dn, tn = a_pin.numpy(), o_pin.numpy(), r_pin.numpy(), d_pin.numpy(), t_pin.numpy()
    out = {}
    out["step_host (H2D + kernel + 4 D2H + sync)"] = timed(lambda: env.step_host(an, on, rn, dn, tn), args.iters, sync)
    out["kernel only, sync each call"] = timed(lambda: (env.step(a_dev), sync()), args.iters, sync)
    out["kernel only, back to back"] = timed(lambda: env.step(a_dev), args.iters, sync)
    out["H2D actions, sync each"] = timed(lambda: (a_dev.copy_(a_pin, non_blocking=True), sync()), args.iters, sync)
    out["D2H obs, sync each"] = timed(lambda: (o_pin.copy_(env.obs, non_blocking=True), sync()), args.iters, sync)
    out["D2H obs+reward+done+terms, sync once"] = timed(
        lambda: (o_pin.copy_(env.obs, non_blocking=True), r_pin.copy_(env.reward, non_blocking=True),
                 d_pin.copy_(env.done, non_blocking=True), t_pin.copy_(env.terms, non_blocking=True), sync()),
        args.iters, sync)
    out["empty ctypes call"] = timed(lambda: env.lib.bio_launch_count(env.handle), args.iters * 10, sync)
    out["bare synchronize"] = timed(sync, args.iters * 10, sync)
    for k, v in out.items():
        print("%-45s %8.1f us" % (k, v))
    print("obs bytes %d -> D2H rate %.1f GB/s" % (o_pin.numel() * 4, o_pin.numel() * 4 / out["D2H obs, sync each"] / 1e3))
    env.close()


if __name__ == "__main__":
    main()
