#!/usr/bin/env python
"""Kernel time of one control step against batch size and launch shape (BIO_COOP_THREADS = lo | hi):
calibrates COOP_SHAPE_COST in bio_coop.cuh.

  python tools/shape_probe.py MuscleWalkingImitation3D-v0 2368 4096 8192 16384 131072
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bioimitation_gym_b200 import backend  # noqa: E402


def main():
    env_id = sys.argv[1]
    sizes = [int(x) for x in sys.argv[2:]]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for n in sizes:
        row = []
        for shape in ("lo", "hi"):
            os.environ["BIO_COOP_THREADS"] = shape
            env = backend.VecEnv(env_id, dict(num_envs=n, seed=1))
            lo, hi = (-1.0, 1.0) if env.spec.torque else (0.0, 1.0)
            a = torch.rand((n, env.n_act), device=env.device) * (hi - lo) + lo
            env.reset()
            for _ in range(60):
                env.step(a)
            steps = 30 if n <= 20000 else 8
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            torch.cuda.synchronize()
            for s, e in ev:
                flush.fill_(1)
                s.record()
                env.step(a)
                e.record()
            torch.cuda.synchronize()
            ms = sum(s.elapsed_time(e) for s, e in ev) / steps
            row.append((env.coop_shape()[1], ms))
            env.close()
        del os.environ["BIO_COOP_THREADS"]
        env = backend.VecEnv(env_id, dict(num_envs=n, seed=1))
        picked = env.coop_shape()[1]
        env.close()
        print("%s n=%d: %s  -> M env-steps/s %s  picked %d" % (
            env_id, n, "  ".join("%d thr %.4f ms" % r for r in row),
            " / ".join("%.2f" % (n / r[1] / 1e3) for r in row), picked), flush=True)


if __name__ == "__main__":
    main()
