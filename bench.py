#!/usr/bin/env python
"""Benchmark of the hot path: batched env steps per second.

  python bench.py --gpus N --steps K --warmup W           # this repo (CUDA)
  python bench.py --impl reference --steps K --warmup W   # CPU arm (see below)

One "step" = one 0.01 s control step of every env of the batch (action
pre-processing, S fixed substeps, observation, reward, done, auto-reset).
Workload at N=1 = BASELINE.json configs[1]: MuscleWalkingImitation2D-v0,
4096 envs per GPU; envs are sharded by env index over ranks (weak scaling, no
collective in the step path; one NCCL all-gather of the rollout statistics
after the timed region).

The reference arm cannot run OpenSim (not installable offline): it
times the CPU restatement of the same algorithm (oracle/, kind "port") on all
host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENV_ID = "MuscleWalkingImitation2D-v0"
ENVS_PER_GPU = 4096


def _flops_per_env_step(task, env_id):
    """Algorithmic FLOPs of one env step from the instrumented-oracle count
    (profiles/flop_count.json, produced by oracle/count_flops.py)."""
    path = os.path.join(ROOT, "profiles", "flop_count.json")
    if not os.path.exists(path):
        return None, None
    d = json.load(open(path))
    e = d.get(env_id)
    if not e:
        return None, None
    from bioimitation_gym_b200 import tasks
    integ = [k for k, v in tasks.INTEGRATORS.items() if v == task.integrator][0]
    if e["integrator"] == integ and e["substeps"] == task.n_substeps:
        return e["flops_per_env_step"], e
    # other scheme / substep count: scale by the number of dynamics evaluations
    rhs_per_sub = {0: 1, 1: 2, 2: 4, 3: 1}[task.integrator]
    evals = task.n_substeps * rhs_per_sub + 1
    return e["flops_per_rhs_eval_incl_overheads"] * evals, e


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the run: NVML polled every 2 ms (the timed region lasts tens
    of milliseconds), nvidia-smi every 100 ms when NVML cannot be loaded.  With CUDA_VISIBLE_DEVICES set
    the NVML index is the visible device's entry of that list."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        super().__init__(daemon=True)
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        try:
            index = int(vis.split(",")[index]) if vis else index
        except (ValueError, IndexError):
            pass
        self.index = index
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag = False
        self.source = "nvml"

    def _nvml(self):
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        bits = [(pynvml.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                (pynvml.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (pynvml.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                (pynvml.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]
        mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        while not self.stop_flag:
            self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            self.mx.append(mx)
            r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            for bit, name in bits:
                if r & bit:
                    self.reasons.add(name)
            time.sleep(0.002)

    def _smi(self):
        self.source = "nvidia-smi"
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                r = [x.strip() for x in out.strip().split(",")]
                self.sm.append(float(r[0]))
                self.mx.append(float(r[1]))
                for nm, v in zip(self.NAMES, r[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.1)

    def run(self):
        try:
            self._nvml()
        except Exception:
            self._smi()

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(max(self.mx)), "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": self.source}


def cpu_arm(env_id, n_envs, steps, warmup, budget_s, threads, integrator=None):
    """Oracle (CPU port) throughput on `threads` host threads; integrator=None: the stated fixed-step
    scheme of the CUDA path, 'adaptive_rkm': error-controlled Runge-Kutta-Merson at accuracy 1e-3, the
    stand-in for the reference's default opensim.Manager integrator (restated algorithm, not OpenSim)."""
    from bioimitation_gym_b200 import registry
    from oracle import oracle as orc
    orc.build()
    spec, cm, ref, task = registry.build_env_tables(env_id, dict(integrator=integrator) if integrator else {})
    rt = orc.RefTables(ref["q"], ref["u"], ref["body_pos"], ref["com_pos"])
    env = orc.OracleVecEnv(cm.tables, task, rt, n_envs, seed=0, threads=threads)
    env.reset()
    rng = np.random.default_rng(0)
    lo, hi = (-1.0, 1.0) if spec.torque else (0.0, 1.0)
    for _ in range(warmup):
        env.step(rng.uniform(lo, hi, (n_envs, cm.tables.n_act)))
    t0 = time.perf_counter()
    done_steps = 0
    for _ in range(steps):
        env.step(rng.uniform(lo, hi, (n_envs, cm.tables.n_act)))
        done_steps += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(value=n_envs * done_steps / dt, unit="env-steps/s", cores=threads, kind="port",
                sample="%d envs x %d control steps of %s on %d threads (oracle/bio_oracle.c, fp64, %s)"
                       % (n_envs, done_steps, env_id, threads,
                          "adaptive Runge-Kutta-Merson, accuracy 1e-3" if integrator else "same fixed-step scheme")), \
        dt / max(done_steps, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--env-id", default=ENV_ID)
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--substeps", type=int, default=None)
    ap.add_argument("--integrator", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        n = cores * 16
        cb, spstep = cpu_arm(args.env_id, n, args.steps, max(1, min(args.warmup, 2)), 120.0, cores)
        from bioimitation_gym_b200 import registry, tasks
        spec, cm, ref, task = registry.build_env_tables(args.env_id, {})
        line = {"impl": "reference", "metric": "env_steps_per_sec", "value": cb["value"], "unit": "env-steps/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": spstep * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "%s, %d envs/GPU (reference arm: bounded sample of %d envs on host CPU)"
                                       % (args.env_id, args.envs_per_gpu, n),
                           "integrator": tasks.DEFAULT_INTEGRATOR, "substeps": task.n_substeps,
                           "note": "OpenSim 4.1 is not installable offline; this is the CPU restatement "
                                   "(oracle port) of the same algorithm, not OpenSim"},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from bioimitation_gym_b200 import backend, tasks

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version banner to stdout when the communicator is created: keep stdout to the one
        # JSON line by pointing fd 1 at stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    dev = torch.device("cuda", local_rank)
    N = args.envs_per_gpu
    cfg = dict(num_envs=N, device=local_rank, dtype=args.dtype, seed=1234, env_offset=rank * N)
    if args.substeps:
        cfg["substeps"] = args.substeps
    if args.integrator:
        cfg["integrator"] = args.integrator
    env = backend.VecEnv(args.env_id, cfg)
    na, D = env.n_act, env.obs_dim
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    lo, hi = (-1.0, 1.0) if env.spec.torque else (0.0, 1.0)
    pool = [torch.rand((N, na), generator=g, device=dev, dtype=env.dtype) * (hi - lo) + lo for _ in range(16)]
    flush = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    env.reset()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks / throttle reasons are sampled from the warm-up to the end of the end-to-end loop (the timed
    # region alone lasts tens of milliseconds, less than one nvidia-smi poll)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for k in range(args.warmup):
        env.step(pool[k % len(pool)])
    barrier()
    launches0 = env.launch_count
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        if flush is not None:
            flush.fill_(k & 0xFF)          # evict L2 between timed iterations (untimed)
        starts[k].record()
        env.step(pool[k % len(pool)])
        stops[k].record()
    barrier()
    launches = env.launch_count - launches0
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    total_ms = float(sum(step_ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = N * world * args.steps / (total_ms * 1e-3)

    # ---- end-to-end through the host-buffer entry point (pinned host memory) ----
    np_dt = np.float32 if env.dtype == torch.float32 else np.float64
    pin = lambda *s, dt=None: torch.empty(s, dtype=dt or env.dtype).pin_memory()
    a_pin = [pin(N, na) for _ in range(4)]
    for i, ap_ in enumerate(a_pin):
        ap_.copy_(pool[i].cpu())
    o_pin, r_pin, d_pin, t_pin = pin(N, D), pin(N), pin(N, dt=torch.uint8), pin(N, env.n_terms)
    e2e_steps = max(10, min(args.steps, 100))
    # page-locked buffers: bio_step_host lets the step kernel read the actions from and write the results into
    # them in place (PCIe traffic inside the timed region: h2d / d2h bytes below), and returns after a stream
    # synchronisation, i.e. when the host can read the results
    a_np = [t_.numpy() for t_ in a_pin]
    o_np, r_np, d_np, t_np = o_pin.numpy(), r_pin.numpy(), d_pin.numpy(), t_pin.numpy()
    for k in range(3):
        env.step_host(a_np[k % 4], o_np, r_np, d_np, t_np)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        env.step_host(a_np[k % 4], o_np, r_np, d_np, t_np)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = N * world * e2e_steps / float(t.item())
    esz = 4 if env.dtype == torch.float32 else 8
    h2d = N * na * esz
    d2h = N * (D + 1 + env.n_terms) * esz + N

    # ---- rollout statistics: the only collective (outside the step path) ----
    stats = env.stats().clone()
    if world > 1:
        gathered = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(gathered, stats)
        stats = torch.stack(gathered).sum(0)
    stats = stats.cpu().numpy()

    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
        flops, fdetail = _flops_per_env_step(env.task, args.env_id)
        ms_launch = total_ms / args.steps
        roofline = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        fp32_peak = None
        try:
            v = float(env.lib.bio_measure_fp32_peak(local_rank))   # FFMA chain, measured now, untimed
            fp32_peak = v if v > 0 else None
        except Exception:
            pass
        if flops is not None:
            achieved = flops * N / (ms_launch * 1e-3) / 1e12
            peak = fp32_peak or 148 * 128 * 2 * 1.965e9 / 1e12
            roofline = {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                        "frac": achieved / peak,
                        "peak_source": "measured on this GPU by bio_measure_fp32_peak (FFMA chain, FMA = 2 flops)" if fp32_peak
                        else "nominal 148 SM x 128 lanes x 2 x 1.965 GHz (no measured FP32 peak yet)",
                        "flops_per_env_step": flops, "traffic": None}
            try:
                roofline["traffic"] = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))[
                    "bytes_per_launch"]
            except Exception:
                pass
        state_bytes = (2 * env.n_dof + 2 * env.n_muscles) * esz
        alg_bytes = (na * esz + 2 * state_bytes + 2 * env.task.horizon * na * esz + D * esz + (2 + env.n_terms) * esz)
        hbm = {"bound": "hbm", "achieved": alg_bytes * N / (ms_launch * 1e-3) / 1e9,
               "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s", "bytes_per_env_step": alg_bytes,
               "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s"}
        hbm["frac"] = hbm["achieved"] / hbm["peak"]
        line = {"metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_launch, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if env.dtype == torch.float32 else "f64", "data": "synthetic",
                "config": {"workload": "%s, %d envs/GPU, actions U[%g,%g] resident in HBM" % (args.env_id, N, lo, hi),
                           "integrator": [k for k, v in tasks.INTEGRATORS.items() if v == env.task.integrator][0],
                           "substeps": env.task.n_substeps, "h_seconds": env.task.dt / env.task.n_substeps,
                           "l2": "flushed between timed iterations (256 MiB write, untimed)" if flush is not None
                           else "not flushed", "sharding": "env index, no collective in the step path"},
                "roofline": roofline, "roofline_hbm": hbm,
                "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "steps": e2e_steps},
                "gpu_launches": int(launches), "clocks": sampler.summary(),
                "rollout": {"env_steps": float(stats[0]), "episodes": float(stats[1]),
                            "mean_return": float(stats[2] / max(stats[1], 1)),
                            "mean_length": float(stats[3] / max(stats[1], 1)),
                            "done_height": float(stats[4]), "done_limit": float(stats[5]),
                            "done_accel": float(stats[6]), "done_horizon": float(stats[7]),
                            "done_feet": float(stats[8]), "done_nonfinite": float(stats[9])}}
        if world == 1 and not args.no_cpu_baseline:
            cb, _ = cpu_arm(args.env_id, cores * 16, 10 ** 9, 1, 15.0, cores)
            # the reference's own integrator setting (adaptive, accuracy 1e-3), restated: context for the ratio
            ad, _ = cpu_arm(args.env_id, cores * 8, 10 ** 9, 1, 8.0, cores, integrator="adaptive_rkm")
            cb["adaptive_scheme"] = {"value": ad["value"], "unit": ad["unit"], "sample": ad["sample"]}
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
