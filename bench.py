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
host cores -- the benchmark build of the oracle (-O3 -march=native, FMA
contraction, OpenMP over envs, compiled on the host that runs it), the SAME
batch as the CUDA arm (4096 envs), INNER control steps per timed "step".

`e2e`: every control step reads its actions from and writes its observation / reward / done / terms rows to
page-locked HOST memory (the step kernel does both in place, over PCIe).  Two public calls are timed and the
better one is `e2e.value` (`e2e.api` names it): the native env-group loop `bio_groups_run` (the batch as 4 groups
on disjoint SMs, a group relaunched when its rows have landed, so one group's PCIe tail hides behind the others'
substeps) and the synchronous whole-batch call `bio_step_host` (`e2e.synchronous`); `e2e.pipelined` is the same
group pipeline driven by a Python send / recv loop.

Before anything is timed the CUDA arm rolls the batch 150 control steps
(untimed, independent of --warmup), so that the timed window holds the
auto-resets of a steady-state rollout (`rollout.episodes` counts the episodes
that ended inside the timed + end-to-end loops).  With --gpus N >= 2 the line
also carries `config5`: BASELINE.json's 1 048 576-env sweep (Palsy3D and
LockedKnee3D at 1 048 576 / N envs per GPU).
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


PREROLL = 150          # untimed control steps before warm-up (steady-state mix of episode ages)
REF_INNER = 4          # control steps of the whole batch per timed "step" of the reference arm


def _fast_oracle():
    """The benchmark build of the oracle for this host; returns the oracle module bound to it."""
    from oracle import oracle as orc
    try:
        os.environ["BIO_ORACLE_LIB"] = orc.build_fast()
        orc._lib = None
        kind = "-O3 -march=native -ffp-contract=fast, OpenMP"
    except Exception as e:                 # no compiler on this host: the parity build that travelled with the repo
        os.environ.pop("BIO_ORACLE_LIB", None)
        orc._lib = None
        orc.build()
        kind = "parity build -O2 (fast build failed: %s)" % type(e).__name__
    return orc, kind


def cpu_arm(env_id, n_envs, steps, warmup, budget_s, threads, integrator=None, inner=1):
    """Oracle (CPU port) throughput on `threads` host threads; integrator=None: the stated fixed-step
    scheme of the CUDA path, 'adaptive_rkm': error-controlled Runge-Kutta-Merson at accuracy 1e-3, the
    stand-in for the reference's default opensim.Manager integrator (restated algorithm, not OpenSim).
    One "step" = `inner` control steps of all n_envs envs."""
    from bioimitation_gym_b200 import registry
    orc, build = _fast_oracle()
    spec, cm, ref, task = registry.build_env_tables(env_id, dict(integrator=integrator) if integrator else {})
    rt = orc.RefTables(ref["q"], ref["u"], ref["body_pos"], ref["com_pos"])
    env = orc.OracleVecEnv(cm.tables, task, rt, n_envs, seed=0, threads=threads)
    env.reset()
    rng = np.random.default_rng(0)
    lo, hi = (-1.0, 1.0) if spec.torque else (0.0, 1.0)
    acts = [rng.uniform(lo, hi, (n_envs, cm.tables.n_act)) for _ in range(4)]
    for k in range(warmup * inner):
        env.step(acts[k % 4])
    t0 = time.perf_counter()
    done_steps = 0
    for k in range(steps):
        for j in range(inner):
            env.step(acts[(k * inner + j) % 4])
        done_steps += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(value=n_envs * inner * done_steps / dt, unit="env-steps/s", cores=threads, kind="port",
                sample="%d envs x %d control steps of %s on %d threads in %.1f s (oracle/bio_oracle.c, fp64, %s, %s)"
                       % (n_envs, done_steps * inner, env_id, threads, dt, build,
                          "adaptive Runge-Kutta-Merson, accuracy 1e-3" if integrator else "same fixed-step scheme")), \
        dt / max(done_steps * inner, 1)


def bind_to_gpu_numa_node(local_rank, world=1):
    """CPU affinity of this rank = the cores of its GPU's NUMA node (sysfs), before any pinned allocation, so
    that the page-locked buffers the step kernel writes over PCIe are node-local.  Where the host hides the
    topology (a VM with one virtual node: numa_node = -1 for every device) the affinity is left alone unless
    BIO_BENCH_BIND_SLICES=1 gives each rank its own contiguous slice of the allowed cores (measured: no gain).
    BIO_BENCH_NO_BIND=1 leaves the affinity alone.  Returns what was done."""
    info = {"numa_node": None, "cpus": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None,
            "bound": False}
    if os.environ.get("BIO_BENCH_NO_BIND") == "1":
        return info
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = int(vis.split(",")[local_rank]) if vis else local_rank
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        base = "/sys/bus/pci/devices/" + bus
        info["numa_node"] = int(open(base + "/numa_node").read())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        want = cpus & allowed
        if want and want != allowed:
            os.sched_setaffinity(0, want)
            info["bound"] = True
        info["cpus"] = len(os.sched_getaffinity(0))
    except Exception as e:                  # no NVML / sysfs: leave the affinity alone
        info["error"] = type(e).__name__
    # measured on an 8 x B200 box (32 vCPUs, one virtual node; gpurun_out/e2e_probe_8_*.json): slices 152.4 M end-to-end
    # env-steps/s, no binding 155.9 M -- so the slices are opt-in
    if not info["bound"] and world > 1 and os.environ.get("BIO_BENCH_BIND_SLICES") == "1" and hasattr(os, "sched_getaffinity"):
        allowed = sorted(os.sched_getaffinity(0))
        per = len(allowed) // world
        if per >= 2:
            os.sched_setaffinity(0, set(allowed[local_rank * per:(local_rank + 1) * per]))
            info["bound"] = True
            info["slice"] = "%d cores from the %d-th of %d slices" % (per, local_rank, world)
            info["cpus"] = per
    return info


def timed_loop(env, pool, steps, flush, barrier):
    """`steps` control steps, each bracketed by CUDA events on the launching stream; returns total ms."""
    import torch
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    barrier()
    for k in range(steps):
        if flush is not None:
            flush.fill_(k & 0xFF)          # evict L2 between timed iterations (untimed)
        starts[k].record()
        env.step(pool[k % len(pool)])
        stops[k].record()
    barrier()
    return float(sum(s.elapsed_time(e) for s, e in zip(starts, stops)))


def e2e_loop(env, pool, steps, barrier):
    """Host-buffer entry point (pinned host memory in, pinned host memory out, synchronous): wall seconds."""
    import torch
    N, na, D = env.num_envs, env.n_act, env.obs_dim
    pin = lambda *s, dt=None: torch.empty(s, dtype=dt or env.dtype).pin_memory()
    a_pin = [pin(N, na) for _ in range(4)]
    for i, ap_ in enumerate(a_pin):
        ap_.copy_(pool[i].cpu())
    o_pin, r_pin, d_pin, t_pin = pin(N, D), pin(N), pin(N, dt=torch.uint8), pin(N, env.n_terms)
    a_np = [t_.numpy() for t_ in a_pin]
    o_np, r_np, d_np, t_np = o_pin.numpy(), r_pin.numpy(), d_pin.numpy(), t_pin.numpy()
    for k in range(3):
        env.step_host(a_np[k % 4], o_np, r_np, d_np, t_np)
    barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        env.step_host(a_np[k % 4], o_np, r_np, d_np, t_np)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


def e2e_pipelined_loop(env_id, cfg, pool, steps, preroll, barrier, n_groups=4):
    """The same end-to-end step through the send / recv entry points (backend.EnvGroups): the batch as
    `n_groups` groups on disjoint SMs, each group's next step sent as soon as its results are in host memory, so
    that one group's PCIe tail overlaps the others' substeps.  Every control step of every group reads its
    actions from and writes its observation rows to page-locked host memory, as in e2e_loop.  Wall seconds for
    `steps` control steps of the whole batch."""
    import torch
    from bioimitation_gym_b200 import backend
    groups = backend.EnvGroups(env_id, cfg, groups=n_groups)
    ng = groups.n_group
    acts = [[pool[i][g * ng:(g + 1) * ng].cpu().numpy().copy() for g in range(n_groups)] for i in range(4)]
    groups.reset()
    for g in range(n_groups):
        groups.send(g, acts[0][g])
    for k in range(1, preroll + 3):           # the same steady-state episode mix as the synchronous loops
        for g in range(n_groups):
            groups.recv(g)
            groups.send(g, acts[k % 4][g])
    for g in range(n_groups):
        groups.recv(g)
    barrier()
    t0 = time.perf_counter()
    for g in range(n_groups):
        groups.send(g, acts[0][g])
    for k in range(1, steps):
        for g in range(n_groups):
            groups.recv(g)
            groups.send(g, acts[k % 4][g])
    for g in range(n_groups):
        groups.recv(g)
    dt = time.perf_counter() - t0
    # the same loop inside the library (bio_groups_run): no host-language code between a group's steps, the four
    # action buffers of every group replayed in place from page-locked memory
    ring = [[torch.as_tensor(acts[i][g]).pin_memory() for i in range(4)] for g in range(n_groups)]
    groups.run(3, action_ring=ring)
    for e in groups.envs:
        e.stats(reset=True)
    barrier()
    t0 = time.perf_counter()
    groups.run(steps, action_ring=ring)
    dt_native = time.perf_counter() - t0
    episodes = float(sum(float(e.stats()[1].item()) for e in groups.envs))
    groups.close()
    return dt, dt_native, episodes


def e2e_steps_for(n_per_gpu, steps):
    """Control steps of the end-to-end loops: at least ~40 ms of wall time (a 4 ms window of 20 small steps lets
    one scheduling hiccup on one of eight ranks move the max-over-ranks figure by 10 %), at most 1000."""
    per_step_ms = max(0.2, n_per_gpu / 4096.0 * 0.2)
    return int(max(min(steps, 1000), min(1000, round(40.0 / per_step_ms)), 10))


def run_config(env_id, n_per_gpu, steps, warmup, rank, world, local_rank, args, preroll, e2e_steps, barrier, flush):
    """One workload: create, pre-roll, warm up, timed loop, end-to-end loop, statistics.  Returns a dict
    (identical on every rank for the all-reduced fields)."""
    import torch
    import torch.distributed as dist
    from bioimitation_gym_b200 import backend
    dev = torch.device("cuda", local_rank)
    cfg = dict(num_envs=n_per_gpu, device=local_rank, dtype=args.dtype, seed=1234, env_offset=rank * n_per_gpu)
    if args.substeps:
        cfg["substeps"] = args.substeps
    if args.integrator:
        cfg["integrator"] = args.integrator
    env = backend.VecEnv(env_id, cfg)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    lo, hi = (-1.0, 1.0) if env.spec.torque else (0.0, 1.0)
    pool = [torch.rand((n_per_gpu, env.n_act), generator=g, device=dev, dtype=env.dtype) * (hi - lo) + lo
            for _ in range(16 if n_per_gpu <= 65536 else 4)]
    env.reset()
    for k in range(preroll + warmup):
        env.step(pool[k % len(pool)])
    barrier()
    env.stats(reset=True)                      # count the episodes of the timed + end-to-end loops only
    launches0 = env.launch_count
    total_ms = timed_loop(env, pool, steps, flush, barrier)
    launches = env.launch_count - launches0
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    e2e_s = e2e_loop(env, pool, e2e_steps, barrier)
    t = torch.tensor([e2e_s, -e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_max, e2e_min = float(t[0].item()), -float(t[1].item())
    pipe = native = None
    if args.pipelined_groups > 1 and n_per_gpu % args.pipelined_groups == 0 and n_per_gpu <= 65536:
        pipe_s, native_s, native_eps = e2e_pipelined_loop(env_id, cfg, pool, e2e_steps, preroll, barrier,
                                                          args.pipelined_groups)
        t = torch.tensor([pipe_s, native_s, -native_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ng = n_per_gpu // args.pipelined_groups
        pipe = {"value": n_per_gpu * world * e2e_steps / float(t[0].item()), "unit": "env-steps/s",
                "groups": args.pipelined_groups, "steps": e2e_steps,
                "api": "backend.EnvGroups send / recv from a Python loop (bio_step_host_begin / _end): %d groups of %d "
                       "envs per GPU on disjoint SMs, same host buffers and bytes per step" % (args.pipelined_groups, ng)}
        native = {"value": n_per_gpu * world * e2e_steps / float(t[1].item()),
                  "per_rank_ms_per_step": {"min": -float(t[2].item()) / e2e_steps * 1e3,
                                           "max": float(t[1].item()) / e2e_steps * 1e3},
                  "episodes_rank0": native_eps,
                  "api": "bio_groups_run (backend.EnvGroups.run): native send / recv loop over %d groups of %d envs per "
                         "GPU on disjoint SMs; every control step of every group reads its actions from and writes "
                         "its observation / reward / done / terms rows to page-locked host memory (the kernel, in "
                         "place), next step launched when the rows have landed" % (args.pipelined_groups, ng)}
    stats = env.stats().clone()
    if world > 1:                              # the only collective: rollout statistics, outside the step path
        gathered = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(gathered, stats)
        stats = torch.stack(gathered).sum(0)
    stats = stats.cpu().numpy()
    esz = 4 if env.dtype == torch.float32 else 8
    # issue-bound kernels (3D) gain nothing from groups: e2e.value is the better of the two public calls, the other
    # one is kept beside it
    sync_better = bool(native) and native["value"] < n_per_gpu * world * e2e_steps / e2e_max
    sync_e2e = {"value": n_per_gpu * world * e2e_steps / e2e_max, "unit": "env-steps/s",
                "per_rank_ms_per_step": {"min": e2e_min / e2e_steps * 1e3, "max": e2e_max / e2e_steps * 1e3},
                "api": "bio_step_host (VecEnv.step_host): the whole batch in one synchronous call, page-locked host "
                       "buffers read and written by the kernel in place"}
    best = sync_e2e if (native is None or sync_better) else native
    out = dict(env=env, value=n_per_gpu * world * steps / (total_ms * 1e-3), ms_per_step=total_ms / steps,
               launches=int(launches), lo=lo, hi=hi,
               e2e={"value": best["value"], "unit": "env-steps/s",
                    "h2d_bytes_per_step": n_per_gpu * env.n_act * esz,
                    "d2h_bytes_per_step": n_per_gpu * (env.obs_dim + 1 + env.n_terms) * esz + n_per_gpu,
                    "steps": e2e_steps,
                    "api": best["api"],
                    "per_rank_ms_per_step": best["per_rank_ms_per_step"],
                    "synchronous": sync_e2e,
                    **({"native_groups": native} if native else {}),
                    **({"pipelined": pipe} if pipe else {})},
               rollout={"env_steps": float(stats[0]), "episodes": float(stats[1]),
                        "mean_return": float(stats[2] / max(stats[1], 1)),
                        "mean_length": float(stats[3] / max(stats[1], 1)),
                        "done_height": float(stats[4]), "done_limit": float(stats[5]),
                        "done_accel": float(stats[6]), "done_horizon": float(stats[7]),
                        "done_feet": float(stats[8]), "done_nonfinite": float(stats[9]),
                        "preroll_steps": preroll})
    return out


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
    ap.add_argument("--no-config5", action="store_true", help="skip the 1M-env sweep lines at --gpus >= 2")
    ap.add_argument("--preroll", type=int, default=PREROLL)
    ap.add_argument("--pipelined-groups", type=int, default=4,
                    help="env groups of the send / recv end-to-end measurement (e2e.pipelined); <= 1: skip")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        warm = max(3, args.warmup)
        cb, spstep = cpu_arm(args.env_id, args.envs_per_gpu, args.steps, min(warm, 3), 240.0, cores, inner=REF_INNER)
        from bioimitation_gym_b200 import registry, tasks
        spec, cm, ref, task = registry.build_env_tables(args.env_id, {})
        line = {"impl": "reference", "metric": "env_steps_per_sec", "value": cb["value"], "unit": "env-steps/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": spstep * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "%s, %d envs/GPU (reference arm: the same batch of %d envs on the host CPU, "
                                       "one timed step = %d control steps of the batch)"
                                       % (args.env_id, args.envs_per_gpu, args.envs_per_gpu, REF_INNER),
                           "integrator": tasks.DEFAULT_INTEGRATOR, "substeps": task.n_substeps,
                           "note": "OpenSim 4.1 is not installable offline; this is the CPU restatement "
                                   "(oracle port) of the same algorithm, not OpenSim"},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    affinity = bind_to_gpu_numa_node(local_rank, world)   # before CUDA / pinned allocations
    import torch
    import torch.distributed as dist
    from bioimitation_gym_b200 import tasks

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
    flush = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks / throttle reasons are sampled from the warm-up to the end of the end-to-end loop (the timed
    # region alone lasts tens of milliseconds, less than one nvidia-smi poll)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    res = run_config(args.env_id, N, args.steps, args.warmup, rank, world, local_rank, args, args.preroll,
                     e2e_steps_for(N, args.steps), barrier, flush)
    env = res["env"]
    shape = dict(zip(("size_class", "threads", "ctas_per_sm"), env.coop_shape()))
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    if res["rollout"]["episodes"] <= 0 and args.preroll >= PREROLL and env.task.auto_reset:
        raise SystemExit("no episode ended inside the timed region: the window does not hold auto-resets")

    # ---- BASELINE.json config 5 at N >= 2: 1 048 576 envs over the N GPUs (kept short: 3 + 10 steps) ----
    config5 = None
    if world > 1 and not args.no_config5 and args.env_id == ENV_ID:
        config5 = {}
        env.close()
        for env5 in ("MusclePalsyImitation3D-v0", "MuscleLockedKneeImitation3D-v0"):
            n5 = 1048576 // world
            r5 = run_config(env5, n5, 10, 3, rank, world, local_rank, args, 0, 5, barrier, None)
            f5, _ = _flops_per_env_step(r5["env"].task, env5)
            r5["env"].close()
            config5[env5] = {"envs_total": n5 * world, "envs_per_gpu": n5, "value": r5["value"],
                             "ms_per_step": r5["ms_per_step"], "e2e": r5["e2e"], "steps": 10, "warmup": 3,
                             "l2": "inputs larger than L2 (no flush)", "episodes": r5["rollout"]["episodes"],
                             "flops_per_env_step": f5}

    if rank == 0:
        flops, fdetail = _flops_per_env_step(env.task, args.env_id)
        ms_launch = res["ms_per_step"]
        roofline = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        is64 = env.dtype == torch.float64
        fp_peak = None
        try:                                       # FMA chain in the handle's scalar type, measured now, untimed
            v = float((env.lib.bio_measure_fp64_peak if is64 else env.lib.bio_measure_fp32_peak)(local_rank))
            fp_peak = v if v > 0 else None
        except Exception:
            pass
        peak = fp_peak or (148 * 128 * 2 * 1.965e9 / 1e12 / (64 if is64 else 1))
        if flops is not None:
            achieved = flops * N / (ms_launch * 1e-3) / 1e12
            roofline = {"bound": "fp64" if is64 else "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                        "frac": achieved / peak,
                        "peak_source": ("measured on this GPU by bio_measure_%s_peak (%s chain, FMA = 2 flops)"
                                        % (("fp64", "DFMA") if is64 else ("fp32", "FFMA"))) if fp_peak
                        else "nominal (no measured peak)",
                        "flops_per_env_step": flops, "traffic": None}
            try:
                roofline["traffic"] = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))[
                    "bytes_per_launch"]
            except Exception:
                pass
        if config5:
            for v5 in config5.values():
                if v5["flops_per_env_step"]:
                    v5["frac_of_fp32_peak_per_gpu"] = v5["flops_per_env_step"] * v5["value"] / world / 1e12 / peak
        esz = 4 if env.dtype == torch.float32 else 8
        na, D = env.n_act, env.obs_dim
        state_bytes = (2 * env.n_dof + 2 * env.n_muscles) * esz
        alg_bytes = (na * esz + 2 * state_bytes + 2 * env.task.horizon * na * esz + D * esz + (2 + env.n_terms) * esz)
        hbm = {"bound": "hbm", "achieved": alg_bytes * N / (ms_launch * 1e-3) / 1e9,
               "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s", "bytes_per_env_step": alg_bytes,
               "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s"}
        hbm["frac"] = hbm["achieved"] / hbm["peak"]
        line = {"metric": "env_steps_per_sec", "value": res["value"], "unit": "env-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_launch, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if env.dtype == torch.float32 else "f64", "data": "synthetic",
                "config": {"workload": "%s, %d envs/GPU, actions U[%g,%g] resident in HBM"
                                       % (args.env_id, N, res["lo"], res["hi"]),
                           "integrator": [k for k, v in tasks.INTEGRATORS.items() if v == env.task.integrator][0],
                           "substeps": env.task.n_substeps, "h_seconds": env.task.dt / env.task.n_substeps,
                           "preroll": "%d untimed control steps before warm-up (steady-state episode mix)" % args.preroll,
                           "l2": "flushed between timed iterations (256 MiB write, untimed)" if flush is not None
                           else "not flushed", "sharding": "env index, no collective in the step path",
                           "kernel_shape": shape},
                "roofline": roofline, "roofline_hbm": hbm, "e2e": res["e2e"],
                "gpu_launches": res["launches"], "clocks": sampler.summary(), "rollout": res["rollout"],
                "affinity": affinity}
        if config5:
            line["config5"] = config5
        if world == 1 and not args.no_cpu_baseline:
            cb, _ = cpu_arm(args.env_id, N, 10 ** 9, 1, 15.0, cores)
            # the reference's own integrator setting (adaptive, accuracy 1e-3), restated: context for the ratio
            ad, _ = cpu_arm(args.env_id, min(N, cores * 32), 10 ** 9, 1, 8.0, cores, integrator="adaptive_rkm")
            cb["adaptive_scheme"] = {"value": ad["value"], "unit": ad["unit"], "sample": ad["sample"]}
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if not config5:
        env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
