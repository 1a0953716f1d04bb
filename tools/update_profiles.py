#!/usr/bin/env python
"""Refresh profiles/ from one `ncu --set full` report and one launch list (both brought back in
gpurun_out/): key metrics (json), DRAM traffic per launch, per-phase instruction / stall-sample
report, launch summary.

  python tools/update_profiles.py gpurun_out/prof.ncu-rep gpurun_out/launches.csv r01
  python tools/update_profiles.py gpurun_out/x.ncu-rep gpurun_out/x_launches.csv r02 --name coop3d --warps 8192
(--name: file stem instead of "coop_kernel" -- only the default stem refreshes dram_traffic.json, the bench's
`roofline.traffic`; --warps / --evals: warps per launch and evaluations per step for the per-warp figures)
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def main():
    rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    opt = sys.argv[4:]
    name = opt[opt.index("--name") + 1] if "--name" in opt else "coop_kernel"
    warps = opt[opt.index("--warps") + 1] if "--warps" in opt else "2048"
    evals = opt[opt.index("--evals") + 1] if "--evals" in opt else "21"
    prof = os.path.join(ROOT, "profiles")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {k: {"value": vals[hdr.index(k)], "unit": units[hdr.index(k)]} for k in WANT if k in hdr}
    for i, k in enumerate(hdr):          # every stall reason, per issued instruction
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
            out[k] = {"value": vals[i], "unit": units[i]}
    out["kernel"] = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
    out["source"] = "ncu --set full --clock-control none, one launch of the bench's timed loop (%s)" % os.path.basename(rep)
    json.dump(out, open(os.path.join(prof, "%s_%s_ncu_metrics.json" % (tag, name)), "w"), indent=1)

    def mb(k):
        v, u = float(out[k]["value"]), out[k]["unit"].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    traffic = int(mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum"))
    out["dram_bytes_per_launch"] = traffic
    json.dump(out, open(os.path.join(prof, "%s_%s_ncu_metrics.json" % (tag, name)), "w"), indent=1)
    if name == "coop_kernel":
        json.dump({"bytes_per_launch": traffic,
                   "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, %s, 4096 envs "
                             "(profiles/%s_coop_kernel_ncu_metrics.json); outputs stay in the 126 MB L2, so writes "
                             "barely reach DRAM" % (out["kernel"][:40], tag)},
                  open(os.path.join(prof, "dram_traffic.json"), "w"))
    src = rep.replace(".ncu-rep", "_src.csv")
    with open(src, "w") as fh:
        subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=fh,
                       stderr=subprocess.DEVNULL)
    rpt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_phase_report.py"), src, warps, evals, "--lines", "40"],
                         capture_output=True, text=True).stdout
    open(os.path.join(prof, "%s_%s_phase_report.txt" % (tag, name)), "w").write(rpt)
    if os.path.exists(launches):
        summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), launches],
                              capture_output=True, text=True).stdout
        open(os.path.join(prof, "%s_%s_launches_summary.txt" % (tag, name) if name != "coop_kernel" else "%s_launches_summary.txt" % tag), "w").write(summ)
        with open(os.path.join(prof, ("%s_%s_launches.csv" % (tag, name)) if name != "coop_kernel" else "%s_launches.csv" % tag), "w") as fh:
            fh.write(open(launches).read())
    print("traffic", traffic, "B/launch;", out.get("gpu__time_duration.sum"))


if __name__ == "__main__":
    main()
