#!/bin/bash
# One measurement pass on the GPU box (run through gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/final_measure.sh'
# Writes gpurun_out/: bench lines (default, reference arm, every BASELINE config), the ncu launch list of the
# default bench command and one `ncu --set full` capture of the step kernel.  tools/update_profiles.py then
# copies the summaries into profiles/.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
: > gpurun_out/bench_configs.jsonl
run() { python bench.py --no-cpu-baseline --env-id "$1" --envs-per-gpu "$2" --steps "$3" --warmup 20 >> gpurun_out/bench_configs.jsonl 2>> gpurun_out/bench_configs.err; }
run TorqueWalkingImitation2D-v0 16384 200
run MuscleRunningImitation2D-v0 16384 200
run MuscleLockedKneeImitation2D-v0 16384 200
run MuscleJumpingImitation2D-v0 16384 200
run MuscleWalkingImitation2D-v0 16384 200
run MuscleWalkingImitation2D-v0 131072 50
run MuscleWalkingImitation3D-v0 8192 100
run MusclePalsyImitation3D-v0 131072 20
run MuscleLockedKneeImitation3D-v0 131072 20
run TorqueWalkingImitation3D-v0 16384 100
# fp64 build of the default workload (roofline against the DFMA chain measured in the same run)
python bench.py --no-cpu-baseline --dtype float64 --steps 50 --warmup 10 > gpurun_out/bench_fp64.json 2> gpurun_out/bench_fp64.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bio_coop_step_kernel -s 10 -c 1 -f -o gpurun_out/prof \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
