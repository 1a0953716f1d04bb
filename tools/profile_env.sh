#!/bin/bash
# ncu evidence for one env id (run through gpurun from the repo root):
#   gpurun --timeout 900 -- 'bash tools/profile_env.sh MuscleWalkingImitation3D-v0 8192 coop3d'
# Writes gpurun_out/<tag>_bench.json (no profiler), <tag>_launches.csv (ncu launch list) and <tag>.ncu-rep
# (one `ncu --set full` capture of the step kernel); tools/update_profiles.py copies the summaries to profiles/.
set -x
ENV_ID=$1; N=$2; TAG=$3
mkdir -p gpurun_out
python bench.py --no-cpu-baseline --env-id "$ENV_ID" --envs-per-gpu "$N" --steps 50 --warmup 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --no-cpu-baseline --env-id "$ENV_ID" --envs-per-gpu "$N" --steps 20 --warmup 3 > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bio_coop_step_kernel -s 10 -c 1 -f -o gpurun_out/${TAG} \
    python bench.py --no-cpu-baseline --env-id "$ENV_ID" --envs-per-gpu "$N" --steps 20 --warmup 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
cat gpurun_out/${TAG}_bench.json
