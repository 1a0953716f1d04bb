#!/bin/bash
# GPU test pass + short bench lines (run through gpurun from the repo root): writes gpurun_out/gputest.log,
# gpurun_out/bench_quick.jsonl
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/gputest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputest.log
tail -5 gpurun_out/gputest.log
: > gpurun_out/bench_quick.jsonl
for cfg in "MuscleWalkingImitation2D-v0 4096" "MuscleWalkingImitation2D-v0 16384" "MuscleWalkingImitation3D-v0 8192" "TorqueWalkingImitation3D-v0 16384"; do
  set -- $cfg
  python bench.py --no-cpu-baseline --env-id $1 --envs-per-gpu $2 --steps 100 --warmup 20 >> gpurun_out/bench_quick.jsonl 2>> gpurun_out/bench_quick.err
done
python - <<'PY'
import json
for ln in open('gpurun_out/bench_quick.jsonl'):
    d=json.loads(ln); print(d['config']['workload'][:48], '%.2fM' % (d['value']/1e6), 'ms %.4f' % d['ms_per_step'], 'e2e %.2fM (sync %.2fM)' % (d['e2e']['value']/1e6, d['e2e']['synchronous']['value']/1e6), 'frac %.3f' % d['roofline']['frac'], 'episodes', d['rollout']['episodes'])
PY
