#!/bin/bash
# End-to-end step rate of N ranks on one box under three host-side settings (run through gpurun --gpus N):
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/e2e_scaling_probe.sh 8'
# per-rank core slices, no binding, staged copies instead of zero-copy rows, default.
N=${1:-8}
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 50 --warmup 10 --no-cpu-baseline --no-config5 > gpurun_out/e2e_probe_${N}_${tag}.json 2> gpurun_out/e2e_probe_${N}_${tag}.err
  python - "$tag" gpurun_out/e2e_probe_${N}_${tag}.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.1fM  e2e %.1fM  per-rank ms %s  affinity %s" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["e2e"]["per_rank_ms_per_step"], d.get("affinity")))
PY
}
nproc
run slices BIO_BENCH_BIND_SLICES=1
run nobind BIO_BENCH_NO_BIND=1
run staged BIO_HOST_ZEROCOPY=0
run default BIO_X=0
