#!/bin/bash
# Cycles per phase of the planar program for a lone warp (dependency latency, no issue contention) and at the
# benchmark batch: profiling build (-DBIO_PHASE_CLOCK), run on the GPU box, then the normal build again.
#   tools/phase_clock.sh            (uses gpurun; prints the per-phase cycle counts)
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build(force=True, extra_flags=['-DBIO_PHASE_CLOCK'])"
/usr/local/graft/bin/gpurun --timeout 300 -- 'python - <<PY
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from bioimitation_gym_b200 import backend
for n in (296, 4096):
    env = backend.VecEnv("MuscleWalkingImitation2D-v0", dict(num_envs=n, seed=1)); env.reset()
    a = torch.rand((n, env.n_act), device=env.device)
    for _ in range(3): env.step(a)
    torch.cuda.synchronize(); print("---- above:", n, "envs (last line = one control step)", flush=True)
    env.close()
PY' 2>&1 | grep -E "^----|phase cycles|step cycles" | awk '/^----/{if(last)print last; print; last=""} /phase cycles/{last=$0} /step cycles/{last2=$0} /^----/{if(last2)print last2; last2=""}'
python -c "import __graft_entry__ as g; g.build(force=True)"
