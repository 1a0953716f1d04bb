import sys, os, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from bioimitation_gym_b200 import backend
N=4096
env = backend.VecEnv("MuscleWalkingImitation2D-v0", dict(num_envs=N, seed=1)); env.reset()
pin = lambda *s, dt=torch.float32: torch.empty(s, dtype=dt).pin_memory()
a=pin(N,env.n_act); a.copy_(torch.rand(N,env.n_act))
o,r,d,t = pin(N,env.obs_dim), pin(N), pin(N,dt=torch.uint8), pin(N,env.n_terms)
an,on,rn,dn,tn = a.numpy(),o.numpy(),r.numpy(),d.numpy(),t.numpy()
for _ in range(10): env.step_host(an,on,rn,dn,tn)
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(300): env.step_host(an,on,rn,dn,tn)
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/300
print(os.environ.get("BIO_HOST_ZEROCOPY","default"), "step_host %.1f us -> %.2f M env-steps/s ; obs checksum %.6f" % (dt*1e6, N/dt/1e6, float(on.sum())))
