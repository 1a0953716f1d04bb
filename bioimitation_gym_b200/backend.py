"""ctypes binding of libbio_b200.so and the batched ``VecEnv``.

``VecEnv`` is the vectorised form of the reference's ``OsimEnv``
(reference ``opensim_environment.py:35-113``): ``reset() -> obs[N, D]`` and
``step(actions[N, A]) -> (obs, reward[N], done[N], info)`` with
``info['all_rewards']`` = the reward terms (reference
``muscle_walking_imitation_env2D.py:358``).  Tensors are torch CUDA tensors
owned by the env and handed to the C ABI as raw device pointers (zero copy);
PyTorch is only the allocator / stream provider.

There is no CPU fallback: constructing a ``VecEnv`` without the compiled
library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Any, Dict, Mapping, Optional

import numpy as np

from . import ctables as ct
from . import registry, tasks

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libbio_b200.so")

_lib = None

C_API = {
    "bio_abi_version": (ctypes.c_int, []),
    "bio_sizeof_model_tables": (ctypes.c_uint64, []),
    "bio_sizeof_task_config": (ctypes.c_uint64, []),
    "bio_last_error": (ctypes.c_char_p, []),
    "bio_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                  ctypes.c_int32, ctypes.c_int32, ctypes.c_uint64, ctypes.c_int64,
                                  ctypes.POINTER(ctypes.c_void_p)]),
    "bio_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "bio_reset": (ctypes.c_int, [ctypes.c_void_p] * 4),
    "bio_step": (ctypes.c_int, [ctypes.c_void_p] * 7),
    "bio_step_host": (ctypes.c_int, [ctypes.c_void_p] * 6),
    "bio_reset_host": (ctypes.c_int, [ctypes.c_void_p] * 3),
    "bio_step_host_begin": (ctypes.c_int, [ctypes.c_void_p] * 6),
    "bio_step_host_end": (ctypes.c_int, [ctypes.c_void_p]),
    "bio_set_grid": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    "bio_groups_run": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64,
                                      ctypes.c_void_p, ctypes.c_void_p]),
    "bio_get_state": (ctypes.c_int, [ctypes.c_void_p] * 3),
    "bio_set_state": (ctypes.c_int, [ctypes.c_void_p] * 3),
    "bio_eval_debug": (ctypes.c_int, [ctypes.c_void_p] * 4),
    "bio_set_step_extra": (ctypes.c_int, [ctypes.c_void_p] * 2),
    "bio_set_host_stream": (ctypes.c_int, [ctypes.c_void_p] * 2),
    "bio_id_apply": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32] + [ctypes.c_void_p] * 5),
    "bio_id_multiply_m": (ctypes.c_int, [ctypes.c_void_p] * 4),
    "bio_id_multiply_minv": (ctypes.c_int, [ctypes.c_void_p] * 4),
    "bio_id_residual": (ctypes.c_int, [ctypes.c_void_p] * 5),
    "bio_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    "bio_kernel_shape": (ctypes.c_int, [ctypes.c_void_p] * 4),
    "bio_launch_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "bio_obs_dim": (ctypes.c_int32, [ctypes.c_void_p]),
    "bio_n_act": (ctypes.c_int32, [ctypes.c_void_p]),
    "bio_measure_fp32_peak": (ctypes.c_double, [ctypes.c_int32]),
    "bio_measure_fp64_peak": (ctypes.c_double, [ctypes.c_int32]),
}


POLICY_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64)


def load_library(path: Optional[str] = None):
    """dlopen the CUDA library and declare every symbol of include/bio_b200.h."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). This backend has no CPU fallback." % p)
    lib = ctypes.CDLL(p)
    for name, (res, args) in C_API.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    ct.check_abi(lib)
    if path is None:
        _lib = lib
    return lib


class BioError(RuntimeError):
    pass


def _check(lib, rc: int, what: str):
    if rc != 0:
        msg = lib.bio_last_error()
        raise BioError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class VecEnv:
    """N independent envs of one env ID on one CUDA device."""

    def __init__(self, env_id: str, config: Optional[Mapping[str, Any]] = None, model=None, ref=None):
        import torch
        self.torch = torch
        self.lib = load_library()
        cfg = tasks.merged_config(config)
        self.env_id = env_id
        self.config = cfg
        self.spec, self.cm, self.ref, self.task = registry.build_env_tables(env_id, cfg, model, ref)
        if not torch.cuda.is_available():
            raise BioError("no CUDA device available: the batched physics step has no CPU fallback")
        self.num_envs = int(cfg["num_envs"])
        self.device_index = int(cfg["device"]) if not isinstance(cfg["device"], str) else \
            torch.device(cfg["device"]).index or 0
        self.device = torch.device("cuda", self.device_index)
        dt = str(cfg["dtype"]).replace("torch.", "")
        if dt in ("float32", "fp32", "f32"):
            self.dtype, prec = torch.float32, ct.MACROS["BIO_PREC_F32"]
        elif dt in ("float64", "fp64", "f64", "double"):
            self.dtype, prec = torch.float64, ct.MACROS["BIO_PREC_F64"]
        else:
            raise ValueError("dtype must be float32 or float64")
        t = self.cm.tables
        self.n_act, self.n_dof, self.n_muscles = t.n_act, t.n_dof, t.n_muscles
        self.obs_dim = self.task.obs_dim
        self.n_terms = self.task.n_reward_terms
        self.action_low = np.ctypeslib.as_array(t.act_min)[:self.n_act].copy()
        self.action_high = np.ctypeslib.as_array(t.act_max)[:self.n_act].copy()
        if self.spec.torque:
            # the action space of a torque env is the actuator range (opensim_wrapper.py:39-53)
            pass
        self._ref_keep = dict(q=np.ascontiguousarray(self.ref["q"], dtype=np.float64),
                              u=np.ascontiguousarray(self.ref["u"], dtype=np.float64),
                              body_pos=np.ascontiguousarray(self.ref["body_pos"], dtype=np.float64),
                              com_pos=np.ascontiguousarray(self.ref["com_pos"], dtype=np.float64))
        r = ct.BioRefTables()
        r.n_rows, r.n_coords = self._ref_keep["q"].shape
        r.n_refbodies = self._ref_keep["body_pos"].shape[1]
        r.q = self._ref_keep["q"].ctypes.data
        r.u = self._ref_keep["u"].ctypes.data
        r.body_pos = self._ref_keep["body_pos"].ctypes.data
        r.com_pos = self._ref_keep["com_pos"].ctypes.data
        h = ctypes.c_void_p()
        rc = self.lib.bio_create(ctypes.addressof(self.cm.tables), ctypes.addressof(self.task),
                                 ctypes.addressof(r), self.num_envs, self.device_index, prec,
                                 int(cfg["seed"]), int(cfg["env_offset"]), ctypes.byref(h))
        _check(self.lib, rc, "bio_create")
        self.handle = h
        N = self.num_envs
        with torch.cuda.device(self.device):
            self.obs = torch.zeros((N, self.obs_dim), dtype=self.dtype, device=self.device)
            self.reward = torch.zeros((N,), dtype=self.dtype, device=self.device)
            self.done = torch.zeros((N,), dtype=torch.uint8, device=self.device)
            self.terms = torch.zeros((N, self.n_terms), dtype=self.dtype, device=self.device)
            self._stats = torch.zeros((16,), dtype=torch.float64, device=self.device)

    # ------------------------------------------------------------------ API
    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self, mask=None):
        """Reference-state reset of all envs (or those with mask != 0)."""
        m = None
        if mask is not None:
            m = mask.to(device=self.device, dtype=self.torch.uint8).contiguous()
        _check(self.lib, self.lib.bio_reset(self.handle, _ptr(m), _ptr(self.obs), self._stream()), "bio_reset")
        return self.obs

    def step(self, actions):
        a = actions
        if a.device != self.device or a.dtype != self.dtype or not a.is_contiguous():
            a = a.to(device=self.device, dtype=self.dtype).contiguous()
        if tuple(a.shape) != (self.num_envs, self.n_act):
            raise ValueError("actions must have shape (%d, %d)" % (self.num_envs, self.n_act))
        _check(self.lib, self.lib.bio_step(self.handle, _ptr(a), _ptr(self.obs), _ptr(self.reward),
                                           _ptr(self.done), _ptr(self.terms), self._stream()), "bio_step")
        return self.obs, self.reward, self.done, {"all_rewards": self.terms}

    def step_host(self, actions: np.ndarray, obs: np.ndarray, reward: np.ndarray, done: np.ndarray,
                  terms: Optional[np.ndarray] = None):
        """Host-buffer call (numpy / pinned arrays in the env's dtype): H2D, step, D2H, sync."""
        _check(self.lib, self.lib.bio_step_host(
            self.handle, actions.ctypes.data, obs.ctypes.data, reward.ctypes.data, done.ctypes.data,
            terms.ctypes.data if terms is not None else None), "bio_step_host")

    def step_host_begin(self, actions: np.ndarray, obs: np.ndarray, reward: np.ndarray, done: np.ndarray,
                        terms: np.ndarray):
        """First half of step_host: launch and return (page-locked buffers only); step_host_end waits."""
        _check(self.lib, self.lib.bio_step_host_begin(
            self.handle, actions.ctypes.data, obs.ctypes.data, reward.ctypes.data, done.ctypes.data,
            terms.ctypes.data), "bio_step_host_begin")

    def step_host_end(self):
        _check(self.lib, self.lib.bio_step_host_end(self.handle), "bio_step_host_end")

    def set_grid(self, ctas: int):
        """Persistent grid of the step kernel (bio_set_grid): `ctas` SMs instead of the whole device."""
        _check(self.lib, self.lib.bio_set_grid(self.handle, int(ctas)), "bio_set_grid")

    def _host_buffers(self):
        """Two sets of page-locked buffers, used alternately: the arrays a call returns stay valid until the
        call after the next one (the reference hands out a fresh array per step)."""
        if getattr(self, "_hb", None) is None:
            torch = self.torch
            pin = lambda *s, dt=None: torch.zeros(s, dtype=dt or self.dtype).pin_memory()
            N = self.num_envs
            self._hb = [dict(a=pin(N, self.n_act), o=pin(N, self.obs_dim), r=pin(N), d=pin(N, dt=torch.uint8),
                             t=pin(N, self.n_terms)) for _ in range(2)]
            for b in self._hb:
                b["np"] = {k: v.numpy() for k, v in b.items()}
            self._hb_turn = 0
        self._hb_turn ^= 1
        return self._hb[self._hb_turn]["np"]

    def step_np(self, actions):
        """numpy in, numpy out (CPU-side trainers): the step kernel reads the actions from and writes obs /
        reward / done / terms into page-locked host memory in place (bio_step_host)."""
        b = self._host_buffers()
        a = np.asarray(actions)
        if a.shape != (self.num_envs, self.n_act):
            raise ValueError("actions must have shape (%d, %d)" % (self.num_envs, self.n_act))
        b["a"][...] = a
        self.step_host(b["a"], b["o"], b["r"], b["d"], b["t"])
        return b["o"], b["r"], b["d"].view(np.bool_), {"all_rewards": b["t"]}

    def reset_np(self):
        b = self._host_buffers()
        self.reset_host(b["o"])
        return b["o"]

    def reset_host(self, obs: np.ndarray, mask: Optional[np.ndarray] = None):
        _check(self.lib, self.lib.bio_reset_host(
            self.handle, mask.ctypes.data if mask is not None else None, obs.ctypes.data), "bio_reset_host")

    # ---------------------------------------------------------------- state
    def _state_struct(self, tensors: Dict[str, Any]):
        p = ct.BioStatePtrs()
        for k, v in tensors.items():
            setattr(p, k, v.data_ptr())
        return p

    def get_state(self):
        torch = self.torch
        N, H = self.num_envs, self.task.horizon
        mk = lambda *s: torch.zeros(s, dtype=self.dtype, device=self.device)
        out = dict(q=mk(N, self.n_dof), u=mk(N, self.n_dof), act=mk(N, self.n_muscles), lm=mk(N, self.n_muscles),
                   last_action=mk(N, self.n_act), history=mk(N, H, self.n_act), old_px=mk(N),
                   istep=torch.zeros(N, dtype=torch.int32, device=self.device),
                   first=torch.zeros(N, dtype=torch.int32, device=self.device),
                   hist_pos=torch.zeros(N, dtype=torch.int32, device=self.device),
                   episode=torch.zeros(N, dtype=torch.int64, device=self.device))
        p = self._state_struct(out)
        _check(self.lib, self.lib.bio_get_state(self.handle, ctypes.addressof(p), self._stream()), "bio_get_state")
        return out

    def set_state(self, state: Dict[str, Any]):
        torch = self.torch
        keep = {}
        for k, v in state.items():
            want = torch.int32 if k in ("istep", "first", "hist_pos") else \
                (torch.int64 if k == "episode" else self.dtype)
            keep[k] = torch.as_tensor(v).to(device=self.device, dtype=want).contiguous()
        p = self._state_struct(keep)
        _check(self.lib, self.lib.bio_set_state(self.handle, ctypes.addressof(p), self._stream()), "bio_set_state")
        torch.cuda.current_stream(self.device).synchronize()

    def eval_debug(self, controls=None):
        """One dynamics evaluation at the current state (no integration)."""
        torch = self.torch
        N, nd, nm = self.num_envs, self.n_dof, self.n_muscles
        mk = lambda *s: torch.zeros(s, dtype=self.dtype, device=self.device)
        out = dict(udot=mk(N, nd), tendon_force=mk(N, nm), fiber_force=mk(N, nm), fiber_vel=mk(N, nm),
                   act_dot=mk(N, nm), path_len=mk(N, nm), path_vel=mk(N, nm), contact=mk(N, 2, 6),
                   limit_force=mk(N, self.cm.tables.n_limits), mass_matrix=mk(N, nd, nd), bias=mk(N, nd))
        p = ct.BioDebugPtrs()
        for k, v in out.items():
            if v.numel():
                setattr(p, k, v.data_ptr())
        c = None
        if controls is not None:
            c = controls.to(device=self.device, dtype=self.dtype).contiguous()
        _check(self.lib, self.lib.bio_eval_debug(self.handle, _ptr(c), ctypes.addressof(p), self._stream()),
               "bio_eval_debug")
        return out

    def enable_step_extra(self, *names):
        """Attach extra per-step outputs of the step kernel (BioStepExtra in the header): any of
        `terminal_obs`, `done_reason` (what a replay buffer needs from an auto-resetting batch) and the
        dynamics read-outs `udot`, `tendon_force`, `fiber_force`, `fiber_vel`, `contact`, `limit_force` of the
        end-of-step evaluation.  Returns the dict of env-owned tensors the following steps fill; no names:
        detach."""
        torch = self.torch
        N, nd, nm = self.num_envs, self.n_dof, self.n_muscles
        mk = lambda *s: torch.zeros(s, dtype=self.dtype, device=self.device)
        shapes = dict(terminal_obs=(N, self.obs_dim), udot=(N, nd), tendon_force=(N, nm), fiber_force=(N, nm),
                      fiber_vel=(N, nm), contact=(N, 2, 6), limit_force=(N, self.cm.tables.n_limits))
        out = {}
        p = ct.BioStepExtra()
        for k in names:
            if k == "done_reason":
                out[k] = torch.zeros(N, dtype=torch.int32, device=self.device)
            elif k in shapes:
                out[k] = mk(*shapes[k])
            else:
                raise KeyError("unknown step output %r" % k)
            if out[k].numel():
                setattr(p, k, out[k].data_ptr())
        self.extra = out
        _check(self.lib, self.lib.bio_set_step_extra(self.handle, ctypes.addressof(p) if names else None),
               "bio_set_step_extra")
        return out

    def id_apply(self, op: str, x, controls=None, shift=None):
        """Inverse-dynamics operator set at the current state (bio_id_apply): op in 'multiply_m',
        'multiply_minv', 'residual', 'solve_shifted'; x [N, n_dof] -> [N, n_dof]."""
        torch = self.torch
        code = ct.MACROS["BIO_ID_" + op.upper()]
        prep = lambda t: None if t is None else t.to(device=self.device, dtype=self.dtype).contiguous()
        xx, cc, ss = prep(x), prep(controls), prep(shift)
        if tuple(xx.shape) != (self.num_envs, self.n_dof):
            raise ValueError("x must have shape (%d, %d)" % (self.num_envs, self.n_dof))
        out = torch.empty_like(xx)
        _check(self.lib, self.lib.bio_id_apply(self.handle, code, _ptr(xx), _ptr(cc), _ptr(ss), _ptr(out),
                                               self._stream()), "bio_id_apply")
        return out

    def stats(self, reset: bool = False):
        """Device tensor of 16 float64 rollout statistics (see bio_stats in the header)."""
        _check(self.lib, self.lib.bio_stats(self.handle, _ptr(self._stats), int(reset), self._stream()), "bio_stats")
        return self._stats

    def coop_shape(self):
        """(size class, threads per CTA, resident CTAs per SM) of the step kernel (bio_kernel_shape)."""
        c, t, k = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        _check(self.lib, self.lib.bio_kernel_shape(self.handle, ctypes.byref(c), ctypes.byref(t), ctypes.byref(k)),
               "bio_kernel_shape")
        return c.value, t.value, k.value

    @property
    def launch_count(self) -> int:
        return int(self.lib.bio_launch_count(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.bio_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class EnvGroups:
    """The batch as G groups stepped in a pipeline (the send / recv pattern of asynchronous vector envs): group g
    holds envs [g N/G, (g+1) N/G) of the same seeded batch (the reset RNG is keyed by the global env index, so
    the union equals one VecEnv of N envs) and steps on its own N_SM / G SMs, so that while one group's
    observation rows drain over PCIe and its caller picks the next actions, the other groups keep computing.

        groups = EnvGroups("MuscleWalkingImitation2D-v0", dict(num_envs=4096), groups=4)
        obs = groups.reset()                       # list of G arrays [N/G, D] (page-locked host memory)
        for g in range(groups.G): groups.send(g, policy(obs[g]))
        while training:
            for g in range(groups.G):
                o, r, d, info = groups.recv(g)     # waits for group g only
                groups.send(g, policy(o))
    """

    def __init__(self, env_id: str, config: Optional[Mapping[str, Any]] = None, groups: int = 4):
        import torch
        cfg = dict(tasks.merged_config(config))
        n = int(cfg["num_envs"])
        if groups < 1 or n % groups:
            raise ValueError("num_envs must be a multiple of the number of groups")
        self.G, self.n_group = groups, n // groups
        self.envs = []
        off0 = int(cfg["env_offset"])
        for g in range(groups):
            c = dict(cfg)
            c["num_envs"] = self.n_group
            c["env_offset"] = off0 + g * self.n_group
            self.envs.append(VecEnv(env_id, c))
        e0 = self.envs[0]
        sms = torch.cuda.get_device_properties(e0.device).multi_processor_count
        if groups > 1:
            for e in self.envs:
                e.set_grid(max(1, sms // groups))
        pin = lambda *s, dt=None: torch.zeros(s, dtype=dt or e0.dtype).pin_memory()
        self._pinned = [dict(a=pin(self.n_group, e0.n_act), o=pin(self.n_group, e0.obs_dim), r=pin(self.n_group),
                             d=pin(self.n_group, dt=torch.uint8), t=pin(self.n_group, e0.n_terms)) for _ in range(groups)]
        self.buf = [{k: v.numpy() for k, v in b.items()} for b in self._pinned]
        self._in_flight = [False] * groups
        self.n_act, self.obs_dim, self.num_envs = e0.n_act, e0.obs_dim, n

    def reset(self):
        for g, e in enumerate(self.envs):
            if self._in_flight[g]:
                e.step_host_end()
                self._in_flight[g] = False
            e.reset_host(self.buf[g]["o"])
        self.envs[0].torch.cuda.synchronize(self.envs[0].device)
        return [b["o"] for b in self.buf]

    def send(self, g: int, actions):
        """Start the next control step of group g with `actions` [N/G, A] (copied into the group's page-locked
        action buffer; pass the buffer itself, groups.buf[g]['a'], to skip the copy)."""
        if self._in_flight[g]:
            raise BioError("group %d is already stepping: recv() it first" % g)
        b = self.buf[g]
        if actions is not b["a"]:
            b["a"][...] = actions
        self.envs[g].step_host_begin(b["a"], b["o"], b["r"], b["d"], b["t"])
        self._in_flight[g] = True

    def recv(self, g: int):
        """Wait for group g's step: (obs, reward, done, info) views of its page-locked buffers, valid until the
        group's next send()."""
        if not self._in_flight[g]:
            raise BioError("group %d has no step in flight" % g)
        self.envs[g].step_host_end()
        self._in_flight[g] = False
        b = self.buf[g]
        return b["o"], b["r"], b["d"].view(np.bool_), {"all_rewards": b["t"]}

    def run(self, steps: int, policy=None, action_ring=None):
        """`steps` control steps of every group through the native sampler loop (bio_groups_run): a group's next step
        is launched the moment its previous one has landed in its page-locked buffers, with no Python between the
        launches.  `policy(g, k)`, if given, is called before step k of group g to fill self.buf[g]['a'] from
        self.buf[g]['o']; else `action_ring` = list over groups of lists of page-locked torch tensors [N/G, A] used
        round-robin (a replayed action sequence); else every step re-uses self.buf[g]['a'].  Results of the last step
        are in self.buf[g]; bit-identical to the same steps through send / recv."""
        for g in range(self.G):
            if self._in_flight[g]:
                raise BioError("group %d is stepping: recv() it first" % g)
        e0 = self.envs[0]
        handles = (ctypes.c_void_p * self.G)(*[e.handle.value for e in self.envs])
        bufs = (ct.BioGroupBuffers * self.G)()
        keep = []
        for g in range(self.G):
            b = self.buf[g]
            bufs[g].actions, bufs[g].obs, bufs[g].reward = b["a"].ctypes.data, b["o"].ctypes.data, b["r"].ctypes.data
            bufs[g].done, bufs[g].reward_terms = b["d"].ctypes.data, b["t"].ctypes.data
            if policy is None and action_ring is not None:
                ring = action_ring[g]
                for t in ring:
                    if not t.is_pinned() or tuple(t.shape) != (self.n_group, self.n_act) or not t.is_contiguous():
                        raise ValueError("action_ring entries must be page-locked contiguous [N/G, A] tensors")
                arr = (ctypes.c_void_p * len(ring))(*[t.data_ptr() for t in ring])
                keep.append(arr)
                bufs[g].action_ring = ctypes.addressof(arr)
                bufs[g].ring_len = len(ring)
        cb = POLICY_FN(lambda user, g, k: policy(int(g), int(k))) if policy is not None else None
        _check(e0.lib, e0.lib.bio_groups_run(ctypes.cast(handles, ctypes.c_void_p), self.G,
                                             ctypes.cast(bufs, ctypes.c_void_p), int(steps),
                                             ctypes.cast(cb, ctypes.c_void_p) if cb is not None else None, None),
               "bio_groups_run")

    def close(self):
        for g, e in enumerate(self.envs):
            if self._in_flight[g]:
                e.step_host_end()
            e.close()
        self.envs = []
