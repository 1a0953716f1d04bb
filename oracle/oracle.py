"""ctypes front-end of the CPU oracle (oracle/bio_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by the product
package.  PARITY UNPINNED (see bio_oracle.c header).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from bioimitation_gym_b200 import ctables as ct

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libbio_oracle.so")

M = ct.MACROS
MAXD, MAXM, MAXA = M["BIO_MAX_DOF"], M["BIO_MAX_MUSCLES"], M["BIO_MAX_ACT"]
c_d = ctypes.c_double


class OrcEval(ctypes.Structure):
    _fields_ = [
        ("udot", c_d * MAXD), ("adot", c_d * MAXM), ("lmdot", c_d * MAXM),
        ("tendon_force", c_d * MAXM), ("fiber_force", c_d * MAXM),
        ("active_fiber_force", c_d * MAXM), ("path_len", c_d * MAXM),
        ("path_vel", c_d * MAXM), ("contact", (c_d * 6) * 2),
        ("limit_force", c_d * M["BIO_MAX_LIMITS"]),
        ("mass_matrix", (c_d * MAXD) * MAXD), ("bias", c_d * MAXD),
        ("obs_pos", (c_d * 3) * M["BIO_MAX_OBSPTS"]),
        ("obs_vel", (c_d * 3) * M["BIO_MAX_OBSPTS"]),
        ("com_pos", c_d * 3), ("com_vel", c_d * 3)]


class OrcEnv(ctypes.Structure):
    _fields_ = [
        ("q", c_d * MAXD), ("u", c_d * MAXD), ("act", c_d * MAXM), ("lm", c_d * MAXM),
        ("last_action", c_d * MAXA),
        ("history", (c_d * MAXA) * M["BIO_MAX_HORIZON"]),
        ("old_px", c_d), ("ep_return", c_d),
        ("hist_pos", ctypes.c_int32), ("istep", ctypes.c_int32),
        ("first", ctypes.c_int32), ("ep_len", ctypes.c_int32),
        ("episode", ctypes.c_int64)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "bio_oracle.c")
    hdr = ct.HEADER
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def build_fast() -> str:
    """The benchmark build (-O3 -march=native, FMA contraction, OpenMP) for the host this runs on; compiled here
    when missing.  Never used by parity tests (its results differ from the parity build in the last bits)."""
    import hashlib
    try:
        flags = [ln for ln in open("/proc/cpuinfo") if ln.startswith(("flags", "model name"))][:2]
    except OSError:
        flags = []
    tag = hashlib.sha1("".join(flags).encode()).hexdigest()[:10]
    path = os.path.join(_HERE, "_build", "libbio_oracle_fast_%s.so" % tag)
    src = os.path.join(_HERE, "bio_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(src), os.path.getmtime(ct.HEADER)):
        subprocess.check_call(["make", "-C", _HERE, "fast", "TAG=" + tag], stdout=subprocess.DEVNULL)
    return path


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        # BIO_ORACLE_LIB: the op-counting build (oracle/count_flops.py) or the benchmark build (build_fast)
        L = ctypes.CDLL(os.environ.get("BIO_ORACLE_LIB", _LIB_PATH))
        L.orc_openmp.restype = ctypes.c_int
        for f in ("orc_sizeof_env", "orc_sizeof_eval", "orc_sizeof_model_tables",
                  "orc_sizeof_task_config", "orc_splitmix64", "orc_rand"):
            getattr(L, f).restype = ctypes.c_uint64
        L.orc_rand.argtypes = [ctypes.c_uint64] * 4
        L.orc_equilibrium_lm.restype = c_d
        L.orc_equilibrium_lm.argtypes = [ctypes.c_void_p, ctypes.c_int, c_d, c_d]
        L.orc_obs_dim.restype = ctypes.c_int
        L.orc_step_env.restype = ctypes.c_int
        assert L.orc_sizeof_env() == ctypes.sizeof(OrcEnv), "OrcEnv mirror out of date"
        assert L.orc_sizeof_eval() == ctypes.sizeof(OrcEval), "OrcEval mirror out of date"
        assert L.orc_sizeof_model_tables() == ctypes.sizeof(ct.BioModelTables)
        assert L.orc_sizeof_task_config() == ctypes.sizeof(ct.BioTaskConfig)
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _vec(x, n):
    a = np.zeros(n, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    a[:x.size] = x
    return a


def func_eval(tables, f, x):
    y, d1, d2 = c_d(), c_d(), c_d()
    lib().orc_func_eval(ctypes.byref(tables), int(f), c_d(x), ctypes.byref(y),
                        ctypes.byref(d1), ctypes.byref(d2))
    return y.value, d1.value, d2.value


def curve_eval(tables, c, x):
    y, d = c_d(), c_d()
    lib().orc_curve_eval(ctypes.byref(tables), int(c), c_d(x), ctypes.byref(y), ctypes.byref(d))
    return y.value, d.value


def path_lengths(tables, q, u=None):
    q = _vec(q, MAXD)
    u = _vec(u if u is not None else [], MAXD)
    L = np.zeros(MAXM)
    Ld = np.zeros(MAXM)
    lib().orc_path_lengths(ctypes.byref(tables), _p(q), _p(u), _p(L), _p(Ld))
    n = tables.n_muscles
    return L[:n].copy(), Ld[:n].copy()


def equilibrium_lm(tables, i, L, a):
    return lib().orc_equilibrium_lm(ctypes.byref(tables), int(i), c_d(L), c_d(a))


def eval_dynamics(tables, q, u, act=None, lm=None, ctrl=None, newton_iters=20,
                  ext_force=None, ext_pt=-1):
    """One RHS evaluation; returns a dict of numpy arrays."""
    nd, nm, na = tables.n_dof, tables.n_muscles, tables.n_act
    ev = OrcEval()
    q = _vec(q, MAXD)
    u = _vec(u, MAXD)
    act = _vec(act if act is not None else [], MAXM)
    lm = _vec(lm if lm is not None else [], MAXM)
    ctrl = _vec(ctrl if ctrl is not None else [], MAXA)
    fx = _vec(ext_force, 3) if ext_force is not None else None
    lib().orc_eval(ctypes.byref(tables), int(newton_iters), _p(q), _p(u), _p(act), _p(lm),
                   _p(ctrl), _p(fx), int(ext_pt), ctypes.byref(ev))
    A = np.ctypeslib.as_array
    return dict(
        udot=A(ev.udot)[:nd].copy(), adot=A(ev.adot)[:nm].copy(), lmdot=A(ev.lmdot)[:nm].copy(),
        tendon_force=A(ev.tendon_force)[:nm].copy(), fiber_force=A(ev.fiber_force)[:nm].copy(),
        active_fiber_force=A(ev.active_fiber_force)[:nm].copy(),
        path_len=A(ev.path_len)[:nm].copy(), path_vel=A(ev.path_vel)[:nm].copy(),
        contact=A(ev.contact).copy(), limit_force=A(ev.limit_force)[:tables.n_limits].copy(),
        mass_matrix=A(ev.mass_matrix)[:nd, :nd].copy(), bias=A(ev.bias)[:nd].copy(),
        obs_pos=A(ev.obs_pos)[:tables.n_obspts].copy(), obs_vel=A(ev.obs_vel)[:tables.n_obspts].copy(),
        com_pos=A(ev.com_pos).copy(), com_vel=A(ev.com_vel).copy())


class RefTables:
    """Keeps the numpy arrays alive next to the BioRefTables that points at them."""

    def __init__(self, q, u, body_pos, com_pos):
        self.q = np.ascontiguousarray(q, dtype=np.float64)
        self.u = np.ascontiguousarray(u, dtype=np.float64)
        self.body_pos = np.ascontiguousarray(body_pos, dtype=np.float64)
        self.com_pos = np.ascontiguousarray(com_pos, dtype=np.float64)
        s = ct.BioRefTables()
        s.n_rows = self.q.shape[0]
        s.n_coords = self.q.shape[1]
        s.n_refbodies = self.body_pos.shape[1]
        s.q = self.q.ctypes.data
        s.u = self.u.ctypes.data
        s.body_pos = self.body_pos.ctypes.data
        s.com_pos = self.com_pos.ctypes.data
        self.struct = s


class OracleVecEnv:
    """Batch of oracle envs with the same call pattern as the CUDA VecEnv."""

    def __init__(self, tables, task, ref: RefTables, n_envs, seed=0, env_offset=0, threads=1):
        self.L = lib()
        self.tables, self.task, self.ref = tables, task, ref
        self.n = int(n_envs)
        self.seed, self.env_offset = int(seed), int(env_offset)
        self.envs = (OrcEnv * self.n)()
        self.obs_dim = self.L.orc_obs_dim(ctypes.byref(tables), ctypes.byref(task))
        self.n_act = tables.n_act
        self.n_terms = task.n_reward_terms
        self.threads = max(1, int(threads))
        self._pool = ThreadPoolExecutor(self.threads) if self.threads > 1 else None
        # like bio_create: every env starts from a reference-state reset of episode 0
        self._reset(bump=0)

    def _slices(self):
        if self.threads == 1:
            return [(0, self.n)]
        step = (self.n + self.threads - 1) // self.threads
        return [(i, min(self.n, i + step)) for i in range(0, self.n, step)]

    def reset(self):
        """Explicit reset = next episode of every env (same as bio_reset)."""
        return self._reset(bump=1)

    def _reset(self, bump):
        obs = np.zeros((self.n, self.obs_dim))
        sz = ctypes.sizeof(OrcEnv)
        base = ctypes.addressof(self.envs)

        def run(sl):
            i0, i1 = sl
            self.L.orc_batch_reset(ctypes.byref(self.tables), ctypes.byref(self.task),
                                   ctypes.byref(self.ref.struct), ctypes.c_void_p(base + i0 * sz),
                                   i1 - i0, ctypes.c_uint64(self.seed),
                                   ctypes.c_int64(self.env_offset + i0), _p(obs[i0:i1]), int(bump))
        if self._pool:
            list(self._pool.map(run, self._slices()))
        else:
            run((0, self.n))
        return obs

    def step(self, actions):
        actions = np.ascontiguousarray(actions, dtype=np.float64).reshape(self.n, self.n_act)
        obs = np.zeros((self.n, self.obs_dim))
        rew = np.zeros(self.n)
        done = np.zeros(self.n, dtype=np.uint8)
        terms = np.zeros((self.n, self.n_terms))
        reasons = np.zeros(self.n, dtype=np.int32)
        sz = ctypes.sizeof(OrcEnv)
        base = ctypes.addressof(self.envs)

        def run(sl):
            i0, i1 = sl
            self.L.orc_batch_step(ctypes.byref(self.tables), ctypes.byref(self.task),
                                  ctypes.byref(self.ref.struct), ctypes.c_void_p(base + i0 * sz),
                                  i1 - i0, ctypes.c_uint64(self.seed),
                                  ctypes.c_int64(self.env_offset + i0), _p(actions[i0:i1]),
                                  _p(obs[i0:i1]), _p(rew[i0:i1]), _p(done[i0:i1]),
                                  _p(terms[i0:i1]), _p(reasons[i0:i1]))
        if self.threads > 1 and self.L.orc_openmp():
            # benchmark build: one call, OpenMP over chunks of envs inside the library
            self.L.orc_batch_step_mt(ctypes.byref(self.tables), ctypes.byref(self.task), ctypes.byref(self.ref.struct),
                                     ctypes.c_void_p(base), self.n, ctypes.c_uint64(self.seed),
                                     ctypes.c_int64(self.env_offset), _p(actions), _p(obs), _p(rew), _p(done), _p(terms),
                                     _p(reasons), int(self.threads))
        elif self._pool:
            list(self._pool.map(run, self._slices()))
        else:
            run((0, self.n))
        return obs, rew, done, terms, reasons

    def step_env_debug(self, i, action):
        """Step env i alone (no auto-reset) and also return the post-step RHS evaluation."""
        ev = OrcEval()
        obs = np.zeros(self.obs_dim)
        rew = c_d()
        terms = np.zeros(8)
        a = _vec(action, MAXA)
        r = self.L.orc_step_env(ctypes.byref(self.tables), ctypes.byref(self.task),
                                ctypes.byref(self.ref.struct), ctypes.byref(self.envs[i]),
                                ctypes.c_uint64(self.seed), ctypes.c_uint64(self.env_offset + i),
                                _p(a), _p(obs), ctypes.byref(rew), _p(terms), ctypes.byref(ev))
        return obs, rew.value, r, terms[:self.n_terms].copy(), ev

    # state access (numpy views per env field)
    def get_state(self):
        nd, nm, na = self.tables.n_dof, self.tables.n_muscles, self.tables.n_act
        h = self.task.horizon
        out = dict(q=np.zeros((self.n, nd)), u=np.zeros((self.n, nd)), act=np.zeros((self.n, nm)),
                   lm=np.zeros((self.n, nm)), last_action=np.zeros((self.n, na)),
                   history=np.zeros((self.n, h, na)), old_px=np.zeros(self.n),
                   istep=np.zeros(self.n, dtype=np.int32), first=np.zeros(self.n, dtype=np.int32),
                   hist_pos=np.zeros(self.n, dtype=np.int32), episode=np.zeros(self.n, dtype=np.int64))
        A = np.ctypeslib.as_array
        for i in range(self.n):
            e = self.envs[i]
            out["q"][i] = A(e.q)[:nd]
            out["u"][i] = A(e.u)[:nd]
            out["act"][i] = A(e.act)[:nm]
            out["lm"][i] = A(e.lm)[:nm]
            out["last_action"][i] = A(e.last_action)[:na]
            hist = A(e.history)[:h, :na]
            # ring buffer -> chronological order is irrelevant for the mean; keep raw
            out["history"][i] = hist
            out["old_px"][i] = e.old_px
            out["istep"][i] = e.istep
            out["first"][i] = e.first
            out["hist_pos"][i] = e.hist_pos
            out["episode"][i] = e.episode
        return out

    def set_state(self, st):
        A = np.ctypeslib.as_array
        for i in range(self.n):
            e = self.envs[i]
            for k in ("q", "u", "act", "lm", "last_action"):
                if k in st:
                    v = np.asarray(st[k][i], dtype=np.float64)
                    A(getattr(e, k))[:v.size] = v
            if "history" in st:
                hv = np.asarray(st["history"][i], dtype=np.float64)
                A(e.history)[:hv.shape[0], :hv.shape[1]] = hv
            if "old_px" in st:
                e.old_px = float(st["old_px"][i])
            if "istep" in st:
                e.istep = int(st["istep"][i])
            if "first" in st:
                e.first = int(st["first"][i])
            if "hist_pos" in st:
                e.hist_pos = int(st["hist_pos"][i])
            if "episode" in st:
                e.episode = int(st["episode"][i])
