"""Host emulation of the planar program of the cooperative step kernel against the CPU oracle.

tests/emul/coop_emul.cu compiles the *same* __host__ __device__ phase functions the GPU kernel
runs (bioimitation_gym_b200/csrc/bio_coop_planar.cuh) for the host and executes them lane by
lane, phase by phase.  This checks the kernel's arithmetic without a GPU: dynamics evaluation
with and without the implicit contact/limit damping and with a perturbation force, on the states
of the golden trajectories.  Tolerance: fp64 1e-8 relative (measured 2e-10)."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "coop_emul.cu")
LIB = os.path.join(HERE, "emul", "_build", "libcoop_emul.so")
CSRC = os.path.join(os.path.dirname(HERE), "bioimitation_gym_b200", "csrc")

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")


@pytest.fixture(scope="module")
def emul_lib():
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["nvcc", "-std=c++17", "-O2", "-shared", "-Xcompiler", "-fPIC", "-gencode",
                               "arch=compute_100a,code=sm_100a", "-o", LIB, SRC])
    return ctypes.CDLL(LIB)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _pad(x, n):
    out = np.zeros(n)
    x = np.asarray(x, dtype=np.float64).ravel()
    out[:x.size] = x
    return out


def _emul(L, t, q, u, act, lm, ctrl, prec=1, h_imp=0.0, ext_fx=0.0, ext_pt=-1, vn0=None, iters=30):
    nd, nm = t.n_dof, t.n_muscles
    udot, adot, lmdot, misc = np.zeros(16), np.zeros(24), np.zeros(24), np.zeros(256)
    q_, u_, a_, l_, c_ = _pad(q, 16), _pad(u, 16), _pad(act, 24), _pad(lm, 24), _pad(ctrl, 24)
    v_ = _pad(vn0, 24) if vn0 is not None else None
    rc = L.emul_planar_eval(ctypes.byref(t), prec, iters, _p(q_), _p(u_), _p(a_), _p(l_), _p(c_),
                            _p(v_) if v_ is not None else None, ctypes.c_double(h_imp), ctypes.c_double(ext_fx), ext_pt, _p(udot), _p(adot),
                            _p(lmdot), _p(misc))
    assert rc == 0, "the model does not match the planar program (rc %d)" % rc
    return dict(udot=udot[:nd], adot=adot[:nm], lmdot=lmdot[:nm], com_pos=misc[0:3], com_vel=misc[3:6],
                contact=misc[6:18].reshape(2, 6), max_limit=misc[18], fiber_force=misc[19:19 + nm])


def _oracle_udot(orc, t, q, u, act, lm, ctrl, h_imp, ext, ext_pt):
    ev = orc.OrcEval()
    q_, u_, a_, l_, c_, f_ = _pad(q, 16), _pad(u, 16), _pad(act, 24), _pad(lm, 24), _pad(ctrl, 24), _pad(ext, 3)
    orc.lib().orc_eval_h(ctypes.byref(t), 20, _p(q_), _p(u_), _p(a_), _p(l_), _p(c_), _p(f_), int(ext_pt),
                         ctypes.c_double(h_imp), ctypes.byref(ev))
    return np.ctypeslib.as_array(ev.udot)[:t.n_dof].copy()


CASES = [("2d_muscle", "config1_muscle_walking_2d.npz"), ("2d_torque", "torque_walking_2d.npz")]


@pytest.mark.parametrize("key,gfile", CASES)
def test_emulated_evaluation_matches_oracle(emul_lib, oracle_lib, models, key, gfile):
    orc = oracle_lib
    t = models[key].tables
    g = np.load(os.path.join(HERE, "golden", gfile))
    worst = {}
    for k in range(0, g["q"].shape[0], 7):
        q, u, act, lm = g["q"][k], g["u"][k], g["act"][k], g["lm"][k]
        ctrl = np.random.default_rng(k).uniform(-50, 50, t.n_act) if t.is_torque else np.clip(g["action"][k], 0, 1)
        o = orc.eval_dynamics(t, q, u, act, lm, ctrl)
        e = _emul(emul_lib, t, q, u, act, lm, ctrl)
        for name, scale in (("udot", 1.0), ("adot", 1.0), ("lmdot", 1e-2), ("com_pos", 1.0), ("com_vel", 1.0),
                            ("contact", 100.0), ("fiber_force", 1.0)):
            a, b = np.asarray(o[name]), np.asarray(e[name])
            if a.size:
                worst[name] = max(worst.get(name, 0.0), float(np.max(np.abs(a - b) / np.maximum(np.abs(a), scale))))
    print(key, worst)
    assert max(worst.values()) < 1e-8


@pytest.mark.parametrize("key,gfile", CASES)
def test_emulated_implicit_damping_and_perturbation(emul_lib, oracle_lib, models, key, gfile):
    orc = oracle_lib
    t = models[key].tables
    g = np.load(os.path.join(HERE, "golden", gfile))
    worst = worst32 = 0.0
    for k in range(0, g["q"].shape[0], 5):
        q, u, act, lm = g["q"][k], g["u"][k], g["act"][k], g["lm"][k]
        ctrl = np.random.default_rng(k).uniform(-50, 50, t.n_act) if t.is_torque else np.clip(g["action"][k], 0, 1)
        ext_pt = 0 if k % 2 else -1
        a = _oracle_udot(orc, t, q, u, act, lm, ctrl, 5e-4, [37.0, 0.0, 0.0], ext_pt)
        e = _emul(emul_lib, t, q, u, act, lm, ctrl, h_imp=5e-4, ext_fx=37.0, ext_pt=ext_pt)
        worst = max(worst, float(np.max(np.abs(a - e["udot"]) / np.maximum(np.abs(a), 1.0))))
        e32 = _emul(emul_lib, t, q, u, act, lm, ctrl, prec=0, h_imp=5e-4, ext_fx=37.0, ext_pt=ext_pt)
        # fp32: accelerations relative to the largest one of the state (the solve is ill-conditioned)
        worst32 = max(worst32, float(np.max(np.abs(a - e32["udot"])) / max(np.max(np.abs(a)), 1.0)))
    print(key, "fp64 %.2e fp32 %.2e" % (worst, worst32))
    assert worst < 1e-8
    assert worst32 < 3e-2


def test_streamed_and_compiled_muscle_paths_agree(emul_lib, oracle_lib, models, monkeypatch):
    """Phase C has two forms of the path geometry: the compiled one (constant length + live segments per
    variant of the conditional points) and the streaming pass over the path points that any model falls back
    to.  BIO_PLANAR_STREAM_PATHS=1 makes the host program take the fallback: both must match the oracle, also
    at coordinates on either side of every conditional point's range."""
    orc = oracle_lib
    t = models["2d_muscle"].tables
    g = np.load(os.path.join(HERE, "golden", "config1_muscle_walking_2d.npz"))
    rng = np.random.default_rng(11)
    for mode in ("0", "1"):
        monkeypatch.setenv("BIO_PLANAR_STREAM_PATHS", mode)
        worst = 0.0
        for k in range(0, g["q"].shape[0], 11):
            q, u, act, lm = g["q"][k].copy(), g["u"][k], g["act"][k], g["lm"][k]
            if k % 2:                       # sweep hip / knee angles across the conditional ranges
                q[3:] = rng.uniform(-2.7, 1.6, q.size - 3)
            ctrl = np.clip(g["action"][k], 0, 1)
            o = orc.eval_dynamics(t, q, u, act, lm, ctrl)
            e = _emul(emul_lib, t, q, u, act, lm, ctrl)
            for name, scale in (("udot", 1.0), ("lmdot", 1e-2), ("fiber_force", 1.0)):
                a, b = np.asarray(o[name]), np.asarray(e[name])
                worst = max(worst, float(np.max(np.abs(a - b) / np.maximum(np.abs(a), scale))))
        print("stream paths = %s: worst relative difference %.2e" % (mode, worst))
        assert worst < 1e-8


def test_warm_started_newton_finds_the_same_root(emul_lib, oracle_lib, models):
    """Any warm start of the fibre-velocity Newton iteration (the kernel keeps the root of the
    previous substep) must end on the root the oracle finds from zero."""
    orc = oracle_lib
    t = models["2d_muscle"].tables
    g = np.load(os.path.join(HERE, "golden", "config1_muscle_walking_2d.npz"))
    rng = np.random.default_rng(5)
    worst = 0.0
    for k in range(0, g["q"].shape[0], 9):
        q, u, act, lm = g["q"][k], g["u"][k], g["act"][k], g["lm"][k]
        ctrl = np.clip(g["action"][k], 0, 1)
        o = orc.eval_dynamics(t, q, u, act, lm, ctrl, newton_iters=60)
        for scale in (1e-14, 1e-6, 1e-3, 0.1, 1.0, 5.0):
            e = _emul(emul_lib, t, q, u, act, lm, ctrl, vn0=rng.uniform(-scale, scale, t.n_muscles), iters=30)
            worst = max(worst, float(np.max(np.abs(o["lmdot"] - e["lmdot"]) / np.maximum(np.abs(o["lmdot"]), 1e-2))))
    print("warm-started Newton, worst relative fibre-velocity difference %.2e" % worst)
    assert worst < 1e-8


def test_both_launch_shapes_fit_the_shared_memory_of_one_sm(emul_lib):
    """Shared-memory budget of the fp32 instantiations: one CTA per SM holds the model block and one work
    buffer per env of its threads; 2D: half a warp per env, 512 (16 warps) or 640 (20 warps) threads; 3D: a warp
    per env, 512 or 896 (28 warps) threads.  227 KB (232448 B) of dynamic shared memory per CTA on sm_100."""
    out = np.zeros(8, dtype=np.int64)
    emul_lib.emul_sizes(_p(out))
    model, work2d, work3d = int(out[0]), int(out[1]), int(out[2])
    base = (model + 15) // 16 * 16
    for threads in (512, 640):
        assert base + threads // 16 * work2d <= 232448, "2D fp32, %d threads: %d B" % (threads, base + threads // 16 * work2d)
    for threads in (512, 896):
        assert base + threads // 32 * work3d <= 232448, "3D fp32, %d threads: %d B" % (threads, base + threads // 32 * work3d)


def test_fast_paths_are_enabled_for_the_shipped_models(emul_lib, models):
    """Guards the host-built program: every 2D model of the reference gets the planar program with scan
    kinematics, every 3D model gets the chain lists (spatial scan kinematics); a silent fallback to the
    general path would only show up as a slower benchmark."""
    for key, cm in models.items():
        out = np.zeros(16, dtype=np.int32)
        emul_lib.emul_prog_info(ctypes.byref(cm.tables), _p(out))
        ok, scan_ok, chain_ok, n_br, n0, n1, n_tasks, n_src, path_ok, n_live, a2_cheap = out.tolist()[:11]
        assert chain_ok == 1 and n_br >= 1, key
        if key.startswith("2d"):
            assert ok == 1 and scan_ok == 1 and n_br == 2 and 1 <= n0 <= 8 and 1 <= n1 <= 8, (key, out)
            assert n_src >= cm.tables.n_spheres + cm.tables.n_muscles
            # compiled muscle paths (one live segment per muscle in the 2D gait models), cheap second round of phase A
            assert path_ok == 1 and n_live <= 1 and a2_cheap == 1, (key, out)
        else:
            assert ok == 0 and max(n0, n1) <= 16, (key, out)


def test_program_rejects_models_it_does_not_cover(emul_lib, models):
    t = models["3d_muscle"].tables
    udot, adot, lmdot, misc = np.zeros(16), np.zeros(24), np.zeros(24), np.zeros(256)
    z16, z24 = np.zeros(16), np.zeros(24)
    rc = emul_lib.emul_planar_eval(ctypes.byref(t), 1, 5, _p(z16), _p(z16), _p(z24), _p(z24), _p(z24), None,
                                   ctypes.c_double(0.0), ctypes.c_double(0.0), -1, _p(udot), _p(adot), _p(lmdot),
                                   _p(misc))
    assert rc == -1
