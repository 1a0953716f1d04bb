"""Reference-motion tables without OpenSim.

The reference envs read four AnalyzeTool outputs per task
(``task_Kinematics_q.sto``, ``_u.sto``, ``task_BodyKinematics_pos_global.sto``,
``_vel_global.sto``; reference ``muscle_walking_imitation_env2D.py:46-71``)
through ``read_from_storage`` (reference ``opensim_utils.py:283-315``:
degrees -> radians, ``resampleLinear(0.01)``).  Those files are not shipped
(reference ``.gitignore:18``), so this module regenerates their content from
the inputs that are: the inverse-kinematics ``.mot`` files, with the recipe of
``walking_reference_data/setup_ka.xml`` (6 Hz low-pass of the coordinates,
speeds by differentiation, BodyKinematics = mass-centre positions of every
body in ground + whole-body centre of mass).

Host-side numpy only; runs once per task at asset-compile time or at env
construction, never per step.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np

from . import ctables as ct
from .model_compiler import CompiledModel, rot_axis


# --------------------------------------------------------------------------
# OpenSim storage (.sto / .mot) text files
# --------------------------------------------------------------------------
def read_storage(path: str):
    """Return (labels, data[T, ncol], in_degrees). Column 0 is time."""
    in_deg = False
    with open(path, "r") as fh:
        lines = fh.read().splitlines()
    i = 0
    while i < len(lines):
        s = lines[i].strip()
        if s.lower().startswith("indegrees"):
            in_deg = s.split("=")[1].strip().lower() in ("yes", "true")
        if s.lower() == "endheader":
            break
        i += 1
    labels = lines[i + 1].split()
    rows = [[float(x) for x in ln.split()] for ln in lines[i + 2:] if ln.strip()]
    return labels, np.asarray(rows, dtype=np.float64), in_deg


def write_storage(path: str, name: str, labels: Sequence[str], data: np.ndarray,
                  in_degrees: bool = False) -> None:
    """Write the layout OpenSim's Storage::print produces (used by the
    trajectory export, reference ``opensim_wrapper.py:334-338``)."""
    data = np.asarray(data)
    with open(path, "w") as fh:
        fh.write("%s\nversion=1\nnRows=%d\nnColumns=%d\ninDegrees=%s\nendheader\n"
                 % (name, data.shape[0], data.shape[1], "yes" if in_degrees else "no"))
        fh.write("\t".join(labels) + "\n")
        for r in data:
            fh.write("\t".join("%.8f" % v for v in r) + "\n")


def resample_linear(t: np.ndarray, x: np.ndarray, dt: float):
    """Storage::resampleLinear: uniform grid from t[0] to t[-1]."""
    n = int(math.floor((t[-1] - t[0]) / dt + 1e-9)) + 1
    tn = t[0] + dt * np.arange(n)
    out = np.stack([np.interp(tn, t, x[:, j]) for j in range(x.shape[1])], axis=1)
    return tn, out


def lowpass_zero_phase(x: np.ndarray, fs: float, fc: float) -> np.ndarray:
    """Zero-phase 2nd-order Butterworth low-pass (forward + backward pass)
    with odd-reflection padding; columns are filtered independently."""
    if fc <= 0:
        return x.copy()
    w = math.tan(math.pi * fc / fs)
    # correct the cut-off for the double pass
    w = w / (2 ** 0.5 - 1) ** 0.25
    k1 = math.sqrt(2) * w
    k2 = w * w
    a0 = k2 / (1 + k1 + k2)
    a1, a2 = 2 * a0, a0
    k3 = 2 * a0 / k2
    b1 = -2 * a0 + k3
    b2 = 1 - 2 * a0 - k3

    def one_pass(y):
        out = np.zeros_like(y)
        out[0], out[1] = y[0], y[1]
        for n in range(2, y.shape[0]):
            out[n] = a0 * y[n] + a1 * y[n - 1] + a2 * y[n - 2] + b1 * out[n - 1] + b2 * out[n - 2]
        return out

    pad = min(x.shape[0] - 1, int(3 * fs / fc))
    head = 2 * x[0] - x[pad:0:-1]
    tail = 2 * x[-1] - x[-2:-pad - 2:-1]
    y = np.concatenate([head, x, tail], axis=0)
    y = one_pass(y)
    y = one_pass(y[::-1])[::-1]
    return y[pad:pad + x.shape[0]]


# --------------------------------------------------------------------------
# numpy forward kinematics over the compiled tables
# --------------------------------------------------------------------------
def _func(t, f: int, x: float):
    kind = t.func_kind[f]
    if kind == ct.MACROS["BIO_FUNC_CONST"]:
        return t.func_c[f][0]
    if kind == ct.MACROS["BIO_FUNC_LINEAR"]:
        return t.func_c[f][0] * x + t.func_c[f][1]
    kb, n = t.func_knot_begin[f], t.func_knot_count[f]
    kx = [t.knot_x[kb + i] for i in range(n)]
    if x <= kx[0]:
        return t.knot_c[kb][0] + t.knot_c[kb][1] * (x - kx[0])
    if x >= kx[-1]:
        return t.knot_c[kb + n - 1][0] + t.knot_c[kb + n - 1][1] * (x - kx[-1])
    i = 0
    while i + 1 < n - 1 and x >= kx[i + 1]:
        i += 1
    dx = x - kx[i]
    c = t.knot_c[kb + i]
    return c[0] + dx * (c[1] + dx * (c[2] + dx * c[3]))


def host_fk(cm: CompiledModel, q: Sequence[float]):
    """Poses (R[nb,3,3], p[nb,3]) of the merged bodies in ground for the free
    coordinates q (dof order)."""
    t = cm.tables
    nb = t.n_bodies
    R = np.zeros((nb, 3, 3))
    p = np.zeros((nb, 3))
    for b in range(nb):
        par = t.body_parent[b]
        Rp = R[par] if par >= 0 else np.eye(3)
        pp = p[par] if par >= 0 else np.zeros(3)
        pos = pp + Rp @ np.asarray(t.body_joint_loc[b][:])
        Rb = Rp.copy()
        for a in range(t.body_axis_begin[b], t.body_axis_begin[b] + t.body_axis_count[b]):
            d = t.axis_dof[a]
            s = _func(t, t.axis_func[a], q[d] if d >= 0 else 0.0)
            ax = np.asarray(t.axis_vec[a][:])
            if t.axis_kind[a] == ct.MACROS["BIO_AXIS_TRANS"]:
                pos = pos + Rp @ ax * s
            else:
                Rb = Rb @ rot_axis(ax, s)
        R[b], p[b] = Rb, pos
    return R, p


def body_kinematics(cm: CompiledModel, q: Sequence[float], bodies: Sequence[str]):
    """What OpenSim's BodyKinematics analysis reports for position: the
    mass-centre of each named (original) body in ground, and the whole-body
    centre of mass."""
    R, p = host_fk(cm, q)
    out = np.zeros((len(bodies), 3))
    for i, b in enumerate(bodies):
        ob = cm.orig_body[b]
        out[i] = p[ob["merged"]] + R[ob["merged"]] @ np.asarray(ob["com"])
    t = cm.tables
    msum = np.zeros(3)
    for b in range(t.n_bodies):
        msum += t.body_mass[b] * (p[b] + R[b] @ np.asarray(t.body_com[b][:]))
    return out, msum / t.total_mass


def sphere_bottoms(cm: CompiledModel, q: Sequence[float]) -> np.ndarray:
    R, p = host_fk(cm, q)
    t = cm.tables
    return np.array([(p[t.sph_body[s]] + R[t.sph_body[s]] @ np.asarray(t.sph_loc[s][:]))[1]
                     - t.sph_radius[s] for s in range(t.n_spheres)])


REF_BODIES = ("pelvis", "femur_r", "tibia_r", "talus_r", "calcn_r", "toes_r",
              "femur_l", "tibia_l", "talus_l", "calcn_l", "toes_l", "torso")


def build_reference(cm: CompiledModel, time: np.ndarray, q_all: np.ndarray,
                    coord_labels: Sequence[str], dt: float = 0.01,
                    lowpass_hz: float = 6.0) -> Dict[str, np.ndarray]:
    """q_all[T, len(coord_labels)] in radians / metres -> reference tables in
    the model's CoordinateSet order (locked coordinates keep the file's
    values, as the reference's q_d frames do)."""
    names = cm.coord_names
    col = {n: i for i, n in enumerate(coord_labels)}
    q = np.zeros((q_all.shape[0], len(names)))
    for j, n in enumerate(names):
        if n in col:
            q[:, j] = q_all[:, col[n]]
        else:
            q[:, j] = cm.tables.coord_const[j]
    tt, q = resample_linear(np.asarray(time, dtype=np.float64), q, dt)
    q = lowpass_zero_phase(q, 1.0 / dt, lowpass_hz)
    u = np.gradient(q, dt, axis=0, edge_order=2)
    bodies = [b for b in REF_BODIES if b in cm.orig_body]
    T = q.shape[0]
    body_pos = np.zeros((T, len(bodies), 3))
    com = np.zeros((T, 3))
    dof_cols = [names.index(n) for n in cm.dof_names]
    for i in range(T):
        body_pos[i], com[i] = body_kinematics(cm, q[i, dof_cols], bodies)
    return dict(time=tt - tt[0], q=q, u=u, body_pos=body_pos, com_pos=com,
                coord_names=list(names), body_names=bodies)


def synth_gait_2d(cm: CompiledModel, curves_deg: Dict[str, Sequence[float]],
                  cycle_steps: int, n_rows: int, speed: float, dt: float = 0.01,
                  penetration: float = 0.008) -> Dict[str, np.ndarray]:
    """Synthetic planar gait: the 2D reference input
    ``healthy_gait.sto`` (reference ``data/2D/walking_reference_data/
    setup_ka.xml:70``) is not shipped, so joint angles follow the healthy
    mean curves embedded in the reference's plotting helper
    (``visualization_utils2D.py:609-745``, 101 samples per cycle, degrees,
    knee flexion positive there / negative in the model), the left leg is the
    right leg half a cycle later, pelvis_tx advances at ``speed`` and
    pelvis_ty is solved per frame so that the lowest contact sphere sinks
    ``penetration`` into the ground."""
    names = cm.coord_names
    phase = (np.arange(n_rows) / float(cycle_steps)) % 1.0

    def curve(name, ph):
        y = np.asarray(curves_deg[name], dtype=np.float64)
        y = np.concatenate([y[:-1], y[:1]])  # periodic closure
        x = np.linspace(0.0, 1.0, y.size)
        return np.deg2rad(np.interp(ph % 1.0, x, y))

    q = np.zeros((n_rows, len(names)))
    for j, n in enumerate(names):
        if n == "pelvis_tilt":
            q[:, j] = curve("pelvis_tilt", phase)
        elif n == "pelvis_tx":
            q[:, j] = speed * dt * np.arange(n_rows)
        elif n.startswith("hip_flexion"):
            q[:, j] = curve("hip_flexion", phase + (0.5 if n.endswith("_l") else 0.0))
        elif n.startswith("knee_angle"):
            q[:, j] = -curve("knee_angle", phase + (0.5 if n.endswith("_l") else 0.0))
        elif n.startswith("ankle_angle"):
            q[:, j] = curve("ankle_angle", phase + (0.5 if n.endswith("_l") else 0.0))
    q = lowpass_zero_phase(q, 1.0 / dt, 6.0)
    jy = names.index("pelvis_ty")
    dof_cols = [names.index(n) for n in cm.dof_names]
    for i in range(n_rows):
        q[i, jy] = 1.0
        low = sphere_bottoms(cm, q[i, dof_cols]).min()
        q[i, jy] = 1.0 - low - penetration
    q[:, jy] = lowpass_zero_phase(q[:, jy:jy + 1], 1.0 / dt, 6.0)[:, 0]
    time = dt * np.arange(n_rows)
    return build_reference(cm, time, q, names, dt=dt, lowpass_hz=-1.0)


def _locked_columns(cm: CompiledModel, q: np.ndarray) -> None:
    """Locked coordinates keep their compiled value in the tables (as the reference's q_d frames do)."""
    for j, n in enumerate(cm.coord_names):
        if n not in cm.dof_names:
            q[:, j] = cm.tables.coord_const[j]


def _solve_pelvis_height(cm: CompiledModel, q: np.ndarray, rows, penetration: float) -> None:
    """pelvis_ty per row such that the lowest contact sphere sinks `penetration` into the ground."""
    names = cm.coord_names
    jy = names.index("pelvis_ty")
    dof_cols = [names.index(n) for n in cm.dof_names]
    for i in rows:
        q[i, jy] = 1.0
        q[i, jy] = 1.0 - sphere_bottoms(cm, q[i, dof_cols]).min() - penetration


def synth_gait(cm: CompiledModel, curves_deg: Dict[str, Sequence[float]], cycle_steps: int, n_rows: int,
               speed: float, dt: float = 0.01, penetration: float = 0.008) -> Dict[str, np.ndarray]:
    """`synth_gait_2d` for any of the gait models, by coordinate name: the sagittal coordinates
    follow the healthy mean curves, every other free coordinate (pelvis list / rotation / tz, hip
    adduction of the 3D models) stays at zero, locked coordinates at their compiled value.  Used for
    the 3D running task, whose reference directory (`running_reference_data`, reference
    ``muscle_running_imitation_env3D.py:47-53``, cycle 70 ``:177``) is not shipped."""
    names = cm.coord_names
    phase = (np.arange(n_rows) / float(cycle_steps)) % 1.0

    def curve(name, ph):
        y = np.asarray(curves_deg[name], dtype=np.float64)
        y = np.concatenate([y[:-1], y[:1]])
        return np.deg2rad(np.interp(ph % 1.0, np.linspace(0.0, 1.0, y.size), y))

    q = np.zeros((n_rows, len(names)))
    for j, n in enumerate(names):
        left = 0.5 if n.endswith("_l") else 0.0
        if n == "pelvis_tilt":
            q[:, j] = curve("pelvis_tilt", phase)
        elif n == "pelvis_tx":
            q[:, j] = speed * dt * np.arange(n_rows)
        elif n.startswith("hip_flexion"):
            q[:, j] = curve("hip_flexion", phase + left)
        elif n.startswith("knee_angle"):
            q[:, j] = -curve("knee_angle", phase + left)
        elif n.startswith("ankle_angle"):
            q[:, j] = curve("ankle_angle", phase + left)
    q = lowpass_zero_phase(q, 1.0 / dt, 6.0)
    _locked_columns(cm, q)
    _solve_pelvis_height(cm, q, range(n_rows), penetration)
    jy = names.index("pelvis_ty")
    q[:, jy] = lowpass_zero_phase(q[:, jy:jy + 1], 1.0 / dt, 6.0)[:, 0]
    return build_reference(cm, dt * np.arange(n_rows), q, names, dt=dt, lowpass_hz=-1.0)


def synth_jump(cm: CompiledModel, n_rows: int = 102, dt: float = 0.01, takeoff_speed: float = 2.2,
               penetration: float = 0.008) -> Dict[str, np.ndarray]:
    """Synthetic counter-movement jump up to the apex.  The jumping envs read
    `highjump_reference_data` / `jumping_reference_data` (reference
    ``muscle_jumping_imitation_env2D.py:47-53``, ``muscle_jumping_imitation_env3D.py:47-53``), which are
    not shipped; they run the table forwards and then mirrored (``:290-294``: index = 2 cycle - istep,
    cycle = N / 2 = rows - 2), so the table holds the way up only: quiet stance, crouch (hip 75 deg,
    knee -95 deg, ankle 25 deg dorsiflexion, trunk 25 deg forward), extension to take-off with the
    ankle plantarflexed, ballistic flight of the pelvis to the apex.  Both legs move together; the feet
    stay on the ground (lowest contact sphere `penetration` deep) until take-off, pelvis_ty follows
    y0 + v t - g t^2 / 2 afterwards with v = `takeoff_speed`."""
    names = cm.coord_names
    g = abs(float(cm.tables.gravity[1]))
    t_fly = takeoff_speed / g                                  # take-off -> apex
    n_fly = int(round(t_fly / dt))
    n_ground = n_rows - n_fly
    if n_ground < 40:
        raise ValueError("synth_jump: too few rows for stance + crouch + extension")
    n_stand = n_ground // 5
    n_ext = n_ground // 5
    n_crouch = n_ground - n_stand - n_ext

    def smooth(a, b, n):                                       # C2 blend a -> b over n samples
        x = np.linspace(0.0, 1.0, n, endpoint=False)
        return a + (b - a) * (10 * x ** 3 - 15 * x ** 4 + 6 * x ** 5)

    def profile(stand, crouch, takeoff, flight):
        return np.concatenate([np.full(n_stand, stand), smooth(stand, crouch, n_crouch),
                               smooth(crouch, takeoff, n_ext), smooth(takeoff, flight, n_fly)])

    deg = np.pi / 180.0
    q = np.zeros((n_rows, len(names)))
    for j, n in enumerate(names):
        if n == "pelvis_tilt":
            q[:, j] = profile(0.0, -25 * deg, 0.0, 0.0)
        elif n.startswith("hip_flexion"):
            q[:, j] = profile(0.0, 75 * deg, -5 * deg, 10 * deg)
        elif n.startswith("knee_angle"):
            q[:, j] = profile(-2 * deg, -95 * deg, -3 * deg, -20 * deg)
        elif n.startswith("ankle_angle"):
            q[:, j] = profile(0.0, 25 * deg, -30 * deg, -15 * deg)
    q = lowpass_zero_phase(q, 1.0 / dt, 6.0)
    _locked_columns(cm, q)
    _solve_pelvis_height(cm, q, range(n_ground + 1), penetration)
    jy = names.index("pelvis_ty")
    tf = dt * np.arange(1, n_fly + 1)
    # flight: continue from the last grounded frame at the take-off speed
    q[n_ground:, jy] = q[n_ground - 1, jy] + takeoff_speed * tf[:n_rows - n_ground] - 0.5 * g * tf[:n_rows - n_ground] ** 2
    q[:, jy] = lowpass_zero_phase(q[:, jy:jy + 1], 1.0 / dt, 6.0)[:, 0]
    return build_reference(cm, dt * np.arange(n_rows), q, names, dt=dt, lowpass_hz=-1.0)


def load_ik_motion(path: str):
    """IK .mot -> (time, q[T, ncol] in rad/m, labels without 'time')."""
    labels, data, in_deg = read_storage(path)
    time = data[:, 0]
    q = data[:, 1:].copy()
    names = labels[1:]
    if in_deg:
        for j, n in enumerate(names):
            if n not in ("pelvis_tx", "pelvis_ty", "pelvis_tz"):
                q[:, j] = np.deg2rad(q[:, j])
    return time, q, names
