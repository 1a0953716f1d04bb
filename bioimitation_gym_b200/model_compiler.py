"""Host-side model compiler: RawModel (osim_parser) -> flat SoA tables.

Replaces, without OpenSim, what the reference obtains from
``opensim.Model(model_predictive.osim)`` + ``initSystem()``
(reference ``opensim_wrapper.py:9-16``) after its in-file model surgery:

* ``construct_predictive_model``      reference ``opensim_utils.py:204-222``
  (contact half-space + 6 spheres + 2 HuntCrossleyForce ``:14-81,144-182``,
  6 CoordinateLimitForce ``:84-141,185-201``, pelvis_ty default 1.02)
* ``convert_model_to_torque_actuated`` reference ``opensim_utils.py:238-270``
* ``convert_model_to_prosthetic``     reference
  ``muscle_locked_knee_imitation_env3D.py:104-125``

The surgery is applied in memory (the reference rewrites a shared file on
disk).  ``compile_model`` then merges welded bodies and
joints whose coordinates are all locked, eliminates locked coordinates, and
fills a ``BioModelTables`` (include/bio_b200.h).
"""
from __future__ import annotations

import copy
import math
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import ctables as ct
from . import curves as _curves

# reference opensim_utils.py:14-51
_SPHERES = [("heel", "calcn", (0.03, 0.02, 0.0), 0.05),
            ("toe1", "toes", (0.02, -0.005, -0.026), 0.025),
            ("toe2", "toes", (0.02, -0.005, 0.026), 0.025)]
# reference opensim_utils.py:54-81
_CONTACT = dict(stiffness=2000000.0, dissipation=1.0, static_friction=0.8,
                dynamic_friction=0.8, viscous_friction=0.6,
                transition_velocity=0.1)
# reference opensim_utils.py:84-141 (degrees)
_LIMITS = [("hip_flexion", 120.0, -30.0), ("knee_angle", 0.0, -140.0),
           ("ankle_angle", 20.0, -40.0)]

# bodies whose origins the envs read (order = body_pos order in the
# observation, reference muscle_walking_imitation_env2D.py:189-200)
OBS_BODIES = ("torso", "calcn_r", "calcn_l", "femur_r", "femur_l",
              "tibia_r", "tibia_l", "talus_r", "talus_l")

# reference muscle_walking_imitation_env2D.py:369 and env3D.py:375-377
SLOW_TWITCH_2D = [0.499, 0.55, 0.5, 0.484, 0.546, 0.759, 0.721] * 2
SLOW_TWITCH_3D = [0.499, 0.55, 0.5, 0.484, 0.546, 0.759, 0.721,
                  0.484, 0.546, 0.759, 0.721, 0.499, 0.55, 0.5,
                  0.484, 0.546, 0.759, 0.721, 0.484, 0.546, 0.759, 0.721]


# --------------------------------------------------------------------------
# model surgery (in memory)
# --------------------------------------------------------------------------
def construct_predictive_model(raw: Dict[str, Any]) -> Dict[str, Any]:
    raw = copy.deepcopy(raw)
    raw["contact_half_spaces"] = [dict(name="platform", body="ground",
                                       loc=[0.0, 0.0, 0.0],
                                       ori=[0.0, 0.0, -math.pi / 2])]
    raw["contact_spheres"] = []
    raw["contact_forces"] = []
    for side in ("r", "l"):
        for nm, body, loc, rad in _SPHERES:
            raw["contact_spheres"].append(dict(
                name="%s_%s" % (nm, side), body="%s_%s" % (body, side),
                loc=list(loc), ori=[0.0, 0.0, 0.0], radius=rad))
    # dict order in the reference: foot_r spheres, then foot_l spheres
    raw["contact_spheres"].sort(key=lambda s: 0 if s["name"].endswith("_r") else 1)
    for side in ("r", "l"):
        raw["contact_forces"].append(dict(
            name="foot_" + side,
            geometries=["platform"] + ["%s_%s" % (n, side) for n, _, _, _ in _SPHERES],
            **_CONTACT))
    raw["limit_forces"] = []
    for nm, up, lo in _LIMITS:
        for side in ("r", "l"):
            raw["limit_forces"].append(dict(
                name="%s_limit_%s" % (nm.split("_angle")[0], side),
                coordinate="%s_%s" % (nm, side), upper_stiffness=20.0,
                upper_limit=up, lower_stiffness=20.0, lower_limit=lo,
                damping=0.25, transition=10.0))
    for c in raw["coordinates"]:
        if c["name"] == "pelvis_ty":
            c["default"] = 1.02
    raw["name"] = "model_predictive"
    return raw


def convert_model_to_torque_actuated(raw, max_actuation: float,
                                     remove_floating_base: bool = True):
    raw = copy.deepcopy(raw)
    raw["muscles"] = []
    raw["actuators"] = []
    for c in raw["coordinates"]:
        if (remove_floating_base and c["name"] in
                ("pelvis_tx", "pelvis_ty", "pelvis_tz")) or c["locked"]:
            continue
        raw["actuators"].append(dict(name=c["name"] + "_actuator",
                                     coordinate=c["name"], optimal_force=1.0,
                                     min_control=-float(max_actuation),
                                     max_control=float(max_actuation)))
    raw["name"] = "model_predictive_no_muscles"
    return raw


def convert_model_to_prosthetic(raw):
    raw = copy.deepcopy(raw)
    raw["muscles"] = [m for m in raw["muscles"]
                      if m["name"] not in ("gastroc_l", "soleus_l", "tib_ant_l")]
    for c in raw["coordinates"]:
        if c["name"] in ("knee_angle_l", "ankle_angle_l"):
            c["default"] = 0.0
            c["locked"] = True
    # a torque model keeps no actuator on a locked coordinate only if the
    # conversion ran after locking; the reference converts to torque first
    # (torque_locked_knee_imitation_env2D.py:56-62), so actuators stay.
    raw["name"] = "model_predictive_prosthetic"
    return raw


# --------------------------------------------------------------------------
# small math helpers
# --------------------------------------------------------------------------
def rot_axis(axis: Sequence[float], ang: float) -> np.ndarray:
    a = np.asarray(axis, dtype=np.float64)
    n = np.linalg.norm(a)
    if n == 0:
        return np.eye(3)
    a = a / n
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)


def euler_xyz_body(o: Sequence[float]) -> np.ndarray:
    """OpenSim frame orientation: body-fixed X-Y-Z rotation sequence."""
    return rot_axis([1, 0, 0], o[0]) @ rot_axis([0, 1, 0], o[1]) @ rot_axis([0, 0, 1], o[2])


def simm_spline_coefficients(x: Sequence[float], y: Sequence[float]):
    """Cubic spline coefficients (b, c, d) of OpenSim's ``SimmSpline``.

    Restated from the published SIMM / Forsythe-Malcolm-Moler ``spline``
    routine (third derivatives at both ends from divided differences; two
    knots give a straight line).  y(x) = y_i + dx*(b_i + dx*(c_i + dx*d_i)),
    dx = x - x_i, linear extrapolation with the end slopes outside.
    PARITY UNPINNED against OpenSim (the source is not in the reference).
    """
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = x.size
    b = np.zeros(n)
    c = np.zeros(n)
    d = np.zeros(n)
    tiny = 1e-30
    if n < 2:
        return b, c, d
    if n == 2:
        t = max(tiny, x[1] - x[0])
        b[:] = (y[1] - y[0]) / t
        return b, c, d
    nm1, nm2 = n - 1, n - 2
    d[0] = max(tiny, x[1] - x[0])
    c[1] = (y[1] - y[0]) / d[0]
    for i in range(1, nm1):
        d[i] = max(tiny, x[i + 1] - x[i])
        b[i] = 2.0 * (d[i - 1] + d[i])
        c[i + 1] = (y[i + 1] - y[i]) / d[i]
        c[i] = c[i + 1] - c[i]
    b[0] = -d[0]
    b[nm1] = -d[nm2]
    c[0] = 0.0
    c[nm1] = 0.0
    if n > 3:
        d31 = max(tiny, x[3] - x[1])
        d20 = max(tiny, x[2] - x[0])
        d1 = max(tiny, x[nm1] - x[n - 3])
        d2 = max(tiny, x[nm2] - x[n - 4])
        d30 = max(tiny, x[3] - x[0])
        d3 = max(tiny, x[nm1] - x[n - 4])
        c[0] = c[2] / d31 - c[1] / d20
        c[nm1] = c[nm2] / d1 - c[n - 3] / d2
        c[0] = c[0] * d[0] * d[0] / d30
        c[nm1] = -c[nm1] * d[nm2] * d[nm2] / d3
    for i in range(1, n):
        t = d[i - 1] / b[i - 1]
        b[i] -= t * d[i - 1]
        c[i] -= t * c[i - 1]
    c[nm1] /= b[nm1]
    for j in range(nm1):
        i = nm2 - j
        c[i] = (c[i] - d[i] * c[i + 1]) / b[i]
    b[nm1] = (y[nm1] - y[nm2]) / d[nm2] + d[nm2] * (c[nm2] + 2.0 * c[nm1])
    for i in range(nm1):
        b[i] = (y[i + 1] - y[i]) / d[i] - d[i] * (c[i + 1] + 2.0 * c[i])
        d[i] = (c[i + 1] - c[i]) / d[i]
        c[i] *= 3.0
    c[nm1] *= 3.0
    d[nm1] = d[nm2]
    return b, c, d


def eval_func(f, q: float) -> float:
    if f[0] == "const":
        return f[1]
    if f[0] == "linear":
        return f[1] * q + f[2]
    xs, ys = np.asarray(f[1]), np.asarray(f[2])
    b, c, d = simm_spline_coefficients(xs, ys)
    if q <= xs[0]:
        return ys[0] + b[0] * (q - xs[0])
    if q >= xs[-1]:
        return ys[-1] + b[-1] * (q - xs[-1])
    i = int(np.searchsorted(xs, q, side="right") - 1)
    dx = q - xs[i]
    return ys[i] + dx * (b[i] + dx * (c[i] + dx * d[i]))


# --------------------------------------------------------------------------
class CompiledModel:
    """BioModelTables plus the python-side metadata (names, reference-body
    COM points) that the env layer and the reference-motion generator use."""

    def __init__(self):
        self.tables = ct.BioModelTables()
        self.name = ""
        self.body_names: List[str] = []        # merged bodies
        self.dof_names: List[str] = []
        self.coord_names: List[str] = []       # all coordinates incl. locked
        self.muscle_names: List[str] = []
        self.actuator_names: List[str] = []
        self.limit_names: List[str] = []
        self.obs_body_names: List[str] = []
        self.orig_body: Dict[str, Any] = {}    # name -> (merged, R_rel, p_rel, mass, com)

    # convenience
    @property
    def n_dof(self):
        return self.tables.n_dof

    def arr(self, name):
        return ct.field_array(self.tables, name)


def _inertia_matrix(i6):
    xx, yy, zz, xy, xz, yz = i6
    return np.array([[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]], dtype=np.float64)


def compile_model(raw: Dict[str, Any], curve_n: Optional[int] = None) -> CompiledModel:
    cm = CompiledModel()
    cm.name = raw.get("name", "")
    t = cm.tables
    t.abi_version = ct.MACROS["BIO_ABI_VERSION"]
    coords = {c["name"]: c for c in raw["coordinates"]}
    coord_order = [c["name"] for c in raw["coordinates"]]
    free = [n for n in coord_order if not coords[n]["locked"]]
    dof_index = {n: i for i, n in enumerate(free)}
    cm.dof_names = list(free)
    cm.coord_names = list(coord_order)

    joint_of_child = {j["child"]: j for j in raw["joints"]}
    # topological order of original bodies
    order: List[str] = []
    placed = {"ground"}
    pending = list(raw["body_order"])
    while pending:
        progressed = False
        for b in list(pending):
            j = joint_of_child.get(b)
            if j is None:
                raise ValueError("body %s has no joint" % b)
            if j["parent"] in placed:
                order.append(b)
                placed.add(b)
                pending.remove(b)
                progressed = True
        if not progressed:
            raise ValueError("kinematic loop or missing parent: %s" % pending)

    def axis_is_free(ax):
        return ax["coord"] is not None and ax["func"][0] != "const" \
            and not coords[ax["coord"]]["locked"]

    def axis_value_const(ax):
        if ax["coord"] is None or ax["func"][0] == "const":
            return eval_func(ax["func"], 0.0)
        return eval_func(ax["func"], coords[ax["coord"]]["default"])

    merged_of: Dict[str, int] = {"ground": -1}
    rel: Dict[str, Any] = {"ground": (np.eye(3), np.zeros(3))}
    merged_names: List[str] = []
    merged_members: List[List[str]] = []
    body_parent: List[int] = []
    joint_loc: List[np.ndarray] = []
    body_axes: List[List[Dict[str, Any]]] = []

    for b in order:
        j = joint_of_child[b]
        axes = j.get("transform", [])
        movable = any(axis_is_free(a) for a in axes)
        Rp_rel, pp_rel = rel[j["parent"]]
        Rjp = euler_xyz_body(j["ori_parent"])
        Rjc = euler_xyz_body(j["ori_child"])
        if movable:
            if not np.allclose(Rp_rel @ Rjp, np.eye(3)) or not np.allclose(Rjc, np.eye(3)) \
                    or not np.allclose(j["loc_child"], 0):
                raise NotImplementedError(
                    "movable joint %s with rotated/offset frames" % j["name"])
            loc = pp_rel + Rp_rel @ np.asarray(j["loc_parent"], dtype=np.float64)
            ax_out = []
            trans = [a for a in axes if a["kind"] == "trans"]
            rots = [a for a in axes if a["kind"] == "rot"]
            for a in trans:
                if axis_is_free(a):
                    ax_out.append(a)
                else:
                    loc = loc + np.asarray(a["axis"]) * axis_value_const(a)
            for a in rots:
                if axis_is_free(a):
                    ax_out.append(a)
                else:
                    v = axis_value_const(a)
                    if v != 0.0:
                        ax_out.append(dict(a, coord=None, func=("const", v)))
            merged_of[b] = len(merged_names)
            rel[b] = (np.eye(3), np.zeros(3))
            merged_names.append(b)
            merged_members.append([b])
            body_parent.append(merged_of[j["parent"]])
            joint_loc.append(loc)
            body_axes.append(ax_out)
        else:
            if merged_of[j["parent"]] < 0:
                raise NotImplementedError("body %s welded to ground" % b)
            Rj = np.eye(3)
            pj = np.zeros(3)
            for a in axes:
                v = axis_value_const(a)
                if a["kind"] == "trans":
                    pj = pj + np.asarray(a["axis"]) * v
                else:
                    Rj = Rj @ rot_axis(a["axis"], v)
            R_rel = Rp_rel @ Rjp @ Rj @ Rjc.T
            p_rel = pp_rel + Rp_rel @ (np.asarray(j["loc_parent"]) + Rjp @ pj) \
                - R_rel @ np.asarray(j["loc_child"])
            merged_of[b] = merged_of[j["parent"]]
            rel[b] = (R_rel, p_rel)
            merged_members[merged_of[b]].append(b)

    nb = len(merged_names)
    if nb > ct.MACROS["BIO_MAX_BODIES"]:
        raise ValueError("too many bodies: %d" % nb)
    cm.body_names = merged_names
    t.n_bodies = nb
    t.n_dof = len(free)
    ct.set_field(t, "gravity", raw["gravity"])

    # mass properties of merged bodies
    masses = np.zeros(nb)
    coms = np.zeros((nb, 3))
    inertias = np.zeros((nb, 6))
    for mb, members in enumerate(merged_members):
        m = sum(raw["bodies"][o]["mass"] for o in members)
        if m <= 0:
            raise ValueError("massless merged body " + merged_names[mb])
        c = np.zeros(3)
        for o in members:
            R_rel, p_rel = rel[o]
            c += raw["bodies"][o]["mass"] * (p_rel + R_rel @ np.asarray(raw["bodies"][o]["com"]))
        c /= m
        I = np.zeros((3, 3))
        for o in members:
            R_rel, p_rel = rel[o]
            mo = raw["bodies"][o]["mass"]
            co = p_rel + R_rel @ np.asarray(raw["bodies"][o]["com"]) - c
            I += R_rel @ _inertia_matrix(raw["bodies"][o]["inertia"]) @ R_rel.T
            I += mo * (np.dot(co, co) * np.eye(3) - np.outer(co, co))
        masses[mb] = m
        coms[mb] = c
        inertias[mb] = [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]]
    for o in order:
        R_rel, p_rel = rel[o]
        cm.orig_body[o] = dict(merged=merged_of[o], R_rel=R_rel, p_rel=p_rel,
                               mass=raw["bodies"][o]["mass"],
                               com=p_rel + R_rel @ np.asarray(raw["bodies"][o]["com"]))
    ct.set_field(t, "body_parent", body_parent)
    ct.set_field(t, "body_mass", masses)
    ct.set_field(t, "body_com", coms)
    ct.set_field(t, "body_inertia", inertias)
    ct.set_field(t, "body_joint_loc", np.asarray(joint_loc))
    t.total_mass = float(sum(raw["bodies"][o]["mass"] for o in order))

    # functions / axes
    func_kind, func_kb, func_kc, func_c = [], [], [], []
    knot_x: List[float] = []
    knot_c: List[List[float]] = []

    def add_func(f) -> int:
        if f[0] == "const":
            func_kind.append(ct.MACROS["BIO_FUNC_CONST"])
            func_kb.append(0)
            func_kc.append(0)
            func_c.append([f[1], 0.0])
        elif f[0] == "linear":
            func_kind.append(ct.MACROS["BIO_FUNC_LINEAR"])
            func_kb.append(0)
            func_kc.append(0)
            func_c.append([f[1], f[2]])
        else:
            xs, ys = np.asarray(f[1], dtype=np.float64), np.asarray(f[2], dtype=np.float64)
            b, c, d = simm_spline_coefficients(xs, ys)
            func_kind.append(ct.MACROS["BIO_FUNC_SPLINE"])
            func_kb.append(len(knot_x))
            func_kc.append(len(xs))
            func_c.append([0.0, 0.0])
            for i in range(len(xs)):
                knot_x.append(float(xs[i]))
                knot_c.append([float(ys[i]), float(b[i]), float(c[i]), float(d[i])])
        return len(func_kind) - 1

    axis_kind, axis_dof, axis_func, axis_vec = [], [], [], []
    ab, ac = [], []
    dof_body = [-1] * len(free)
    for mb in range(nb):
        ab.append(len(axis_kind))
        for a in body_axes[mb]:
            axis_kind.append(ct.MACROS["BIO_AXIS_ROT"] if a["kind"] == "rot"
                             else ct.MACROS["BIO_AXIS_TRANS"])
            if a["coord"] is not None and a["func"][0] != "const":
                di = dof_index[a["coord"]]
                if dof_body[di] not in (-1, mb):
                    raise NotImplementedError("coordinate %s spans joints" % a["coord"])
                dof_body[di] = mb
                axis_dof.append(di)
            else:
                axis_dof.append(-1)
            axis_func.append(add_func(a["func"]))
            v = np.asarray(a["axis"], dtype=np.float64)
            axis_vec.append(v / np.linalg.norm(v))
        ac.append(len(axis_kind) - ab[-1])
    if any(b < 0 for b in dof_body):
        raise ValueError("free coordinate without axis")
    for i in range(1, len(dof_body)):
        if dof_body[i] < dof_body[i - 1]:
            raise NotImplementedError("coordinates not in topological order")
    t.n_axes = len(axis_kind)
    ct.set_field(t, "body_axis_begin", ab)
    ct.set_field(t, "body_axis_count", ac)
    ct.set_field(t, "axis_kind", axis_kind)
    ct.set_field(t, "axis_dof", axis_dof)
    ct.set_field(t, "axis_func", axis_func)
    ct.set_field(t, "axis_vec", np.asarray(axis_vec))
    ct.set_field(t, "dof_body", dof_body)
    ct.set_field(t, "dof_default_q", [coords[n]["default"] for n in free])

    def is_anc(a, b):  # merged body a is ancestor-or-self of b
        while b >= 0:
            if a == b:
                return True
            b = body_parent[b]
        return False
    masks = []
    for i in range(len(free)):
        m = 0
        for jd in range(i + 1):
            if is_anc(dof_body[jd], dof_body[i]):
                m |= 1 << jd
        masks.append(m)
    ct.set_field(t, "dof_anc_mask", masks)

    def to_merged(body, loc):
        R_rel, p_rel = rel[body]
        return merged_of[body], p_rel + R_rel @ np.asarray(loc, dtype=np.float64)

    # muscles
    nm = len(raw["muscles"])
    cm.muscle_names = [m["name"] for m in raw["muscles"]]
    t.n_muscles = nm
    pk, pb, pd, pf, pl, pr = [], [], [], [], [], []
    mpb, mpc = [], []
    max_pen = math.acos(0.1)
    for m in raw["muscles"]:
        mpb.append(len(pk))
        for p in m["points"]:
            kind = p["kind"]
            mbody, loc = to_merged(p["body"], p.get("loc", [0, 0, 0]))
            dof = -1
            funcs = [0, 0, 0]
            rng = [0.0, 0.0]
            if kind == "conditional":
                cn = p["coord"]
                if coords[cn]["locked"]:
                    v = coords[cn]["default"]
                    if not (p["range"][0] <= v <= p["range"][1]):
                        continue  # never active
                    kind = "fixed"
                else:
                    dof = dof_index[cn]
                    rng = list(p["range"])
            elif kind == "moving":
                cns = set(c for c, f in zip(p["coords"], p["funcs"]) if f[0] != "const")
                if len(cns) > 1:
                    raise NotImplementedError("moving point with several coordinates")
                cn = cns.pop() if cns else None
                if cn is None or coords[cn]["locked"]:
                    v = coords[cn]["default"] if cn else 0.0
                    mbody, loc = to_merged(p["body"], [eval_func(f, v) for f in p["funcs"]])
                    kind = "fixed"
                else:
                    R_rel, p_rel = rel[p["body"]]
                    if not (np.allclose(R_rel, np.eye(3)) and np.allclose(p_rel, 0)):
                        raise NotImplementedError("moving point on a merged body")
                    dof = dof_index[cn]
                    funcs = [add_func(f) for f in p["funcs"]]
                    loc = np.zeros(3)
            pk.append({"fixed": 0, "conditional": 1, "moving": 2}[kind])
            pb.append(mbody)
            pd.append(dof)
            pf.append(funcs)
            pl.append(loc)
            pr.append(rng)
        mpc.append(len(pk) - mpb[-1])
    t.n_pathpts = len(pk)
    if nm:
        ct.set_field(t, "mus_pt_begin", mpb)
        ct.set_field(t, "mus_pt_count", mpc)
        ct.set_field(t, "pt_kind", pk)
        ct.set_field(t, "pt_body", pb)
        ct.set_field(t, "pt_dof", pd)
        ct.set_field(t, "pt_func", np.asarray(pf))
        ct.set_field(t, "pt_loc", np.asarray(pl))
        ct.set_field(t, "pt_range", np.asarray(pr))
        g = lambda k: [m[k] for m in raw["muscles"]]
        fiso = np.asarray(g("max_isometric_force"))
        lopt = np.asarray(g("optimal_fiber_length"))
        a0 = np.asarray(g("pennation_angle_at_optimal"))
        height = lopt * np.sin(a0)
        # minimum fibre length: max(active-curve lower bound, pennation limit)
        min_pen = np.where(a0 > 1e-12, height / math.sin(max_pen), lopt * 0.01)
        lm_min = np.maximum(0.4441 * lopt, min_pen)
        ct.set_field(t, "mus_fiso", fiso)
        ct.set_field(t, "mus_lopt", lopt)
        ct.set_field(t, "mus_lts", g("tendon_slack_length"))
        ct.set_field(t, "mus_alpha0", a0)
        ct.set_field(t, "mus_vmax", g("max_contraction_velocity"))
        ct.set_field(t, "mus_tact", g("activation_time_constant"))
        ct.set_field(t, "mus_tdeact", g("deactivation_time_constant"))
        ct.set_field(t, "mus_amin", g("minimum_activation"))
        ct.set_field(t, "mus_beta", g("fiber_damping"))
        ct.set_field(t, "mus_default_act", g("default_activation"))
        ct.set_field(t, "mus_default_lm", g("default_fiber_length"))
        ct.set_field(t, "mus_height", height)
        ct.set_field(t, "mus_lm_min", lm_min)
        ct.set_field(t, "mus_cot_mass", fiso / 0.25e6 * 1059.7 * lopt)
        table = SLOW_TWITCH_2D if nm <= 14 else SLOW_TWITCH_3D
        # the reference indexes its hard-coded list by muscle position
        ct.set_field(t, "mus_slow_twitch", [table[i] for i in range(nm)])

    t.n_funcs = len(func_kind)
    t.n_knots = len(knot_x)
    ct.set_field(t, "func_kind", func_kind)
    ct.set_field(t, "func_knot_begin", func_kb)
    ct.set_field(t, "func_knot_count", func_kc)
    ct.set_field(t, "func_c", np.asarray(func_c))
    if knot_x:
        ct.set_field(t, "knot_x", knot_x)
        ct.set_field(t, "knot_c", np.asarray(knot_c))

    # contact
    sph = raw.get("contact_spheres", [])
    t.n_spheres = len(sph)
    if sph:
        hs = raw["contact_half_spaces"]
        if len(hs) != 1 or hs[0]["body"] != "ground" or \
                not np.allclose(hs[0]["ori"], [0, 0, -math.pi / 2]) or \
                not np.allclose(hs[0]["loc"], 0):
            raise NotImplementedError("only the ground half-space y<0 is supported")
        sb, sl, sg, sr = [], [], [], []
        par = {k: [] for k in ("k", "c", "us", "ud", "uv", "vt")}
        for s in sph:
            grp = None
            for gi, f in enumerate(raw["contact_forces"]):
                if s["name"] in f["geometries"]:
                    grp, cf = gi, f
            if grp is None:
                raise ValueError("sphere %s not used by any force" % s["name"])
            mbody, loc = to_merged(s["body"], s["loc"])
            sb.append(mbody)
            sl.append(loc)
            sg.append(grp)
            sr.append(s["radius"])
            par["k"].append(0.5 * cf["stiffness"] ** (2.0 / 3.0))
            par["c"].append(cf["dissipation"])
            par["us"].append(cf["static_friction"])
            par["ud"].append(cf["dynamic_friction"])
            par["uv"].append(cf["viscous_friction"])
            par["vt"].append(cf["transition_velocity"])
        ct.set_field(t, "sph_body", sb)
        ct.set_field(t, "sph_loc", np.asarray(sl))
        ct.set_field(t, "sph_group", sg)
        ct.set_field(t, "sph_radius", sr)
        for k, v in par.items():
            ct.set_field(t, "sph_" + k, v)

    # coordinate limit forces (degrees -> radians)
    lims = [l for l in raw.get("limit_forces", []) if not coords[l["coordinate"]]["locked"]]
    t.n_limits = len(lims)
    cm.limit_names = [l["name"] for l in lims]
    if lims:
        r2d = 180.0 / math.pi
        ct.set_field(t, "lim_dof", [dof_index[l["coordinate"]] for l in lims])
        ct.set_field(t, "lim_kup", [l["upper_stiffness"] * r2d for l in lims])
        ct.set_field(t, "lim_qup", [l["upper_limit"] / r2d for l in lims])
        ct.set_field(t, "lim_klo", [l["lower_stiffness"] * r2d for l in lims])
        ct.set_field(t, "lim_qlo", [l["lower_limit"] / r2d for l in lims])
        ct.set_field(t, "lim_damp", [l["damping"] * r2d for l in lims])
        ct.set_field(t, "lim_w", [l["transition"] / r2d for l in lims])

    # actuators
    acts = raw.get("actuators", [])
    if acts and nm:
        raise NotImplementedError("mixed muscle/torque actuation")
    if acts:
        t.is_torque = 1
        t.n_act = len(acts)
        cm.actuator_names = [a["name"] for a in acts]
        ad = []
        for a in acts:
            if coords[a["coordinate"]]["locked"]:
                ad.append(-1)  # actuator on a locked coordinate does nothing
            else:
                ad.append(dof_index[a["coordinate"]])
        ct.set_field(t, "act_dof", ad)
        ct.set_field(t, "act_min", [a["min_control"] * a["optimal_force"] for a in acts])
        ct.set_field(t, "act_max", [a["max_control"] * a["optimal_force"] for a in acts])
    else:
        t.is_torque = 0
        t.n_act = nm
        cm.actuator_names = list(cm.muscle_names)
        ct.set_field(t, "act_dof", [-1] * nm)
        ct.set_field(t, "act_min", [m["min_control"] for m in raw["muscles"]])
        ct.set_field(t, "act_max", [m["max_control"] for m in raw["muscles"]])

    # observation points
    obs = [b for b in OBS_BODIES if b in merged_of]
    cm.obs_body_names = obs
    t.n_obspts = len(obs)
    ob, ol = [], []
    for b in obs:
        mbody, loc = to_merged(b, [0, 0, 0])
        ob.append(mbody)
        ol.append(loc)
    ct.set_field(t, "obs_body", ob)
    ct.set_field(t, "obs_loc", np.asarray(ol))

    # all coordinates
    t.n_coords = len(coord_order)
    ct.set_field(t, "coord_dof", [dof_index.get(n, -1) for n in coord_order])
    ct.set_field(t, "coord_const", [coords[n]["default"] if coords[n]["locked"] else 0.0
                                    for n in coord_order])
    ct.set_field(t, "coord_pelvis_trans",
                 [{"pelvis_tx": 1, "pelvis_ty": 2, "pelvis_tz": 3}.get(n, 0)
                  for n in coord_order])

    # muscle curves
    n = ct.MACROS["BIO_CURVE_N"]
    x0s, x1s = [], []
    tab = np.zeros((4, n + 1, 2))
    for ci, cv in enumerate(_curves.default_curves()):
        x0, x1, _, _, tb = _curves.tabulate(cv, n)
        x0s.append(x0)
        x1s.append(x1)
        tab[ci] = tb
    ct.set_field(t, "curve_x0", x0s)
    ct.set_field(t, "curve_x1", x1s)
    ct.set_field(t, "curve_tab", tab)
    return cm
