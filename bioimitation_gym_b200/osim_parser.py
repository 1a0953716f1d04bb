"""OpenSim ``.osim`` reader (no OpenSim dependency).

Parses both XML dialects that the reference's data directory uses
(SURVEY.md section 0.6):

* ``OpenSimDocument Version="30000"`` (2D model): joints are nested inside
  ``<Body>``, offsets are ``<location_in_parent>``/``<location>``, path points
  name their body with ``<body>``.
* ``OpenSimDocument Version="40000"`` (3D / palsy models): ``<JointSet>`` with
  ``<PhysicalOffsetFrame>`` children, path points use ``<socket_parent_frame>``.

The output is a plain-python ``RawModel`` (dicts and lists) that
``model_compiler`` turns into the flat structure-of-arrays tables the CUDA
kernels consume.  It replaces what the reference obtains from
``opensim.Model(model_path)`` (reference ``opensim_wrapper.py:9``).

Functions are represented as tuples:
    ("const", v) | ("linear", slope, intercept) | ("spline", xs, ys)
``MultiplierFunction`` is folded into its inner function at parse time.
"""
from __future__ import annotations

import re
import xml.etree.ElementTree as ET
from typing import Any, Dict, List, Optional

Func = tuple


def _floats(text: Optional[str]) -> List[float]:
    if text is None:
        return []
    return [float(t) for t in text.split()]


def _text(el, tag, default=None):
    c = el.find(tag)
    if c is None or c.text is None:
        return default
    return c.text.strip()


def _bool(el, tag, default=False):
    t = _text(el, tag)
    if t is None:
        return default
    return t.lower() == "true"


def _scale_func(f: Func, s: float) -> Func:
    if f[0] == "const":
        return ("const", f[1] * s)
    if f[0] == "linear":
        return ("linear", f[1] * s, f[2] * s)
    if f[0] == "spline":
        return ("spline", list(f[1]), [y * s for y in f[2]])
    raise ValueError(f[0])


def parse_function(el) -> Func:
    """Parse one OpenSim Function element (Constant / LinearFunction /
    SimmSpline / MultiplierFunction)."""
    tag = el.tag
    if tag == "Constant":
        return ("const", float(_text(el, "value", "0")))
    if tag == "LinearFunction":
        c = _floats(_text(el, "coefficients", "1 0"))
        return ("linear", c[0], c[1])
    if tag in ("SimmSpline", "NaturalCubicSpline"):
        return ("spline", _floats(_text(el, "x")), _floats(_text(el, "y")))
    if tag == "MultiplierFunction":
        inner_holder = el.find("function")
        inner = parse_function(list(inner_holder)[0])
        return _scale_func(inner, float(_text(el, "scale", "1")))
    raise ValueError("unsupported OpenSim function <%s>" % tag)


def _function_child(el) -> Func:
    """A holder element whose (single) child is a Function; in v3 the child
    sits inside <function>, in v4 it is a direct child named 'function'."""
    holder = el.find("function")
    if holder is not None and len(list(holder)) > 0:
        return parse_function(list(holder)[0])
    for c in el:
        if c.tag in ("Constant", "LinearFunction", "SimmSpline",
                     "NaturalCubicSpline", "MultiplierFunction"):
            return parse_function(c)
    raise ValueError("no function under <%s>" % el.tag)


def _parse_coordinate(el) -> Dict[str, Any]:
    rng = _floats(_text(el, "range", "-1e9 1e9"))
    return dict(name=el.get("name"),
                default=float(_text(el, "default_value", "0")),
                default_speed=float(_text(el, "default_speed_value", "0")),
                range=rng,
                locked=_bool(el, "locked"),
                clamped=_bool(el, "clamped"))


def _parse_spatial_transform(el, coord_names) -> List[Dict[str, Any]]:
    out = []
    for ta in el.findall("TransformAxis"):
        name = ta.get("name")
        kind = "rot" if name.startswith("rotation") else "trans"
        coords = (_text(ta, "coordinates", "") or "").split()
        out.append(dict(name=name, kind=kind, axis=_floats(_text(ta, "axis")),
                        coord=coords[0] if coords else None,
                        func=_function_child(ta)))
    # OpenSim order inside SpatialTransform is rotation1..3, translation1..3
    order = {"rotation1": 0, "rotation2": 1, "rotation3": 2,
             "translation1": 3, "translation2": 4, "translation3": 5}
    out.sort(key=lambda t: order[t["name"]])
    return out


def _pin_transform(coord_name: str) -> List[Dict[str, Any]]:
    return [dict(name="rotation1", kind="rot", axis=[0.0, 0.0, 1.0],
                 coord=coord_name, func=("linear", 1.0, 0.0))]


def _strip_path(s: str) -> str:
    return s.strip().split("/")[-1]


def _parse_path_points(gp, v4: bool) -> List[Dict[str, Any]]:
    pts = []
    objs = gp.find("PathPointSet/objects")
    for p in list(objs):
        if p.tag not in ("PathPoint", "ConditionalPathPoint", "MovingPathPoint"):
            raise ValueError("unsupported path point <%s>" % p.tag)
        if v4:
            body = _strip_path(_text(p, "socket_parent_frame"))
        else:
            body = _text(p, "body")
        d = dict(name=p.get("name"), kind="fixed", body=body,
                 loc=_floats(_text(p, "location", "0 0 0")))
        if p.tag == "ConditionalPathPoint":
            d["kind"] = "conditional"
            d["range"] = _floats(_text(p, "range"))
            d["coord"] = _strip_path(_text(p, "socket_coordinate")) if v4 \
                else _text(p, "coordinate")
        elif p.tag == "MovingPathPoint":
            d["kind"] = "moving"
            funcs, coords = [], []
            for ax in "xyz":
                funcs.append(_function_child(p.find(ax + "_location")))
                if v4:
                    coords.append(_strip_path(_text(p, "socket_%s_coordinate" % ax)))
                else:
                    coords.append(_text(p, ax + "_coordinate"))
            d["funcs"] = funcs
            d["coords"] = coords
        pts.append(d)
    return pts


_MUSCLE_DEFAULTS = dict(  # Millard2012EquilibriumMuscle property defaults
    max_contraction_velocity=10.0, activation_time_constant=0.010,
    deactivation_time_constant=0.040, minimum_activation=0.01,
    fiber_damping=0.1, default_activation=0.05, default_fiber_length=0.1,
    min_control=0.0, max_control=1.0, pennation_angle_at_optimal=0.0)


def _parse_muscle(el, v4: bool) -> Dict[str, Any]:
    m = dict(name=el.get("name"), type=el.tag)
    for key in ("max_isometric_force", "optimal_fiber_length",
                "tendon_slack_length"):
        m[key] = float(_text(el, key))
    for key, dv in _MUSCLE_DEFAULTS.items():
        t = _text(el, key)
        m[key] = float(t) if t is not None else dv
    m["points"] = _parse_path_points(el.find("GeometryPath"), v4)
    if el.find("GeometryPath/PathWrapSet/objects") is not None and \
            len(list(el.find("GeometryPath/PathWrapSet/objects"))) > 0:
        raise ValueError("PathWrap objects are not supported (%s)" % m["name"])
    return m


def parse_osim(path: str) -> Dict[str, Any]:
    """Read an .osim file into a RawModel dict."""
    with open(path, "r") as fh:
        txt = fh.read()
    # tags such as <HuntCrossleyForce::ContactParameters> break expat
    txt = re.sub(r"(</?[A-Za-z0-9_]+)::", r"\1__", txt)
    root = ET.fromstring(txt)
    version = int(root.get("Version", "0"))
    v4 = version >= 40000
    model = root.find("Model")
    raw: Dict[str, Any] = dict(name=model.get("name"), version=version,
                               gravity=_floats(_text(model, "gravity", "0 -9.80665 0")))
    bodies: Dict[str, Any] = {}
    joints: List[Dict[str, Any]] = []
    body_order: List[str] = []

    for b in model.find("BodySet/objects").findall("Body"):
        name = b.get("name")
        if name == "ground":
            continue
        if v4:
            inertia = _floats(_text(b, "inertia", "0 0 0 0 0 0"))
        else:
            inertia = [float(_text(b, "inertia_" + k, "0"))
                       for k in ("xx", "yy", "zz", "xy", "xz", "yz")]
        bodies[name] = dict(mass=float(_text(b, "mass", "0")),
                            com=_floats(_text(b, "mass_center", "0 0 0")),
                            inertia=inertia)
        body_order.append(name)
        if not v4:
            jh = b.find("Joint")
            jl = list(jh) if jh is not None else []
            if not jl:
                continue
            j = jl[0]
            coords = [_parse_coordinate(c) for c in
                      (j.find("CoordinateSet/objects").findall("Coordinate")
                       if j.find("CoordinateSet/objects") is not None else [])]
            jd = dict(name=j.get("name"), type=j.tag,
                      parent=_text(j, "parent_body"), child=name,
                      loc_parent=_floats(_text(j, "location_in_parent", "0 0 0")),
                      ori_parent=_floats(_text(j, "orientation_in_parent", "0 0 0")),
                      loc_child=_floats(_text(j, "location", "0 0 0")),
                      ori_child=_floats(_text(j, "orientation", "0 0 0")),
                      coords=coords)
            if _bool(j, "reverse"):
                raise ValueError("reversed joints are not supported")
            if j.tag == "CustomJoint":
                jd["transform"] = _parse_spatial_transform(
                    j.find("SpatialTransform"), [c["name"] for c in coords])
            elif j.tag == "PinJoint":
                jd["transform"] = _pin_transform(coords[0]["name"])
            elif j.tag == "WeldJoint":
                jd["transform"] = []
            else:
                raise ValueError("unsupported joint type " + j.tag)
            joints.append(jd)

    if v4:
        for j in list(model.find("JointSet/objects")):
            frames = {}
            fh = j.find("frames")
            for f in (fh.findall("PhysicalOffsetFrame") if fh is not None else []):
                frames[f.get("name")] = dict(
                    body=_strip_path(_text(f, "socket_parent")),
                    t=_floats(_text(f, "translation", "0 0 0")),
                    o=_floats(_text(f, "orientation", "0 0 0")))
            pf = frames[_text(j, "socket_parent_frame")]
            cf = frames[_text(j, "socket_child_frame")]
            ch = j.find("coordinates")
            coords = [_parse_coordinate(c) for c in
                      (ch.findall("Coordinate") if ch is not None else [])]
            jd = dict(name=j.get("name"), type=j.tag, parent=pf["body"],
                      child=cf["body"], loc_parent=pf["t"], ori_parent=pf["o"],
                      loc_child=cf["t"], ori_child=cf["o"], coords=coords)
            if j.tag == "CustomJoint":
                jd["transform"] = _parse_spatial_transform(
                    j.find("SpatialTransform"), [c["name"] for c in coords])
            elif j.tag == "PinJoint":
                jd["transform"] = _pin_transform(coords[0]["name"])
            elif j.tag == "WeldJoint":
                jd["transform"] = []
            else:
                raise ValueError("unsupported joint type " + j.tag)
            joints.append(jd)

    raw["bodies"] = bodies
    raw["body_order"] = body_order
    raw["joints"] = joints
    raw["coordinates"] = [c for j in joints for c in j["coords"]]

    muscles, contact_spheres, contact_forces, limit_forces, actuators = [], [], [], [], []
    fs = model.find("ForceSet/objects")
    for f in (list(fs) if fs is not None else []):
        if f.tag in ("Millard2012EquilibriumMuscle",):
            muscles.append(_parse_muscle(f, v4))
        elif f.tag == "HuntCrossleyForce":
            cp = f.find(".//HuntCrossleyForce__ContactParameters")
            contact_forces.append(dict(
                name=f.get("name"),
                geometries=(_text(cp, "geometry", "") or "").split(),
                stiffness=float(_text(cp, "stiffness")),
                dissipation=float(_text(cp, "dissipation")),
                static_friction=float(_text(cp, "static_friction")),
                dynamic_friction=float(_text(cp, "dynamic_friction")),
                viscous_friction=float(_text(cp, "viscous_friction")),
                transition_velocity=float(_text(f, "transition_velocity", "0.01"))))
        elif f.tag == "CoordinateLimitForce":
            limit_forces.append(dict(
                name=f.get("name"), coordinate=_text(f, "coordinate"),
                upper_stiffness=float(_text(f, "upper_stiffness")),
                upper_limit=float(_text(f, "upper_limit")),
                lower_stiffness=float(_text(f, "lower_stiffness")),
                lower_limit=float(_text(f, "lower_limit")),
                damping=float(_text(f, "damping")),
                transition=float(_text(f, "transition"))))
        elif f.tag == "CoordinateActuator":
            actuators.append(dict(
                name=f.get("name"), coordinate=_text(f, "coordinate"),
                optimal_force=float(_text(f, "optimal_force", "1")),
                min_control=float(_text(f, "min_control", "-inf")),
                max_control=float(_text(f, "max_control", "inf"))))
        else:
            raise ValueError("unsupported force <%s>" % f.tag)
    cg = model.find("ContactGeometrySet/objects")
    half_spaces = []
    for g in (list(cg) if cg is not None else []):
        body = _strip_path(_text(g, "socket_frame")) if v4 else _text(g, "body_name")
        d = dict(name=g.get("name"), body=body,
                 loc=_floats(_text(g, "location", "0 0 0")),
                 ori=_floats(_text(g, "orientation", "0 0 0")))
        if g.tag == "ContactSphere":
            d["radius"] = float(_text(g, "radius"))
            contact_spheres.append(d)
        elif g.tag == "ContactHalfSpace":
            half_spaces.append(d)
    # MarkerSet (what calc_markers_info reads, reference opensim_wrapper.py:261-282): name, body, location
    markers = []
    ms = model.find("MarkerSet/objects")
    for mk in (ms.findall("Marker") if ms is not None else []):
        body = _strip_path(_text(mk, "socket_parent_frame")) if v4 else _text(mk, "body")
        markers.append(dict(name=mk.get("name"), body=body, loc=_floats(_text(mk, "location", "0 0 0"))))
    raw.update(muscles=muscles, contact_spheres=contact_spheres,
               contact_half_spaces=half_spaces, contact_forces=contact_forces,
               limit_forces=limit_forces, actuators=actuators, markers=markers)
    return raw
