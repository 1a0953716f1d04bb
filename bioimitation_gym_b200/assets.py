"""Compiled model / reference-motion tables shipped with the package.

The reference keeps its models as OpenSim ``.osim`` XML under
``bioimitation/imitation_envs/data`` and rebuilds ``model_predictive.osim``
on first use (reference ``muscle_walking_imitation_env2D.py:55-59``).  Here the
same models are compiled ONCE on the host into flat tables
(``tools/compile_assets.py`` -> ``bioimitation_gym_b200/data/*.npz``); a
user who has the reference data directory can also compile from the XML at
run time with ``model_from_osim``.
"""
from __future__ import annotations

import json
import os
from typing import Dict

import numpy as np

from . import ctables as ct
from . import model_compiler as mc
from .osim_parser import parse_osim

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

# key -> (subject dir below the reference data dir, surgery steps)
MODEL_SPECS = {
    "2d_muscle": ("2D", ()),
    "2d_torque": ("2D", ("torque",)),
    "2d_torque_prosthetic": ("2D", ("torque", "prosthetic")),
    "3d_muscle": ("3D", ()),
    "3d_muscle_prosthetic": ("3D", ("prosthetic",)),
    "3d_torque": ("3D", ("torque",)),
    "3d_torque_prosthetic": ("3D", ("torque", "prosthetic")),
    "palsy_muscle": ("02905/02905_PRE", ()),
}
MODEL_KEYS = tuple(MODEL_SPECS)


def model_from_osim(path: str, surgery=(), max_actuation: float = 200.0) -> mc.CompiledModel:
    """model_scaled.osim -> predictive model (+ torque / prosthetic variants)."""
    raw = parse_osim(path)
    if not raw.get("contact_spheres"):
        raw = mc.construct_predictive_model(raw)
    for s in surgery:
        if s == "torque":
            raw = mc.convert_model_to_torque_actuated(raw, max_actuation)
        elif s == "prosthetic":
            raw = mc.convert_model_to_prosthetic(raw)
        else:
            raise ValueError("unknown surgery step %r" % s)
    return mc.compile_model(raw)


def save_model(cm: mc.CompiledModel, path: str) -> None:
    d = ct.struct_to_dict(cm.tables)
    meta = dict(name=cm.name, body_names=cm.body_names, dof_names=cm.dof_names,
                coord_names=cm.coord_names, muscle_names=cm.muscle_names,
                actuator_names=cm.actuator_names, limit_names=cm.limit_names,
                obs_body_names=cm.obs_body_names,
                orig_body={k: dict(merged=int(v["merged"]), com=[float(x) for x in v["com"]],
                                   p_rel=[float(x) for x in v["p_rel"]],
                                   R_rel=[[float(x) for x in r] for r in v["R_rel"]],
                                   mass=float(v["mass"]))
                           for k, v in cm.orig_body.items()})
    arrays = {"t_" + k: np.asarray(v) for k, v in d.items()}
    np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)


_cache: Dict[str, mc.CompiledModel] = {}


def load_model_file(path: str) -> mc.CompiledModel:
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    d = {k[2:]: z[k] for k in z.files if k.startswith("t_")}
    cm = mc.CompiledModel()
    cm.tables = ct.struct_from_dict(ct.BioModelTables, d)
    if cm.tables.abi_version != ct.MACROS["BIO_ABI_VERSION"]:
        raise RuntimeError("%s was compiled for table ABI %d, package has %d; re-run "
                           "tools/compile_assets.py" % (path, cm.tables.abi_version,
                                                        ct.MACROS["BIO_ABI_VERSION"]))
    for k in ("name", "body_names", "dof_names", "coord_names", "muscle_names",
              "actuator_names", "limit_names", "obs_body_names"):
        setattr(cm, k, meta[k])
    cm.orig_body = {k: dict(merged=v["merged"], com=np.asarray(v["com"]),
                            p_rel=np.asarray(v["p_rel"]), R_rel=np.asarray(v["R_rel"]),
                            mass=v["mass"]) for k, v in meta["orig_body"].items()}
    return cm


def load_model(key: str) -> mc.CompiledModel:
    if key not in _cache:
        _cache[key] = load_model_file(os.path.join(DATA_DIR, "model_%s.npz" % key))
    return _cache[key]


def load_ref(key: str) -> Dict[str, np.ndarray]:
    """Reference-motion tables: q,u [T,n_coords]; body_pos [T,n_refbodies,3];
    com_pos [T,3]; plus the column names."""
    z = np.load(os.path.join(DATA_DIR, "ref_%s.npz" % key), allow_pickle=False)
    out = {k: z[k] for k in z.files if k != "meta"}
    out.update(json.loads(bytes(z["meta"]).decode()))
    return out
