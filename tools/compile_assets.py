#!/usr/bin/env python
"""Compile the reference's model XML and motion inputs into the table
fixtures the package ships (bioimitation_gym_b200/data/*.npz).

Runs in the build container only (reads /root/reference, which does not exist
on the GPU box).  The outputs are derived numeric tables, not reference
source; they are committed together with this script.

  python tools/compile_assets.py [--ref /root/reference]
"""
from __future__ import annotations

import argparse
import ast
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bioimitation_gym_b200 import assets, refmotion  # noqa: E402


def scone_mean_curves(py_path: str):
    """Healthy-gait mean joint-angle curves (degrees, 101 samples per cycle)
    from the reference's plotting helper add_healthy_range_scone
    (visualization_utils2D.py:589-745)."""
    src = open(py_path).read()
    start = src.index("def add_healthy_range_scone")
    body = src[start:]
    out = {}

    def grab(blk, which):
        mm = re.search(r"norm_%s\s*=\s*(?:np\.(multiply|add)\()?\s*(\[.*?\])(?:\s*,\s*(-?[0-9.]+)\s*\))?"
                       % which, blk, re.S)
        if mm is None:
            return None
        vals = np.asarray(ast.literal_eval(mm.group(2)), dtype=np.float64)
        if mm.group(1) == "multiply":
            vals = vals * float(mm.group(3))
        elif mm.group(1) == "add":
            vals = vals + float(mm.group(3))
        return vals

    for var in ("pelvis_tilt", "hip_flexion", "knee_angle", "ankle_angle"):
        m = re.search(r'if "%s" in var_name:(.*?)(?=\n    if |\Z)' % var, body, re.S)
        blk = m.group(1)
        mean = grab(blk, "mean")
        if mean is None:  # the ankle block only carries min/max
            mean = 0.5 * (grab(blk, "min") + grab(blk, "max"))
        out[var] = mean
    return out


def save_ref(key, ref):
    meta = dict(coord_names=ref["coord_names"], body_names=ref["body_names"])
    np.savez_compressed(os.path.join(assets.DATA_DIR, "ref_%s.npz" % key),
                        meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
                        time=ref["time"], q=ref["q"], u=ref["u"], body_pos=ref["body_pos"],
                        com_pos=ref["com_pos"])
    print("ref_%s: rows=%d coords=%d bodies=%d" % (key, ref["q"].shape[0], ref["q"].shape[1],
                                                  ref["body_pos"].shape[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    data = os.path.join(args.ref, "bioimitation", "imitation_envs", "data")
    os.makedirs(assets.DATA_DIR, exist_ok=True)
    models = {}
    for key, (sub, surgery) in assets.MODEL_SPECS.items():
        path = os.path.join(data, sub, "scale", "model_scaled.osim")
        cm = assets.model_from_osim(path, surgery)
        assets.save_model(cm, os.path.join(assets.DATA_DIR, "model_%s.npz" % key))
        models[key] = cm
        t = cm.tables
        print("model_%s: bodies=%d dof=%d muscles=%d act=%d pts=%d mass=%.3f" % (
            key, t.n_bodies, t.n_dof, t.n_muscles, t.n_act, t.n_pathpts, t.total_mass))

    # 3D / palsy walking: IK .mot + setup_ka.xml recipe
    for key, sub, mkey in (("3d_walking", "3D", "3d_muscle"),
                           ("palsy_walking", "02905/02905_PRE", "palsy_muscle")):
        t, q, names = refmotion.load_ik_motion(
            os.path.join(data, sub, "inverse_kinematics", "task_InverseKinematics.mot"))
        save_ref(key, refmotion.build_reference(models[mkey], t, q, names, dt=0.01, lowpass_hz=6.0))

    # 2D: synthetic (inputs not shipped)
    curves = scone_mean_curves(os.path.join(args.ref, "bioimitation", "imitation_envs", "utils",
                                            "visualization_utils2D.py"))
    cm2 = models["2d_muscle"]
    save_ref("2d_walking", refmotion.synth_gait_2d(cm2, curves, cycle_steps=132, n_rows=400, speed=1.0))
    save_ref("2d_running", refmotion.synth_gait_2d(cm2, curves, cycle_steps=70, n_rows=282, speed=1.9))
    # tasks whose reference directories the reference does not ship (highjump / jumping / 3D running)
    save_ref("2d_jumping", refmotion.synth_jump(cm2, n_rows=102))
    cm3 = models["3d_muscle"]
    save_ref("3d_running", refmotion.synth_gait(cm3, curves, cycle_steps=70, n_rows=282, speed=1.6))
    save_ref("3d_jumping", refmotion.synth_jump(cm3, n_rows=102, takeoff_speed=2.0))


if __name__ == "__main__":
    main()
