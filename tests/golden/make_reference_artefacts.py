#!/usr/bin/env python
"""Extract the OpenSim-PRODUCED numeric artefacts the reference repository ships for the 3D
subjects into small fixtures (tests/golden/opensim_artefacts_*.npz).

These are the only numbers under /root/reference that came out of OpenSim 4.x itself, so they
are the only non-self-referential pins of the oracle's conventions (tests/test_reference_artefacts.py):

  scale/static.mot + experimental_data/static.trc + <MarkerSet> of scale/model_scaled.osim
      ScaleTool's marker placer moved every (non-fixed) model marker onto the averaged
      experimental marker in the static pose USING OPENSIM'S OWN FORWARD KINEMATICS: FK of the
      compiled model at the static.mot coordinates must reproduce the TRC means (joint frames,
      rotation order, knee splines, weld merging).
  inverse_kinematics/task_InverseKinematics.mot + experimental_data/task.trc
      OpenSim's IK solution: the model markers track the measured ones within the IK residual.
  experimental_data/task_grf.mot
      measured ground reaction: Newton's law on the whole-body centre of mass of the compiled
      model along the IK motion (masses, mass centres, FK).
  static_optimization/task_StaticOptimization_controls.xml (3D subject only)
      OpenSim's StaticOptimization solution (22 muscle activations, 6 pelvis residuals, 11
      reserves per frame): the equations of motion assembled from the ORACLE's mass matrix, bias,
      moment arms and Millard force-length-velocity curves must balance with OpenSim's numbers.

Reads /root/reference (build container only); the outputs are derived numeric tables and are
committed with this script.

  python tests/golden/make_reference_artefacts.py
"""
from __future__ import annotations

import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from bioimitation_gym_b200 import osim_parser, refmotion  # noqa: E402

DATA = "/root/reference/bioimitation/imitation_envs/data"
SUBJECTS = (("3d", "3D"), ("palsy", "02905/02905_PRE"))


def read_trc(path):
    lines = open(path).read().splitlines()
    names = [h for h in lines[3].split("\t")[2:] if h.strip()]
    rows = []
    for ln in lines[5:]:
        if not ln.strip():
            continue
        v = [float(x) if x.strip() else np.nan for x in ln.split("\t")][:2 + 3 * len(names)]
        v += [np.nan] * (2 + 3 * len(names) - len(v))
        rows.append(v)
    d = np.asarray(rows)
    return names, d[:, 1], d[:, 2:].reshape(len(rows), len(names), 3) / 1000.0     # mm -> m


def read_controls(path):
    txt = open(path).read()
    out = {}
    for m in re.finditer(r'<ControlLinear name="([^"]+)">(.*?)</ControlLinear>', txt, re.S):
        t = [float(x) for x in re.findall(r"<t>([^<]+)</t>", m.group(2))]
        v = [float(x) for x in re.findall(r"<value>([^<]+)</value>", m.group(2))]
        out[m.group(1)] = (np.asarray(t), np.asarray(v))
    return out


def main():
    for key, sub in SUBJECTS:
        base = os.path.join(DATA, sub)
        raw = osim_parser.parse_osim(os.path.join(base, "scale", "model_scaled.osim"))
        out = {}
        out["marker_names"] = np.asarray([m["name"] for m in raw["markers"]])
        out["marker_body"] = np.asarray([m["body"] for m in raw["markers"]])
        out["marker_loc"] = np.asarray([m["loc"] for m in raw["markers"]], dtype=np.float64)
        # static pose (state-path labels ".../<coordinate>/value")
        labels, data, in_deg = refmotion.read_storage(os.path.join(base, "scale", "static.mot"))
        assert not in_deg
        cn = [l.split("/")[-2] for l in labels if l.endswith("/value")]
        cv = [data[0, i] for i, l in enumerate(labels) if l.endswith("/value")]
        out["static_coord_names"] = np.asarray(cn)
        out["static_q"] = np.asarray(cv)
        names, tt, xyz = read_trc(os.path.join(base, "experimental_data", "static.trc"))
        out["static_trc_names"] = np.asarray(names)
        out["static_trc_mean"] = np.nanmean(xyz, axis=0)
        # IK solution (radians / metres) and the measured markers of the same frames
        t, q, qn = refmotion.load_ik_motion(os.path.join(base, "inverse_kinematics", "task_InverseKinematics.mot"))
        out["ik_time"], out["ik_q"], out["ik_coord_names"] = t, q, np.asarray(qn)
        names, tt, xyz = read_trc(os.path.join(base, "experimental_data", "task.trc"))
        assert len(tt) == len(t) and np.allclose(tt, t, atol=1e-6)
        keep = [i for i, n in enumerate(names) if n in set(out["marker_names"].tolist())]
        out["trc_names"] = np.asarray([names[i] for i in keep])
        out["trc_xyz"] = xyz[:, keep, :].astype(np.float32)
        labels, data, _ = refmotion.read_storage(os.path.join(base, "experimental_data", "task_grf.mot"))
        out["grf_labels"] = np.asarray(labels)
        out["grf"] = data
        so = os.path.join(base, "static_optimization", "task_StaticOptimization_controls.xml")
        if os.path.exists(so):
            ctl = read_controls(so)
            nm = list(ctl)
            out["so_names"] = np.asarray(nm)
            out["so_time"] = ctl[nm[0]][0]
            out["so_value"] = np.stack([ctl[n][1] for n in nm], axis=1)
        path = os.path.join(HERE, "opensim_artefacts_%s.npz" % key)
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), "bytes;", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
