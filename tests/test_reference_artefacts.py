"""Pins of the oracle against numbers OPENSIM ITSELF produced (CPU; `-m "not gpu"`).

The reference ships no golden dynamics, but its data directory holds OpenSim 4.x outputs for the
3D subjects (tests/golden/make_reference_artefacts.py extracts them into
tests/golden/opensim_artefacts_*.npz): the ScaleTool static pose + placed markers, the
InverseKinematics solution next to the measured markers and force plates, and a
StaticOptimization solution.  They are the first checks of the restated joint / spline / frame
conventions, mass distribution, rigid-body dynamics, moment arms and Millard curves that do not
come from this repository's own formulas (reference call site of the marker read-out:
opensim_wrapper.py:261-282 `calc_markers_info`).

What stays unpinned: Hunt-Crossley contact, CoordinateLimitForce, tendon compliance / damped
equilibrium, activation dynamics, the integrator (no OpenSim output exists for any of them).
"""
import ctypes
import os

import numpy as np
import pytest

from bioimitation_gym_b200 import assets, ctables as ct, refmotion

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SUBJECTS = {"3d": "3d_muscle", "palsy": "palsy_muscle"}


def _load(key):
    return np.load(os.path.join(GOLD, "opensim_artefacts_%s.npz" % key))


def _tables_with_points(cm, bodies, locs):
    """Copy of the model tables whose observation points are the given body-fixed points
    (original, un-merged body names): the oracle's `obs_pos` read-out then is their FK."""
    t = ct.BioModelTables.from_buffer_copy(bytes(cm.tables))
    assert len(bodies) <= ct.MACROS["BIO_MAX_OBSPTS"]
    ob, ol = [], []
    for b, loc in zip(bodies, locs):
        o = cm.orig_body[str(b)]
        ob.append(int(o["merged"]))
        ol.append(np.asarray(o["p_rel"]) + np.asarray(o["R_rel"]) @ np.asarray(loc))
    t.n_obspts = len(ob)
    ct.set_field(t, "obs_body", ob + [0] * (ct.MACROS["BIO_MAX_OBSPTS"] - len(ob)))
    ct.set_field(t, "obs_loc", np.asarray(ol + [[0, 0, 0]] * (ct.MACROS["BIO_MAX_OBSPTS"] - len(ol))))
    return t


def _oracle_points(orc, cm, q_dof, bodies, locs):
    out = []
    K = ct.MACROS["BIO_MAX_OBSPTS"]
    for s in range(0, len(bodies), K):
        t = _tables_with_points(cm, bodies[s:s + K], locs[s:s + K])
        ev = orc.eval_dynamics(t, q_dof, np.zeros(len(q_dof)), act=np.full(t.n_muscles, 0.05),
                               lm=np.asarray([t.mus_lopt[i] for i in range(t.n_muscles)]))
        out.append(ev["obs_pos"])
    return np.concatenate(out, axis=0)


@pytest.mark.parametrize("key", sorted(SUBJECTS))
def test_static_pose_reproduces_opensims_marker_placement(oracle_lib, models, key):
    """ScaleTool's MarkerPlacer put every model marker on the averaged static-trial marker using
    OpenSim's own FK at the static.mot pose (all 14 free + 3 locked coordinates non-zero): the
    oracle's FK of the compiled (weld-merged, lock-eliminated) model must land on the same points.
    Measured: 3D 4e-7 m, palsy 5e-7 m (the .mot prints 8 decimals)."""
    z = _load(key)
    cm = models[SUBJECTS[key]]
    sq = dict(zip(z["static_coord_names"].tolist(), z["static_q"].tolist()))
    for i, n in enumerate(cm.coord_names):               # the locked coordinates were compiled at these values
        if n not in cm.dof_names:
            assert abs(sq[n] - cm.tables.coord_const[i]) < 1e-6, n
    q = np.asarray([sq[n] for n in cm.dof_names])
    names = z["marker_names"].tolist()
    trc = dict(zip(z["static_trc_names"].tolist(), z["static_trc_mean"]))
    use = [i for i, n in enumerate(names) if n in trc]
    assert len(use) >= 24
    x = _oracle_points(oracle_lib, cm, q, z["marker_body"][use].tolist(), z["marker_loc"][use])
    err = np.linalg.norm(x - np.asarray([trc[names[i]] for i in use]), axis=1)
    assert err.max() < 5e-6, dict(zip([names[i] for i in use], err))
    # and the host FK used for the reference-motion tables agrees with the oracle
    R, p = refmotion.host_fk(cm, q)
    for k, i in enumerate(use):
        o = cm.orig_body[str(z["marker_body"][i])]
        xh = p[o["merged"]] + R[o["merged"]] @ (np.asarray(o["p_rel"]) + np.asarray(o["R_rel"]) @ z["marker_loc"][i])
        assert np.abs(xh - x[k]).max() < 1e-9


@pytest.mark.parametrize("key,rms_max,worst_max", [("3d", 0.015, 0.045), ("palsy", 0.07, 0.36)])
def test_ik_solution_tracks_the_measured_markers(oracle_lib, models, key, rms_max, worst_max):
    """FK at every 4th row of OpenSim's IK solution against the measured markers of the same frame:
    within the IK residual (measured: 3D subject RMS 1.19 cm / max 3.9 cm; palsy subject RMS 6.1 cm /
    max 32 cm -- the same model reproduces its static pose to 5e-7 m, so that is OpenSim's own
    tracking error for this subject, whose hip rotations are locked)."""
    z = _load(key)
    cm = models[SUBJECTS[key]]
    cn = z["ik_coord_names"].tolist()
    cols = [cn.index(n) for n in cm.dof_names]
    names = z["marker_names"].tolist()
    tn = z["trc_names"].tolist()
    idx = [names.index(n) for n in tn]
    errs = []
    for i in range(0, z["ik_q"].shape[0], 4):
        x = _oracle_points(oracle_lib, cm, z["ik_q"][i, cols], z["marker_body"][idx].tolist(), z["marker_loc"][idx])
        errs.append(np.linalg.norm(x - z["trc_xyz"][i], axis=1))
    e = np.asarray(errs)
    assert np.sqrt(np.nanmean(e ** 2)) < rms_max
    assert np.nanmax(e) < worst_max


def _filtered_spline(z, cm):
    from scipy.interpolate import make_interp_spline
    cn = z["ik_coord_names"].tolist()
    cols = [cn.index(n) for n in cm.dof_names]
    qf = refmotion.lowpass_zero_phase(z["ik_q"][:, cols], 100.0, 6.0)     # setup_*.xml: 6 Hz low-pass
    return make_interp_spline(z["ik_time"], qf, k=5)


def test_force_plates_obey_newtons_law_on_the_model_com(oracle_lib, models):
    """3D subject, the window OpenSim's own analyses use (static_optimization/setup_so.xml: 1.24-2.0 s,
    every contact on a force plate): m (a_com + g) from the ORACLE's whole-body centre of mass along
    the IK motion against the measured vertical ground reaction.  Measured: correlation 0.95, RMS
    difference 29 N of 384 N RMS, mean 387 N against 378 N (m g = 407 N)."""
    z = _load("3d")
    cm = models["3d_muscle"]
    spl = _filtered_spline(z, cm)
    t = z["ik_time"]
    w = (t >= 1.24) & (t <= 2.0)
    T = cm.tables
    com = []
    for tt in t:
        ev = oracle_lib.eval_dynamics(T, spl(tt), spl(tt, 1), act=np.full(T.n_muscles, 0.05),
                                      lm=np.asarray([T.mus_lopt[i] for i in range(T.n_muscles)]))
        com.append(ev["com_pos"])
    com = np.asarray(com)
    acc = np.gradient(np.gradient(com, 0.01, axis=0), 0.01, axis=0)
    fy = T.total_mass * (acc[:, 1] + 9.80665)
    lab = z["grf_labels"].tolist()
    meas = z["grf"][:len(t), lab.index("left_ground_force_vy")] + z["grf"][:len(t), lab.index("right_ground_force_vy")]
    assert np.corrcoef(fy[w], meas[w])[0, 1] > 0.92
    assert np.sqrt(np.mean((fy[w] - meas[w]) ** 2)) < 0.11 * np.sqrt(np.mean(meas[w] ** 2))
    assert abs(fy[w].mean() - meas[w].mean()) < 0.05 * meas[w].mean()


def test_static_optimization_solution_balances_the_oracles_equations_of_motion(oracle_lib, models):
    """OpenSim's StaticOptimization output for the 3D subject (77 frames, 22 activations + 6 pelvis
    residual actuators + reserves) put into the ORACLE's equations of motion:

        M(q) qdd + bias(q, qd) - J^T GRF  =  sum_m (-dL_m/dq) a_m F_iso f_L f_V cos(alpha)  +  residuals

    with q from the IK solution (6 Hz low-pass, quintic spline derivatives), M / bias / path lengths
    from the oracle and the rigid-tendon active force OpenSim's StaticOptimization uses.
    * pelvis rows hold no muscle: OpenSim's residual actuators FX..MZ are its inverse dynamics, i.e. they
      pin mass matrix, gravity / velocity bias and the application of the plate forces
      (measured RMS mismatch 0.4-2.5 N|N.m against 5.5-29.7 RMS signal: 4-20 %, the size expected of a
      different low-pass / differentiation);
    * hip and knee rows pin moment arms (-dL/dq) and the active force-length / force-velocity curves
      (measured 0.6-3.3 N.m of 9-17 N.m RMS);
    * the ankle rows do NOT close: with OpenSim's activations the plantarflexors deliver 25-30 % less
      moment than the inverse dynamics asks for (8 N.m RMS of 19-29).  There the rigid-tendon fibre
      length sits at 1.4-1.55 l_opt on the descending limb (1 mm of path = 2.4 % l_opt for the soleus), so
      the row is hypersensitive, and the shipped solution may predate the shipped muscle parameters;
      recorded as measured, bounded loosely, cause not identified."""
    z = _load("3d")
    cm, cmt = models["3d_muscle"], models["3d_torque"]
    T, TT = cm.tables, cmt.tables
    nd, nm = T.n_dof, T.n_muscles
    spl = _filtered_spline(z, cm)
    so_n = z["so_names"].tolist()
    grf, gl = z["grf"], z["grf_labels"].tolist()

    def fk_point(q, body, loc):
        R, p = refmotion.host_fk(cm, q)
        o = cm.orig_body[body]
        Rb = R[o["merged"]] @ np.asarray(o["R_rel"])
        return p[o["merged"]] + R[o["merged"]] @ np.asarray(o["p_rel"]) + Rb @ np.asarray(loc), Rb

    def jac(q, body, pw):
        x0, R0 = fk_point(q, body, [0, 0, 0])
        loc = R0.T @ (pw - x0)
        Jv, Jw, e = np.zeros((3, nd)), np.zeros((3, nd)), 1e-6
        for j in range(nd):
            qp, qm = q.copy(), q.copy()
            qp[j] += e
            qm[j] -= e
            xp, Rp = fk_point(qp, body, loc)
            xm, Rm = fk_point(qm, body, loc)
            Jv[:, j] = (xp - xm) / (2 * e)
            dR = (Rp - Rm) / (2 * e) @ R0.T
            Jw[:, j] = [dR[2, 1], dR[0, 2], dR[1, 0]]
        return Jv, Jw

    res, sig = [], []
    for k in range(0, len(z["so_time"]), 2):
        tt = z["so_time"][k]
        q, u, a = spl(tt), spl(tt, 1), spl(tt, 2)
        ql = q.copy()
        ql[cm.dof_names.index("pelvis_ty")] += 10.0          # off the ground: bias without contact forces
        ev = oracle_lib.eval_dynamics(TT, ql, u, ctrl=np.zeros(TT.n_act))
        b = ev["bias"].copy()
        for l in range(TT.n_limits):                         # ... and without the limit forces
            b[TT.lim_dof[l]] += ev["limit_force"][l]
        tau = ev["mass_matrix"] @ a + b
        i = int(round(tt / 0.01))
        for side, body in (("left", "calcn_l"), ("right", "calcn_r")):
            F, P, M = (np.array([grf[i, gl.index("%s_ground_%s%s" % (side, kind, c))] for c in "xyz"])
                       for kind in ("force_v", "force_p", "torque_"))
            if np.abs(F).sum() + np.abs(M).sum() > 0:
                Jv, Jw = jac(q, body, P)
                tau -= Jv.T @ F + Jw.T @ M
        # OpenSim's residual (pelvis point / torque actuators, model/reserve_actuators.xml) and reserve actuators
        Jv, Jw = jac(q, "pelvis", fk_point(q, "pelvis", [-0.0707, 0, 0])[0])
        Q = Jv.T @ np.array([z["so_value"][k, so_n.index(n)] for n in ("FX", "FY", "FZ")]) + \
            Jw.T @ np.array([z["so_value"][k, so_n.index(n)] for n in ("MX", "MY", "MZ")])
        for j, n in enumerate(cm.dof_names):
            if n + "_reserve" in so_n:
                Q[j] += z["so_value"][k, so_n.index(n + "_reserve")]
        L, Ld = oracle_lib.path_lengths(T, q, u)
        for j in range(nd):
            qp, qm = q.copy(), q.copy()
            qp[j] += 1e-6
            qm[j] -= 1e-6
            dL = (oracle_lib.path_lengths(T, qp)[0] - oracle_lib.path_lengths(T, qm)[0]) / 2e-6
            for m in range(nm):
                la = L[m] - T.mus_lts[m]
                lm = np.hypot(la, T.mus_height[m])
                ca = la / lm
                fal = oracle_lib.curve_eval(T, 0, lm / T.mus_lopt[m])[0]
                fv = oracle_lib.curve_eval(T, 1, Ld[m] * ca / (T.mus_vmax[m] * T.mus_lopt[m]))[0]
                Q[j] -= dL[m] * z["so_value"][k, so_n.index(cm.muscle_names[m])] * T.mus_fiso[m] * fal * fv * ca
        res.append(tau - Q)
        sig.append(tau)
    res, sig = np.sqrt(np.mean(np.square(res), axis=0)), np.sqrt(np.mean(np.square(sig), axis=0))
    rel = dict(zip(cm.dof_names, res / sig))
    absr = dict(zip(cm.dof_names, res))
    for n in ("pelvis_tilt", "pelvis_list", "pelvis_rotation", "pelvis_tx", "pelvis_ty", "pelvis_tz"):
        assert rel[n] < 0.30 and absr[n] < 4.0, (n, rel[n], absr[n])
    assert rel["pelvis_ty"] < 0.10                            # 1.9 N of 29.7 N: gravity + vertical inertia
    for n in ("hip_flexion_r", "hip_adduction_r", "knee_angle_r", "hip_flexion_l", "hip_adduction_l", "knee_angle_l"):
        assert rel[n] < 0.35 and absr[n] < 5.0, (n, rel[n], absr[n])
    for n in ("ankle_angle_r", "ankle_angle_l"):              # recorded mismatch, see the docstring
        assert rel[n] < 0.55 and absr[n] < 12.0, (n, rel[n], absr[n])


@pytest.mark.needs_reference
def test_fixtures_are_what_the_reference_files_hold(tmp_path):
    """The committed fixtures equal a fresh extraction from /root/reference (build container only)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mra", os.path.join(GOLD, "make_reference_artefacts.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.HERE = str(tmp_path)
    mod.main()
    for key in SUBJECTS:
        a, b = _load(key), np.load(os.path.join(str(tmp_path), "opensim_artefacts_%s.npz" % key))
        assert sorted(a.files) == sorted(b.files)
        for f in a.files:
            if a[f].dtype.kind in "fc":
                np.testing.assert_allclose(a[f], b[f], rtol=0, atol=0, equal_nan=True)
            else:
                assert (a[f] == b[f]).all()
