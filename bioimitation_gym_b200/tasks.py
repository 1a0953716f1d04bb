"""Per-env-ID task tables: everything the reference's 17 env classes hard-code.

Each ``EnvSpec`` row restates the constants of one reference class (SURVEY
App. B has the file:line of every number); ``make_task_config`` merges it
with the user config (keys of reference ``configs/env_default.py:7-15``) and
the backend keys (``substeps``, ``integrator`` ...) into a ``BioTaskConfig``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Mapping, Optional, Sequence, Tuple

import numpy as np

from . import ctables as ct
from .model_compiler import CompiledModel

M = ct.MACROS

# reference configs/env_default.py:7-15
DEFAULT_ENV_CONFIG: Dict[str, Any] = dict(
    visualize=False, max_actuation=200, mode="train", log=False,
    r_weights=[0.8, 0.2, 0.1], apply_perturbations=False,
    use_target_obs=True, use_GRF=True, horizon=5)

# backend keys (not in the reference): fixed-step integrator and batching
DEFAULT_BACKEND_CONFIG: Dict[str, Any] = dict(
    num_envs=1, device=0, dtype="float32", substeps=None, integrator=None,
    newton_iters=None, seed=0, env_offset=0, auto_reset=True)

INTEGRATORS = {"semi_implicit_euler": M["BIO_INT_SEMI_IMPLICIT_EULER"],
               "rk2": M["BIO_INT_RK2_MIDPOINT"], "rk4": M["BIO_INT_RK4"],
               "implicit_damping": M["BIO_INT_IMPLICIT_DAMPING"],
               "adaptive_rkm": M["BIO_INT_ADAPTIVE_RKM"]}
# Stated fixed-step scheme replacing OpenSim's adaptive Manager integrator
# (opensim_wrapper.py:287-301): semi-implicit Euler with the dissipative
# contact / limit forces linearly implicit in the speeds, h = 0.01 s / 20 =
# 0.5 ms (DESIGN.md "Integrator" has the stability / accuracy measurements).
DEFAULT_INTEGRATOR = "implicit_damping"
DEFAULT_SUBSTEPS = 20
DEFAULT_NEWTON_ITERS = 30      # cap; the loop stops at |delta| < tolerance (typically 2-5 iterations)


@dataclass
class EnvSpec:
    env_id: str
    cls_name: str
    model: str                 # assets.MODEL_SPECS key
    ref: str                   # reference-motion key
    spatial: bool
    torque: bool
    cycle: Optional[int]       # None: N/2 (jumping)
    n_train: Optional[int]     # None: rows-2 ; "2x": 2*(rows-2)
    n_test_mult: int           # N(test) = mult*(rows-2)
    reset_max: str             # "N/2" | "cycle"
    feed_mean: bool
    term_height: float
    term_limit: float
    term_acc: float
    action_r_scale: float
    perturb_thresh: float
    jumping: bool = False
    perturb_negative_only: bool = False
    broken_as_shipped: str = ""   # reference bug that prevents the class from running


def _spec(*a, **k):
    return EnvSpec(*a, **k)


# SURVEY App. B (each number carries its reference file:line there)
ENV_SPECS: Dict[str, EnvSpec] = {s.env_id: s for s in [
    _spec("MuscleWalkingImitation2D-v0", "MuscleWalkingImitationEnv2D", "2d_muscle", "2d_walking",
          False, False, 132, 264, 1, "N/2", True, 0.75, 1000.0, 1e4, 1.0, 1.8),
    _spec("MuscleRunningImitation2D-v0", "MuscleRunningImitationEnv2D", "2d_muscle", "2d_running",
          False, False, 70, None, 1, "N/2", False, 0.75, 1e4, 1e6, 1.0, 1.8,
          broken_as_shipped="self.w_effort never set (muscle_running_imitation_env2D.py:30-31)"),
    _spec("MuscleJumpingImitation2D-v0", "MuscleJumpingImitationEnv2D", "2d_muscle", "2d_jumping",
          False, False, None, None, 2, "N/2", False, 0.3, 1e4, 1e6, 1.0, 1.8, jumping=True),
    _spec("MuscleLockedKneeImitation2D-v0", "MuscleLockedKneeImitationEnv2D", "2d_muscle", "2d_walking",
          False, False, 132, 264, 1, "N/2", False, 0.75, 1e4, 1e6, 1.0, 1.8),
    _spec("MuscleWalkingImitation3D-v0", "MuscleWalkingImitationEnv3D", "3d_muscle", "3d_walking",
          True, False, 50, None, 1, "cycle", True, 0.75, 1e4, 1e6, 0.5, 1.5),
    _spec("MuscleRunningImitation3D-v0", "MuscleRunningImitationEnv3D", "3d_muscle", "3d_running",
          True, False, 70, None, 1, "N/2", True, 0.75, 1e4, 1e6, 0.5, 1.5),
    _spec("MuscleJumpingImitation3D-v0", "MuscleJumpingImitationEnv3D", "3d_muscle", "3d_jumping",
          True, False, None, None, 2, "cycle", True, 0.3, 1e4, 1e6, 0.5, 1.5, jumping=True,
          broken_as_shipped="self.N used before assignment (muscle_jumping_imitation_env3D.py:73)"),
    _spec("MuscleLockedKneeImitation3D-v0", "MuscleLockedKneeImitationEnv3D", "3d_muscle_prosthetic",
          "3d_walking", True, False, 50, None, 1, "cycle", True, 0.75, 1e4, 1e6, 0.5, 1.5),
    _spec("MusclePalsyImitation3D-v0", "MusclePalsyImitationEnv3D", "palsy_muscle", "palsy_walking",
          True, False, 50, None, 1, "cycle", False, 0.75, 1e4, 1e6, 0.5, 1.5,
          perturb_negative_only=True),
    _spec("TorqueWalkingImitation2D-v0", "TorqueWalkingImitationEnv2D", "2d_torque", "2d_walking",
          False, True, 132, 264, 1, "N/2", True, 0.75, 1000.0, 1e5, 1.0, 1.5),
    _spec("TorqueRunningImitation2D-v0", "TorqueRunningImitationEnv2D", "2d_torque", "2d_running",
          False, True, 70, None, 1, "N/2", True, 0.75, 1e4, 1e6, 1.0, 1.5),
    _spec("TorqueJumpingImitation2D-v0", "TorqueJumpingImitationEnv2D", "2d_torque", "2d_jumping",
          False, True, None, None, 2, "N/2", True, 0.3, 1e4, 1e6, 1.0, 1.5, jumping=True),
    _spec("TorqueLockedKneeImitation2D-v0", "TorqueLockedKneeImitationEnv2D", "2d_torque_prosthetic",
          "2d_walking", False, True, 132, 264, 1, "N/2", True, 0.75, 1e4, 1e6, 1.0, 1.5),
    _spec("TorqueWalkingImitation3D-v0", "TorqueWalkingImitationEnv3D", "3d_torque", "3d_walking",
          True, True, 132, 264, 1, "N/2", True, 0.75, 1e4, 1e6, 1.0, 1.5),
    _spec("TorqueRunningImitation3D-v0", "TorqueRunningImitationEnv3D", "3d_torque", "3d_running",
          True, True, 70, None, 1, "N/2", True, 0.75, 1e4, 1e6, 1.0, 1.5),
    _spec("TorqueJumpingImitation3D-v0", "TorqueJumpingImitationEnv3D", "3d_torque", "3d_walking",
          True, True, None, None, 2, "N/2", True, 0.3, 1e4, 1e6, 1.0, 1.5, jumping=True),
    _spec("TorqueLockedKneeImitation3D-v0", "TorqueLockedKneeImitationEnv3D", "3d_torque_prosthetic",
          "3d_walking", True, True, 132, 264, 1, "N/2", True, 0.75, 1e4, 1e6, 1.0, 1.5),
]}

# torque env2D.py:134-139 / torque env3D.py:131-137 (index lists reproduced
# verbatim, including the 3D duplication of 8 and the locked hip rotations)
_PD_2D = dict(x=[0, 1, 2, 3, 4, 5, 6], v=[0, 3, 4, 5, 6, 7, 8],
              kp=[100, 100, 100, 50, 100, 100, 50], kv=[5, 5, 5, 2, 5, 5, 2])
_PD_3D = dict(x=[0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10], v=[0, 4, 5, 6, 8, 8, 9, 10, 11, 12, 13],
              kp=[100, 100, 100, 100, 100, 50, 100, 100, 100, 100, 50],
              kv=[5, 5, 5, 5, 5, 2, 5, 5, 5, 5, 2])


def merged_config(config: Optional[Mapping[str, Any]]) -> Dict[str, Any]:
    out = dict(DEFAULT_ENV_CONFIG)
    out.update(DEFAULT_BACKEND_CONFIG)
    if config is not None:
        items = config.items() if hasattr(config, "items") else dict(config).items()
        for k, v in items:
            out[k] = v
    return out


def make_task_config(spec: EnvSpec, cm: CompiledModel, ref: Mapping[str, Any],
                     config: Mapping[str, Any]) -> ct.BioTaskConfig:
    cfg = merged_config(config)
    t = cm.tables
    c = ct.BioTaskConfig()
    c.abi_version = M["BIO_ABI_VERSION"]
    c.dt = 0.01                                       # env2D.py:41
    c.n_substeps = int(cfg["substeps"] or DEFAULT_SUBSTEPS)
    integ = cfg["integrator"] or DEFAULT_INTEGRATOR
    c.integrator = INTEGRATORS[integ] if isinstance(integ, str) else int(integ)
    c.newton_iters = int(cfg["newton_iters"] or DEFAULT_NEWTON_ITERS)
    c.horizon = int(cfg["horizon"])
    if not 1 <= c.horizon <= M["BIO_MAX_HORIZON"]:
        raise ValueError("horizon must be in 1..%d" % M["BIO_MAX_HORIZON"])
    c.feed_mean_action = int(spec.feed_mean)
    test = cfg["mode"] == "test"
    c.test_mode = int(test)
    rows = int(np.asarray(ref["q"]).shape[0])
    if test or spec.n_train is None:
        n_steps = spec.n_test_mult * (rows - 2)
    else:
        n_steps = spec.n_train
    c.n_steps = n_steps
    c.cycle = spec.cycle if spec.cycle is not None else n_steps // 2
    if test:
        c.reset_max_index = 0
    elif spec.reset_max == "cycle":
        c.reset_max_index = c.cycle
    else:
        c.reset_max_index = n_steps // 2
    c.reset_max_index = min(c.reset_max_index, rows - 2)
    c.ref_mirror = int(spec.jumping)
    c.auto_reset = int(bool(cfg["auto_reset"]))
    # termination
    obs_names = cm.obs_body_names
    c.term_obspt = obs_names.index("torso")
    c.term_feet_cross = int(spec.spatial)
    ct.set_field(c, "feet_obspt", [obs_names.index("calcn_r"), obs_names.index("calcn_l")])
    c.term_height = spec.term_height
    c.term_limit_force = spec.term_limit
    c.term_acc = spec.term_acc
    # reward
    c.w_imitate, c.w_effort, c.w_action = [float(x) for x in cfg["r_weights"][:3]]
    c.action_r_scale = spec.action_r_scale
    c.max_actuation = float(cfg["max_actuation"])
    c.height = 1.80
    c.reward_use_feet = int(spec.spatial)
    c.effort_torque = int(spec.torque)
    c.effort_use_dy = int(spec.jumping)
    c.n_reward_terms = 4 if spec.torque else 5
    ref_bodies = list(ref["body_names"])
    rew_o, rew_r = [], []
    for side in ("r", "l"):
        names = ["%s_%s" % (b, side) for b in ("calcn", "femur", "tibia", "talus")]
        rew_o.append([obs_names.index(n) for n in names])
        rew_r.append([ref_bodies.index(n) for n in names])
    ct.set_field(c, "rew_obspt", np.asarray(rew_o))
    ct.set_field(c, "rew_refbody", np.asarray(rew_r))
    # PD (torque envs)
    c.use_pd = int(spec.torque)
    if spec.torque:
        pd = _PD_3D if spec.spatial else _PD_2D
        non_pelvis = [i for i in range(t.n_coords) if t.coord_pelvis_trans[i] == 0]
        c.n_pd = len(pd["x"])
        ct.set_field(c, "pd_x_coord", [non_pelvis[i] for i in pd["x"]])
        ct.set_field(c, "pd_v_coord", pd["v"])
        ct.set_field(c, "pd_kp", pd["kp"])
        ct.set_field(c, "pd_kv", pd["kv"])
    # observation
    c.use_target_obs = int(bool(cfg["use_target_obs"]))
    c.use_grf = int(bool(cfg["use_GRF"]))
    c.n_obs_bodies = len(obs_names)
    c.n_obs_body_vel = 3
    c.obs_dim = obs_dim(t, c)
    # perturbation
    c.perturb = int(bool(cfg["apply_perturbations"]))
    c.perturb_obspt = obs_names.index("torso")
    c.perturb_negative_only = int(spec.perturb_negative_only)
    c.perturb_thresh = spec.perturb_thresh
    c.perturb_force = 50.0
    return c


def obs_dim(t, c) -> int:
    n_pel = sum(1 for i in range(t.n_coords) if t.coord_pelvis_trans[i] != 0)
    n_tx = sum(1 for i in range(t.n_coords) if t.coord_pelvis_trans[i] == 1)
    d = 1 + (t.n_coords - n_pel) + 2 * t.n_coords
    if c.use_target_obs:
        d += 2 * (t.n_coords - n_tx)
    d += 3 * (c.n_obs_bodies + 1) + 3 * (c.n_obs_body_vel + 1) + 3 * t.n_muscles
    if c.use_grf:
        d += 12
    return d
