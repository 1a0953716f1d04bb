"""Single-env views with the reference's gym call pattern.

The 17 classes keep the reference's names (``bioimitation/__init__.py:23-133``)
and signature ``Env(config)``; ``reset(obs_as_dict=False) -> np.ndarray`` and
``step(action, obs_as_dict=False) -> [obs, reward, done, {'all_rewards': [...]}]``
(reference ``opensim_environment.py:88-113``).  Each instance is a ``VecEnv``
of ``config['num_envs']`` (default 1) envs on the GPU; with ``num_envs == 1`` the
return types are the reference's (1-D numpy obs, python float, python bool,
list), otherwise batched numpy arrays.  RLlib / SAC callers
(``tests/sample_rllib_training.py:42-62``, ``tests/sample_baselines_training.py:59-87``)
only use ``action_space``, ``observation_space``, ``reset`` and ``step``.
"""
from __future__ import annotations

from typing import Any, Dict, Mapping, Optional

import numpy as np

from . import tasks


class Box:
    """Minimal stand-in for ``gym.spaces.Box`` (gym is not a dependency)."""

    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


def _make_box(low, high):
    try:  # use the real thing when gym / gymnasium is installed
        from gym import spaces  # type: ignore
        return spaces.Box(np.asarray(low, dtype=np.float32), np.asarray(high, dtype=np.float32))
    except Exception:
        try:
            from gymnasium import spaces  # type: ignore
            return spaces.Box(np.asarray(low, dtype=np.float32), np.asarray(high, dtype=np.float32))
        except Exception:
            return Box(low, high)


class Specification:
    """reference opensim_environment.py:9-14"""

    def __init__(self, timestep_limit):
        self.id = 0
        self.timestep_limit = timestep_limit


class BioImitationEnv:
    """Common implementation of the reference's env classes."""

    ENV_ID: str = ""
    metadata: Dict[str, Any] = {}

    def __init__(self, config: Optional[Mapping[str, Any]] = None):
        from .backend import VecEnv
        cfg = tasks.merged_config(config)
        if config is None or "auto_reset" not in dict(config):
            cfg["auto_reset"] = int(cfg["num_envs"]) > 1
        self.config = cfg
        self.vec = VecEnv(self.ENV_ID, cfg)
        self.num_envs = self.vec.num_envs
        # action bounds = actuator min/max (opensim_wrapper.py:39-53); obs bounds +-inf
        self.action_space = _make_box(self.vec.action_low, self.vec.action_high)
        d = self.vec.obs_dim
        self.observation_space = _make_box([-np.inf] * d, [np.inf] * d)
        self.timestep_limit = 1e10
        self.spec = Specification(self.timestep_limit)
        self.spec.action_space = self.action_space
        self.spec.observation_space = self.observation_space
        self._np_dtype = np.float32 if str(self.vec.dtype).endswith("float32") else np.float64

    # ---- reference API -------------------------------------------------
    def reset(self, obs_as_dict: bool = False):
        obs = self.vec.reset_np()
        return self._format_obs(obs, obs_as_dict)

    def step(self, action, obs_as_dict: bool = False):
        # numpy in / numpy out through page-locked buffers the step kernel reads and writes in place
        a = np.asarray(action, dtype=self._np_dtype).reshape(self.num_envs, -1)
        obs, rew, done, info = self.vec.step_np(a)
        terms = info["all_rewards"]
        if self.num_envs == 1:
            return [self._format_obs(obs, obs_as_dict), float(rew[0]), bool(done[0]),
                    {"all_rewards": [float(x) for x in terms[0]]}]
        return [self._format_obs(obs, obs_as_dict), rew, done, {"all_rewards": terms}]

    def get_observation_space_size(self):
        return self.vec.obs_dim

    def get_action_space_size(self):
        return self.vec.n_act

    def render(self, mode="human", close=False):
        return None

    def seed(self, seed=None):
        return [seed]

    def close(self):
        self.vec.close()

    # ---- helpers ---------------------------------------------------------
    def _format_obs(self, obs: np.ndarray, as_dict: bool):
        if as_dict:
            if self.num_envs != 1:
                raise ValueError("obs_as_dict needs num_envs == 1")
            return self.obs_to_dict(obs[0])
        return obs[0].copy() if self.num_envs == 1 else obs.copy()

    def obs_to_dict(self, flat: np.ndarray) -> Dict[str, Any]:
        """Nested observation dict in the reference's insertion order
        (reference muscle_walking_imitation_env2D.py:158-230)."""
        cm, task = self.vec.cm, self.vec.task
        t = cm.tables
        names = cm.coord_names
        pel = [i for i in range(t.n_coords) if t.coord_pelvis_trans[i] != 0]
        tx = [i for i in range(t.n_coords) if t.coord_pelvis_trans[i] == 1]
        o = 0
        out: Dict[str, Any] = {"phase": float(flat[o])}
        o += 1
        out["coordinate_pos"] = {}
        for i, n in enumerate(names):
            if i not in pel:
                out["coordinate_pos"][n] = float(flat[o]); o += 1
        for key in ("coordinate_vel", "coordinate_acc"):
            out[key] = {}
            for n in names:
                out[key][n] = float(flat[o]); o += 1
        if task.use_target_obs:
            for key in ("target_coordinate_pos", "target_coordinate_vel"):
                out[key] = {}
                for i, n in enumerate(names):
                    if i not in tx:
                        out[key][n] = float(flat[o]); o += 1
        out["body_pos"] = {}
        for b in cm.obs_body_names[:task.n_obs_bodies] + ["center_of_mass"]:
            out["body_pos"][b] = [float(x) for x in flat[o:o + 3]]; o += 3
        out["body_vel"] = {}
        for b in cm.obs_body_names[:task.n_obs_body_vel] + ["center_of_mass"]:
            out["body_vel"][b] = [float(x) for x in flat[o:o + 3]]; o += 3
        if t.n_muscles:
            out["muscles"] = {}
            for mname in cm.muscle_names:
                out["muscles"][mname] = dict(activation=float(flat[o]), fiber_length=float(flat[o + 1]),
                                             fiber_velocity=float(flat[o + 2]))
                o += 3
        if task.use_grf:
            out["contact_forces"] = {}
            for f in ("foot_r", "foot_l"):
                out["contact_forces"][f] = [float(x) for x in flat[o:o + 6]]; o += 6
        assert o == flat.shape[0]
        return out


def _cls(env_id: str):
    spec = tasks.ENV_SPECS[env_id]
    return type(spec.cls_name, (BioImitationEnv,), {"ENV_ID": env_id, "__doc__":
                "Drop-in for the reference class %s (%s)." % (spec.cls_name, env_id)})


ENV_CLASSES = {env_id: _cls(env_id) for env_id in tasks.ENV_SPECS}
globals().update({c.__name__: c for c in ENV_CLASSES.values()})


def make(env_id: str, config: Optional[Mapping[str, Any]] = None):
    """``gym.make(env_id, config=config)`` without gym."""
    return ENV_CLASSES[env_id](config)


def register_all() -> Dict[str, bool]:
    """Register the 17 IDs with gym / gymnasium and ray.tune when those packages
    are installed (reference bioimitation/__init__.py:23-143); returns what was
    registered.  None of them is required."""
    done = {"gym": False, "gymnasium": False, "ray": False}
    for mod in ("gym", "gymnasium"):
        try:
            reg = __import__(mod + ".envs.registration", fromlist=["register"]).register
            for env_id, cls in ENV_CLASSES.items():
                try:
                    reg(id=env_id, entry_point="bioimitation_gym_b200.envs:%s" % cls.__name__)
                except Exception:
                    pass
            done[mod] = True
        except Exception:
            pass
    try:
        from ray.tune.registry import register_env  # type: ignore
        for env_id, cls in ENV_CLASSES.items():
            register_env(env_id, lambda config, _c=cls: _c(config))
        done["ray"] = True
    except Exception:
        pass
    return done
