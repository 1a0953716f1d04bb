"""`import bioimitation` shim: the reference's import path on top of the B200 backend.

The reference registers its 17 env IDs with gym (and ray.tune) when `bioimitation` is imported
(reference bioimitation/__init__.py:23-143) and its callers do

    import gym, bioimitation
    env = gym.make("MuscleWalkingImitation2D-v0", config=cfg)        # tests/test_env.py:15-26

This package keeps that working unmodified: importing it registers the same IDs (when gym /
gymnasium / ray are installed; none is required) with entry points under the reference's own module
paths (`bioimitation.imitation_envs.envs.muscle.planar.muscle_walking_imitation_env2D:
MuscleWalkingImitationEnv2D`, ...), which resolve to the drop-in classes of
`bioimitation_gym_b200.envs`.  Nothing of the reference's OpenSim code is here.
"""
from __future__ import annotations

import re
import sys
import types

from bioimitation_gym_b200 import envs as _envs
from bioimitation_gym_b200 import tasks as _tasks


def _module_path(spec) -> str:
    """Reference module of an env class, e.g. TorqueLockedKneeImitationEnv3D ->
    bioimitation.imitation_envs.envs.torque.spatial.torque_locked_knee_imitation_env3D."""
    words = re.findall(r"[A-Z][a-z]*|\d+[A-Z]", spec.cls_name)          # ..., 'Env', '2D'
    snake = "_".join(w.lower() for w in words[:-1]) + words[-1]         # ..._env + 2D
    return "bioimitation.imitation_envs.envs.%s.%s.%s" % (
        "torque" if spec.torque else "muscle", "spatial" if spec.spatial else "planar", snake)


def _install_modules():
    made = {}
    for env_id, spec in _tasks.ENV_SPECS.items():
        path = _module_path(spec)
        parts = path.split(".")
        for k in range(2, len(parts) + 1):               # parents below `bioimitation`
            name = ".".join(parts[:k])
            if name not in sys.modules:
                mod = types.ModuleType(name, "reference module path, served by bioimitation_gym_b200")
                mod.__path__ = []                        # a package as far as importlib is concerned
                sys.modules[name] = mod
                setattr(sys.modules[".".join(parts[:k - 1])], parts[k - 1], mod)
        cls = _envs.ENV_CLASSES[env_id]
        setattr(sys.modules[path], cls.__name__, cls)
        made[env_id] = "%s:%s" % (path, cls.__name__)
    return made


ENTRY_POINTS = _install_modules()
globals().update({c.__name__: c for c in _envs.ENV_CLASSES.values()})


def register_all():
    """gym / gymnasium `register` and ray.tune `register_env` for the 17 IDs (what the reference's
    __init__ does at import); returns which registries were found."""
    done = {"gym": False, "gymnasium": False, "ray": False}
    for mod in ("gym", "gymnasium"):
        try:
            reg = __import__(mod + ".envs.registration", fromlist=["register"]).register
        except Exception:
            continue
        for env_id, ep in ENTRY_POINTS.items():
            try:
                reg(id=env_id, entry_point=ep)
            except Exception:                           # already registered
                pass
        done[mod] = True
    try:
        from ray.tune.registry import register_env  # type: ignore
        for env_id, cls in _envs.ENV_CLASSES.items():
            register_env(env_id, lambda config, _c=cls: _c(config))
        done["ray"] = True
    except Exception:
        pass
    return done


REGISTERED = register_all()
make = _envs.make
