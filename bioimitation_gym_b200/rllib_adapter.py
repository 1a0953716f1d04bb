"""RLlib `VectorEnv` view of the batched backend.

The reference trains with RLlib PPO on one OpenSim env per rollout worker
(tests/sample_rllib_training.py, bioimitation/__init__.py:63-70 registers the classes with
`ray.tune.registry.register_env`).  RLlib can also drive a whole batch per worker through
`ray.rllib.env.VectorEnv` (`vector_reset`, `reset_at`, `vector_step`, `get_sub_environments`);
this adapter offers that interface on top of `VecEnv`, so one worker steps thousands of envs per
call.  ray is optional: when it is installed the class derives from `VectorEnv`, otherwise it is a
plain class with the same methods (duck-typed), which is what the tests use.

Episode ends: the step kernel resets finished envs itself (reference-state initialisation,
muscle_walking_imitation_env2D.py:133-156), so the observation of a finished env returned by
`vector_step` is already the first observation of its next episode and `reset_at(i)` only hands
that row out again; nothing is launched per env.

Every array handed to RLlib is a COPY: `VecEnv.step_np` returns views of two alternating page-locked
buffers that the GPU overwrites in place two calls later, while RLlib's sample collectors keep
observation references for a whole rollout fragment.

Supported interface: the pre-gymnasium `VectorEnv` of ray <= 2.2 (4-tuple `vector_step`, `reset_at(index)`),
i.e. the ray 1.8 the reference pins (scripts/requirements.txt); `vector_step_v26` returns the
terminated / truncated split newer ray versions expect (truncated = the episode hit its step limit).
"""
from __future__ import annotations

from typing import Any, List, Mapping, Optional

import numpy as np

from . import envs as _envs
from .backend import VecEnv

try:  # pragma: no cover - ray is not part of this image
    from ray.rllib.env.vector_env import VectorEnv as _Base  # type: ignore
except Exception:  # noqa: BLE001
    _Base = object


class BioVectorEnv(_Base):
    def __init__(self, env_id: str, config: Optional[Mapping[str, Any]] = None, backend_env=None):
        cfg = dict(config or {})
        self.env = backend_env if backend_env is not None else VecEnv(env_id, cfg)
        self.num_envs = self.env.num_envs
        self.observation_space = _envs._make_box(-np.inf * np.ones(self.env.obs_dim, dtype=np.float32),
                                                 np.inf * np.ones(self.env.obs_dim, dtype=np.float32))
        self.action_space = _envs._make_box(self.env.action_low.astype(np.float32),
                                            self.env.action_high.astype(np.float32))
        if _Base is not object:  # pragma: no cover
            super().__init__(self.observation_space, self.action_space, self.num_envs)
        self._obs = None

    def vector_reset(self) -> List[np.ndarray]:
        self._obs = self.env.reset_np().copy()
        return [o.copy() for o in self._obs]

    def reset_at(self, index: Optional[int] = None) -> np.ndarray:
        if self._obs is None:
            self.vector_reset()
        return self._obs[0 if index is None else int(index)].copy()

    def vector_step(self, actions):
        # host buffers in, host buffers out: page-locked arrays the step kernel reads / writes in place
        a = np.asarray(actions, dtype=np.float32 if self.env.dtype == self.env.torch.float32 else np.float64)
        obs, rew, done, info = self.env.step_np(a)
        self._obs = obs.copy()
        terms = info["all_rewards"]
        infos = [{"all_rewards": terms[i].tolist()} for i in range(self.num_envs)]
        return [o.copy() for o in self._obs], rew.tolist(), done.tolist(), infos

    def vector_step_v26(self, actions):
        """(obs, rewards, terminateds, truncateds, infos): `truncated` = the env ended on its step limit
        (BIO_DONE_HORIZON), from the step kernel's per-env done reason."""
        from . import ctables as ct
        if getattr(self, "_reason", None) is None:
            self._reason = self.env.enable_step_extra("done_reason")["done_reason"]
        obs, rew, done, infos = self.vector_step(actions)
        reason = self._reason.cpu().numpy()
        trunc = (reason == ct.MACROS["BIO_DONE_HORIZON"])
        term = np.asarray(done, dtype=bool) & ~trunc
        return obs, rew, term.tolist(), trunc.tolist(), infos

    def get_sub_environments(self):
        return []          # the envs live on the device; there are no per-env Python objects

    def get_unwrapped(self):   # older RLlib name
        return []

    def close(self):
        self.env.close()
