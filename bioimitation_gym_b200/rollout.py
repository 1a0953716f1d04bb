"""Device-resident rollout storage and trajectory export for the batched envs.

* `DeviceReplayBuffer` -- the reference's `ReplayBuffer` / `Dataset.sample`
  (bioimitation/learning_algorithm/datasets/replay_buffer.py:9-74, dataset.py:46-66) with the
  same field names, but the storage is torch tensors on the env's device and a whole batch of N
  transitions (one per env) is inserted per step, so the learner never round-trips through the
  host (`sample_baselines_training.py:59-87` inserts one transition per Python iteration).
* `TrajectoryRecorder` -- per-step states of selected envs, written in the layout of OpenSim's
  state storage, i.e. what `OsimModel.save_simulation` produces as `simulation_States.sto`
  (opensim_wrapper.py:334-338), so the reference's visualisation scripts keep working.

No physics here: both only move tensors the step kernel produced.
"""
from __future__ import annotations

import collections
import os
from typing import Optional, Sequence

import numpy as np

from . import refmotion

Batch = collections.namedtuple("Batch", ["observations", "actions", "rewards", "masks", "next_observations"])


class DeviceReplayBuffer:
    def __init__(self, obs_dim: int, action_dim: int, capacity: int, device="cuda", dtype=None):
        import torch
        self.torch = torch
        dtype = dtype or torch.float32
        self.capacity = int(capacity)
        mk = lambda *s: torch.empty(s, dtype=dtype, device=device)
        self.observations = mk(self.capacity, obs_dim)
        self.actions = mk(self.capacity, action_dim)
        self.rewards = mk(self.capacity)
        self.masks = mk(self.capacity)
        self.dones_float = mk(self.capacity)
        self.next_observations = mk(self.capacity, obs_dim)
        self.size = 0
        self.insert_index = 0

    def insert(self, observation, action, reward, mask, done_float, next_observation):
        """One transition (reference signature, replay_buffer.py:63-74) or a batch of N (leading dimension)."""
        torch = self.torch
        obs = torch.as_tensor(observation, device=self.observations.device, dtype=self.observations.dtype)
        if obs.dim() == 1:
            obs = obs[None]
        n = obs.shape[0]
        if n > self.capacity:
            raise ValueError("batch of %d transitions does not fit a buffer of %d" % (n, self.capacity))
        idx = (torch.arange(n, device=obs.device) + self.insert_index) % self.capacity
        as_t = lambda x, like: torch.as_tensor(x, device=like.device, dtype=like.dtype).reshape((n,) + like.shape[1:])
        self.observations[idx] = obs
        self.actions[idx] = as_t(action, self.actions)
        self.rewards[idx] = as_t(reward, self.rewards)
        self.masks[idx] = as_t(mask, self.masks)
        self.dones_float[idx] = as_t(done_float, self.dones_float)
        self.next_observations[idx] = as_t(next_observation, self.next_observations)
        self.insert_index = (self.insert_index + n) % self.capacity
        self.size = min(self.size + n, self.capacity)

    def insert_step(self, observation, action, reward, done, next_observation, time_limit_done=None):
        """Batch insert of one step of N envs: mask = 0 where the episode really ended
        (sample_baselines_training.py:72-75: a time-limit end keeps mask 1).  `observation` and
        `next_observation` must be different storage: `VecEnv.step` returns the SAME env-owned tensor every
        call, and after an auto-reset it holds the first observation of the next episode, not the terminal one
        -- use `attach` + `step_and_insert` for a VecEnv."""
        if hasattr(observation, "data_ptr") and hasattr(next_observation, "data_ptr") and \
                observation.data_ptr() == next_observation.data_ptr():
            raise ValueError("observation and next_observation alias one tensor (VecEnv.step returns its own "
                             "buffer): clone the observation before stepping, or use attach/step_and_insert")
        done_f = done.to(self.rewards.dtype)
        mask = 1.0 - done_f
        if time_limit_done is not None:
            mask = self.torch.where(time_limit_done.bool(), self.torch.ones_like(mask), mask)
        self.insert(observation, action, reward, mask, done_f, next_observation)

    # ---- feeding from a VecEnv (the learner loop of sample_baselines_training.py:59-87, batched) ----
    def attach(self, env, observation=None):
        """Start collecting from `env`: asks the step kernel for the terminal observation and the done reason
        of finished envs (BioStepExtra) and keeps a private copy of the current observation (pass the tensor
        `env.reset()` returned, default: the env's observation buffer)."""
        from . import ctables as ct
        self._env = env
        self._extra = env.enable_step_extra("terminal_obs", "done_reason")
        self._horizon_bit = ct.MACROS["BIO_DONE_HORIZON"]
        self._prev = (env.obs if observation is None else observation).detach().clone()

    def step_and_insert(self, actions):
        """env.step(actions) and insert the N transitions: next_observation of a finished env is its TERMINAL
        observation (the step kernel has already put the first observation of the next episode into the
        returned batch), mask stays 1 when the episode only hit its step limit.  Returns what env.step does."""
        torch = self.torch
        env = self._env
        obs, rew, done, info = env.step(actions)
        fin = done.bool()
        nxt = torch.where(fin[:, None], self._extra["terminal_obs"], obs)
        self.insert_step(self._prev, actions, rew, done, nxt, self._extra["done_reason"] == self._horizon_bit)
        self._prev.copy_(obs)
        return obs, rew, done, info

    def sample(self, batch_size: int, generator=None) -> Batch:
        if self.size == 0:
            raise ValueError("empty buffer")
        idx = self.torch.randint(self.size, (int(batch_size),), device=self.observations.device, generator=generator)
        return Batch(observations=self.observations[idx], actions=self.actions[idx], rewards=self.rewards[idx],
                     masks=self.masks[idx], next_observations=self.next_observations[idx])

    def __len__(self):
        return self.size


def state_labels(cm) -> list:
    """Column labels of an OpenSim 4 states storage for the compiled model: time, then value / speed
    per free coordinate, activation / fiber_length per muscle."""
    labels = ["time"]
    for name in cm.dof_names:
        labels += ["%s/value" % name, "%s/speed" % name]
    for name in cm.muscle_names:
        labels += ["%s/activation" % name, "%s/fiber_length" % name]
    return labels


class TrajectoryRecorder:
    def __init__(self, env, env_indices: Optional[Sequence[int]] = None):
        self.env = env
        self.idx = list(env_indices) if env_indices is not None else [0]
        self.rows = {i: [] for i in self.idx}

    def record(self):
        """Append the current state of the recorded envs (call after reset and after every step)."""
        st = self.env.get_state()
        dt = self.env.task.dt
        q, u = st["q"].cpu().numpy(), st["u"].cpu().numpy()
        a, lm = st["act"].cpu().numpy(), st["lm"].cpu().numpy()
        istep = st["istep"].cpu().numpy()
        for i in self.idx:
            row = [istep[i] * dt]
            for d in range(q.shape[1]):
                row += [q[i, d], u[i, d]]
            for k in range(a.shape[1]):
                row += [a[i, k], lm[i, k]]
            self.rows[i].append(row)

    def save_simulation(self, base_dir: str, env_index: Optional[int] = None) -> str:
        """Write `<base_dir>/simulation_States.sto` (reference opensim_wrapper.py:334-338) for one recorded env."""
        i = self.idx[0] if env_index is None else env_index
        os.makedirs(base_dir, exist_ok=True)
        path = os.path.join(base_dir, "simulation_States.sto")
        refmotion.write_storage(path, "simulation_States", state_labels(self.env.cm), np.asarray(self.rows[i]))
        return path
