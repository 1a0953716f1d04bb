"""Batched inverse-dynamics operator set on the GPU.

B200 equivalent of the reference's only native component, the boost.python class
`InverseDynamics` (bioimitation/imitation_envs/inverse_dynamics/inverse_dynamics.cpp:44-200),
and of the controllers built on it in tests/example_position_control.py:113-198.

The reference wraps one OpenSim model per process and takes Python lists; here N independent
states are evaluated per call, and every operator runs INSIDE the native library
(`bio_id_apply` / `bio_id_multiply_m` / `bio_id_multiply_minv` / `bio_id_residual`,
include/bio_b200.h): joint-space inertia from the composite-inertia pass and the sparse L^T D L
factorisation along the kinematic tree that the step kernel's own solve uses -- no cuBLAS /
cuSOLVER, torch only allocates the tensors.  Method names, argument order (`t` first) and sign
conventions follow the reference:

    M(q) qddot + c(q, qdot) = g(q) + tau_applied        (inverse_dynamics.cpp:122,139)
    calculateTotalForces   f  with  M qddot + f = tau    (:96-121)
    calculateResidualForces   M qddot + f_internal - f_applied   (:65-94)

Like the reference (muscles "setAppliesForce(false)", :50-52) the operators see no muscle
forces: they are evaluated on the coordinate-actuated variant of the model (contact spheres
and coordinate limits stay applied forces).  Coordinates are the free generalized coordinates
of the compiled model, in its dof order (`dof_names`); locked coordinates are eliminated at
compile time.  Lists / 1-D inputs are treated as a batch of one and returned as lists.
"""
from __future__ import annotations

import copy
from typing import Optional

from . import assets, ctables as ct, tasks
from .backend import VecEnv

_TORQUE_ENV = {"2D": "TorqueWalkingImitation2D-v0", "3D": "TorqueWalkingImitation3D-v0",
               "2D_prosthetic": "TorqueLockedKneeImitation2D-v0"}


def _bare_copy(cm):
    """The same multibody tree with no force elements (no spheres, limits)."""
    out = copy.copy(cm)
    out.tables = type(cm.tables).from_buffer_copy(cm.tables)
    out.tables.n_spheres = 0
    out.tables.n_limits = 0
    return out


class InverseDynamics:
    def __init__(self, model: str = "2D", num_envs: int = 1, device: int = 0, dtype: str = "float64"):
        env_id = _TORQUE_ENV.get(model, model)
        spec = tasks.ENV_SPECS[env_id]
        if not spec.torque:
            raise ValueError("InverseDynamics needs a coordinate-actuated env id (muscles apply no force in "
                             "the reference's InverseDynamics either); got %r" % env_id)
        cfg = dict(num_envs=int(num_envs), device=device, dtype=dtype, auto_reset=False)
        self.full = VecEnv(env_id, cfg)
        self.bare = VecEnv(env_id, cfg, model=_bare_copy(assets.load_model(spec.model)))
        self.num_envs = int(num_envs)
        self.n_dof = self.full.n_dof
        self.dof_names = list(self.full.cm.dof_names)
        self.torch = self.full.torch

    # ------------------------------------------------------------------ helpers
    def _in(self, x):
        torch = self.torch
        t = torch.as_tensor(x, dtype=self.full.dtype, device=self.full.device)
        if t.dim() == 1:
            t = t[None]
        if t.shape != (self.num_envs, self.n_dof):
            raise ValueError("expected [%d, %d], got %s" % (self.num_envs, self.n_dof, tuple(t.shape)))
        return t.contiguous()

    def _out(self, t, like):
        if isinstance(like, (list, tuple)):
            return t[0].tolist()
        return t

    def _eval(self, env, q, u):
        env.set_state(dict(q=q, u=u))
        zeros = self.torch.zeros((self.num_envs, env.n_act), dtype=env.dtype, device=env.device)
        return env.eval_debug(zeros)

    def mass_matrix(self, q):
        """M(q), [N, n_dof, n_dof] (the reference reads it through calcM, example_position_control.py:166-173)."""
        qq = self._in(q)
        return self._eval(self.bare, qq, self.torch.zeros_like(qq))["mass_matrix"]

    # ------------------------------------------------------------------ reference API
    def setStateAndRealizeDynamics(self, t, q, qDot):
        """inverse_dynamics.cpp:55-61; kept for call compatibility (every operator sets its own state)."""
        self.full.set_state(dict(q=self._in(q), u=self._in(qDot)))

    def calculateResidualForces(self, t, q, qDot, qDDot):
        """tau_residual = M qddot + f_internal - f_applied (inverse_dynamics.cpp:65-94)."""
        qq, uu, aa = self._in(q), self._in(qDot), self._in(qDDot)
        self.full.set_state(dict(q=qq, u=uu))
        return self._out(self.full.id_apply("residual", aa), q)

    def calculateTotalForces(self, t, q, qDot):
        """f with M qddot + f = tau: Coriolis - gravity - contact - limits (inverse_dynamics.cpp:96-121)."""
        return self._out(self._eval(self.full, self._in(q), self._in(qDot))["bias"], q)

    def calculateGravity(self, t, q):
        """g with M qddot + c = g + tau (inverse_dynamics.cpp:123-137)."""
        qq = self._in(q)
        return self._out(-self._eval(self.bare, qq, self.torch.zeros_like(qq))["bias"], q)

    def calculateCoriolis(self, t, q, qDot):
        """c with M qddot + c = g + tau (inverse_dynamics.cpp:139-156)."""
        qq, uu = self._in(q), self._in(qDot)
        b1 = self._eval(self.bare, qq, uu)["bias"].clone()
        b0 = self._eval(self.bare, qq, self.torch.zeros_like(qq))["bias"]
        return self._out(b1 - b0, q)

    def _at(self, q):
        qq = self._in(q)
        self.bare.set_state(dict(q=qq, u=self.torch.zeros_like(qq)))
        return qq

    def multiplyByM(self, t, q, a):
        """M(q) a (inverse_dynamics.cpp:158-174)."""
        self._at(q)
        return self._out(self.bare.id_apply("multiply_m", self._in(a)), q)

    def multiplyByMInv(self, t, q, tau):
        """M(q)^-1 tau (inverse_dynamics.cpp:176-192)."""
        self._at(q)
        return self._out(self.bare.id_apply("multiply_minv", self._in(tau)), q)

    # ------------------------------------------------------------------ controllers
    def computed_torque(self, t, q, qDot, qDDot_des, tau_pd):
        """tau = M tau_pd + ID(q, qdot, qddot_des)  (example_position_control.py:113-141)."""
        qq = self._in(q)
        res = self.calculateResidualForces(t, qq, self._in(qDot), self._in(qDDot_des))
        return self._out(self.multiplyByM(t, qq, self._in(tau_pd)) + res, q)

    def stable_pd(self, t, q, qDot, tau_pd, Kd, step_size):
        """Stable PD (example_position_control.py:143-190): qddot = (M + Kd h)^-1 (tau_pd - ID(q, qdot, 0)),
        tau = tau_pd - Kd h qddot, with the reference's diagonal gain (`np.diagflat([Kd * step_size] * n)`,
        :181): Kd is a scalar, [n_dof] or [N, n_dof]."""
        torch = self.torch
        qq, uu, tp = self._in(q), self._in(qDot), self._in(tau_pd)
        res = self.calculateResidualForces(t, qq, uu, torch.zeros_like(qq))
        kd = torch.as_tensor(Kd, dtype=qq.dtype, device=qq.device)
        if kd.dim() > 2 or (kd.dim() == 2 and tuple(kd.shape) != (self.num_envs, self.n_dof)):
            raise ValueError("Kd must be a scalar, [n_dof] or [N, n_dof] (diagonal gain)")
        shift = (kd * float(step_size)).expand(self.num_envs, self.n_dof).contiguous()
        self._at(qq)
        qdd = self.bare.id_apply("solve_shifted", tp - res, shift=shift)
        return self._out(tp - shift * qdd, q)

    def close(self):
        self.full.close()
        self.bare.close()
