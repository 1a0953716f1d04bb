"""B200-native batched physics step for bioimitation-gym (drop-in backend).

Only the hot path of the reference is here: reset/step of the 17 imitation envs
for thousands of independent envs per GPU, behind the reference's env IDs,
config keys and observation/action/reward layout.  See DESIGN.md.
"""
from .tasks import ENV_SPECS, DEFAULT_ENV_CONFIG  # noqa: F401

__all__ = ["ENV_SPECS", "DEFAULT_ENV_CONFIG", "make_vec", "make"]


def make_vec(env_id, config=None, **kw):
    """Batched env (torch CUDA tensors in/out)."""
    from .backend import VecEnv
    return VecEnv(env_id, config, **kw)


def make(env_id, config=None, **kw):
    """Single-env view with the reference's gym call pattern."""
    from .envs import make as _make
    return _make(env_id, config, **kw)
