"""Env-ID -> (model tables, reference-motion tables, task config).

Mirrors what each reference env constructor assembles before it builds
``OsimEnv`` (reference ``muscle_walking_imitation_env2D.py:18-81``): pick the
subject directory, make/load the predictive model, read the four reference
``.sto`` frames.  Here all three come from the compiled fixtures in
``bioimitation_gym_b200/data`` (or from user-supplied tables).
"""
from __future__ import annotations

from typing import Any, Mapping, Optional

from . import assets, tasks


def build_env_tables(env_id: str, config: Optional[Mapping[str, Any]] = None,
                     model=None, ref=None):
    if env_id not in tasks.ENV_SPECS:
        raise KeyError("unknown env id %r (known: %s)" % (env_id, ", ".join(tasks.ENV_SPECS)))
    spec = tasks.ENV_SPECS[env_id]
    cm = model if model is not None else assets.load_model(spec.model)
    if ref is None:
        try:
            ref = assets.load_ref(spec.ref)
        except FileNotFoundError as e:
            raise FileNotFoundError(
                "no reference motion %r for %s: the reference repository does not ship it "
                "(SURVEY 0.3); pass ref=dict(q,u,body_pos,com_pos,body_names)" %
                (spec.ref, env_id)) from e
    task = tasks.make_task_config(spec, cm, ref, config or {})
    return spec, cm, ref, task
