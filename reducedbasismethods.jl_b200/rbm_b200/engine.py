"""Process-wide default context: one process drives one GPU (LOCAL_RANK selects it)."""
from __future__ import annotations

import os

from ._lib import Context

_ctx = None


def default_context() -> Context:
    global _ctx
    if _ctx is None:
        _ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _ctx


def set_default_context(ctx: Context | None):
    global _ctx
    _ctx = ctx
