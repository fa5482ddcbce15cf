"""The reference's error convention at the drop-in boundary.

Inside the package a failed native call raises ``GravityNativeError`` (code + message of ``gv2_last_error``).  The modules
that mirror the reference's import paths (``gravity2_b200.app.utils.*``) wrap their functions with ``reference_errors``:
a native failure then ends the run the way the reference ends it, ``raise_gravity_error(msg)`` ->
``SystemExit("FATAL ERROR: ...")`` (app/utils/error_handlers.py:3-4).  Python-level errors the reference itself raises
(ZeroDivisionError of shared_norm_pphmm_ratio, ValueError / IndexError of bad shapes) pass through unchanged.
"""
from __future__ import annotations

import functools

from ._lib import GravityNativeError


def raise_gravity_error(msg):
    """app/utils/error_handlers.py:3-4 (the reference colours the prefix red with termcolor when it is installed)."""
    try:
        from termcolor import colored
        prefix = colored("FATAL ERROR", "red")
    except Exception:
        prefix = "FATAL ERROR"
    raise SystemExit(f"{prefix}: {msg}")


def reference_errors(fn):
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        try:
            return fn(*args, **kwargs)
        except GravityNativeError as e:
            raise_gravity_error(f"libgravity_b200 ({fn.__name__}): {e}")
    return wrapped
