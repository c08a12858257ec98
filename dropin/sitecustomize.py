"""Runs at interpreter start-up when this directory is on PYTHONPATH: installs the import hook that points the reference's bare-name
imports at enerf_b200 (see enerf_b200/dropin_hook.py), then hands over to any other `sitecustomize` further down the path so that this
file shadows nothing."""
import importlib.machinery
import importlib.util
import os
import sys

try:
    _mode = os.environ.get("ENERF_DROPIN", "all").lower()
    if _mode != "off":
        from enerf_b200 import dropin_hook
        dropin_hook.install(mirrors=(_mode != "packages"))
except Exception as e:  # noqa: BLE001  (never break interpreter start-up)
    sys.stderr.write(f"[enerf_b200 dropin] import hook not installed: {type(e).__name__}: {e}\n")

_here = os.path.dirname(os.path.abspath(__file__))
for _p in sys.path:
    try:
        if not _p or os.path.abspath(_p) == _here:
            continue
        _f = os.path.join(_p, "sitecustomize.py")
        if os.path.isfile(_f):
            _spec = importlib.util.spec_from_file_location("_chained_sitecustomize", _f)
            _mod = importlib.util.module_from_spec(_spec)
            _spec.loader.exec_module(_mod)
            break
    except Exception as e:  # noqa: BLE001
        sys.stderr.write(f"[enerf_b200 dropin] chained sitecustomize failed: {type(e).__name__}: {e}\n")
        break
