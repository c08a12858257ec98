"""Launcher: run a script of the unmodified reference with this repository's hot path.

    cd /path/to/enerf && python -m enerf_b200.run_reference main_nerf.py --config configs/... [--cuda_ray --ff]

Equivalent to `python main_nerf.py ...` with the import hook of enerf_b200/dropin_hook.py installed first."""
import os
import runpy
import sys

from . import dropin_hook


def main():
    if len(sys.argv) < 2:
        raise SystemExit("usage: python -m enerf_b200.run_reference <script.py> [args...]")
    script = os.path.abspath(sys.argv[1])
    dropin_hook.install()
    sys.argv = sys.argv[1:]
    sys.path.insert(0, os.path.dirname(script))          # what `python script.py` would have done
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
