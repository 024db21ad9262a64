"""Launcher for the reference's scripts with the mirror ahead of the reference checkout.

  python -m crnerf_b200.run /path/to/CR-NeRF-PyTorch/eval.py --root_dir ... (the script's own flags)

``python eval.py`` puts the script's directory at ``sys.path[0]``, ahead of ``PYTHONPATH``, so
the reference's own ``models`` / ``losses`` would win the import.  This launcher runs the
script as ``__main__`` (``runpy.run_path``) with ``sys.path = [mirror, script dir, ...]``:
``models.rendering`` / ``models.nerf`` / ``models.linearStyleTransfer`` /
``models.nerf_decoder_stylenerf`` / ``losses`` resolve to the mirror, every other module of the
reference (``models.esrgan``, ``models.networks``, ``models.lightweight_seg``, ``datasets``,
``utils``, ``opt`` ...) to the reference's unmodified files (see ``models/__init__.py``).
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    script = os.path.abspath(argv[0])
    if not os.path.isfile(script):
        raise SystemExit(f"crnerf_b200.run: no such script: {script}")
    mirror = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script_dir = os.path.dirname(script)
    sys.path[:] = [mirror, script_dir] + [p for p in sys.path if p not in (mirror, script_dir, "")]
    for name in [m for m in sys.modules if m == "models" or m.startswith("models.") or m == "losses"]:
        del sys.modules[name]                      # re-resolve against the new path order
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
