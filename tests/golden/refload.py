"""Loader for the UNMODIFIED reference (lorenzrichter/path-space-PDE-solver) on CPU.

Test infrastructure only.  Used in the build container (where /root/reference is
mounted) to generate the golden fixtures in this directory; nothing on the GPU
box imports it.  Recipe = SURVEY.md Appendix B: stub matplotlib, swap the
hard-coded ``pt.device('cuda')`` literal (solver.py:36,573,947; problems.py:11;
utilities.py:293,440) for the requested device, exec the four files as modules.
No reference source is copied into this repository.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("PSPDE_REF", "/root/reference")


def load_reference(device="cpu"):
    if not os.path.isdir(REF_ROOT):
        raise FileNotFoundError("reference tree not mounted at %s" % REF_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    mods = {}
    for name in ("function_space", "problems", "utilities", "solver"):
        path = os.path.join(REF_ROOT, name + ".py")
        with open(path) as fh:
            src = fh.read()
        src = src.replace("pt.device('cuda')", "pt.device(%r)" % device)
        mod = types.ModuleType(name)
        mod.__file__ = path
        sys.modules[name] = mod
        exec(compile(src, path, "exec"), mod.__dict__)
        mods[name] = mod
    return mods
