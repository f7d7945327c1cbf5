"""Import the real scikit-fusion sources from /root/reference *in memory*.

Only usable in the build container (the GPU box has no /root/reference).  Nothing is copied
into this repository: the module sources are read, four py3.12 / numpy-2 compatibility edits
 are applied to the text in memory, and the result is exec'd as a private package
named ``_skfusion_reference``:

  (a) ``from collections import ... Iterable``   -> collections.abc        fusion_graph.py:4
  (b) generated ``skfusion/version.py`` is absent -> skfusion/__init__.py is bypassed (only skfusion.fusion is loaded)
  (c) ``entry != []`` on an ndarray               -> explicit list test     _dfmf.py:74, _dfmc.py:74
  (d) ``np.float``                                -> ``float``              _dfmf.py:296,428 _dfmc.py:366

Used by make_golden.py and by the (container-only) oracle pinning test.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("SKFUSION_REFERENCE", "/root/reference")
PKG = "_skfusion_reference"


def available():
    """REF_ROOT may be repointed (module attribute) before the first load(), e.g. at baseline/_ref on the GPU box."""
    return os.path.isdir(os.path.join(REF_ROOT, "skfusion", "fusion"))


def _patch(relpath, text):
    if relpath.endswith("fusion_graph.py"):
        text = text.replace(
            "from collections import defaultdict, OrderedDict, Iterable",
            "from collections import defaultdict, OrderedDict\nfrom collections.abc import Iterable")
    if relpath.endswith("_dfmf.py") or relpath.endswith("_dfmc.py"):
        text = text.replace("if entry != []}", "if not (isinstance(entry, list) and len(entry) == 0)}")
        text = text.replace("np.finfo(np.float)", "np.finfo(float)")
    return text


def _exec_module(fullname, path, is_pkg):
    spec = importlib.util.spec_from_loader(fullname, loader=None, is_package=is_pkg)
    mod = types.ModuleType(fullname)
    mod.__file__ = path
    mod.__spec__ = spec
    if is_pkg:
        mod.__path__ = [os.path.dirname(path)]
        mod.__package__ = fullname
    else:
        mod.__package__ = fullname.rpartition(".")[0]
    sys.modules[fullname] = mod
    with open(path, "r", encoding="utf-8") as fh:
        src = _patch(path, fh.read())
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def load():
    """Return the reference's ``skfusion.fusion`` package (patched in memory)."""
    if PKG + ".fusion" in sys.modules:
        return sys.modules[PKG + ".fusion"]
    if not available():
        raise RuntimeError("reference sources not found under %s" % REF_ROOT)
    base = os.path.join(REF_ROOT, "skfusion")
    root = types.ModuleType(PKG)
    root.__path__ = [base]
    root.__package__ = PKG
    sys.modules[PKG] = root
    # package shells first (so relative imports resolve), then leaves / package bodies bottom-up
    shells = {"fusion": "fusion/__init__.py", "fusion.base": "fusion/base/__init__.py",
              "fusion.decomposition": "fusion/decomposition/__init__.py"}
    for name, rel in shells.items():
        full = PKG + "." + name
        m = types.ModuleType(full)
        m.__path__ = [os.path.dirname(os.path.join(base, rel))]
        m.__package__ = full
        m.__file__ = os.path.join(base, rel)
        sys.modules[full] = m

    def body(name):
        full = PKG + "." + name
        path = os.path.join(base, shells[name])
        with open(path, "r", encoding="utf-8") as fh:
            exec(compile(fh.read(), path, "exec"), sys.modules[full].__dict__)

    for leaf in ("fusion/base/base.py", "fusion/base/fusion_graph.py"):
        _exec_module(PKG + "." + leaf[:-3].replace("/", "."), os.path.join(base, leaf), False)
    body("fusion.base")
    for leaf in ("fusion/decomposition/_init.py", "fusion/decomposition/_dfmf.py",
                 "fusion/decomposition/_dfmc.py", "fusion/decomposition/dfmf.py",
                 "fusion/decomposition/dfmc.py"):
        _exec_module(PKG + "." + leaf[:-3].replace("/", "."), os.path.join(base, leaf), False)
    body("fusion.decomposition")
    body("fusion")
    return sys.modules[PKG + ".fusion"]


def functions():
    """(dfmf, dfmc, transform, initialize) free functions of the reference."""
    load()
    d = sys.modules[PKG + ".fusion.decomposition._dfmf"]
    c = sys.modules[PKG + ".fusion.decomposition._dfmc"]
    i = sys.modules[PKG + ".fusion.decomposition._init"]
    return d.dfmf, c.dfmc, d.transform, i.initialize
