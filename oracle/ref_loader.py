"""Loaders for the real reference.  TEST INFRASTRUCTURE (see oracle/__init__.py).

* ``load_ref_dedup()`` loads oracle/_ref/_set_redundant_lpj_to_low_CPU*.so — the reference's only
  native routine, compiled by ``make -C oracle ref`` from the .pyx where it lies under
  /root/reference.  The .so travels to the GPU box; /root/reference does not.
* ``import_reference()`` imports the unmodified python reference from /root/reference (build
  container only).  h5py and munch, which the reference imports at module level
  (tvo/utils/__init__.py:2, tvo/utils/parallel.py:7, tvo/exp/_experiments.py:30) but which are
  absent from this image, are satisfied by empty stub modules.  Used by tools/make_golden.py to
  generate tests/golden/*.npz and by the `needs_reference` tests; never at GPU-box run time.
"""
import glob
import importlib.machinery
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("TVO_REFERENCE_ROOT", "/root/reference")
_EXT_NAME = "_set_redundant_lpj_to_low_CPU"


def _ext_path():
    hits = sorted(glob.glob(os.path.join(_HERE, "_ref", _EXT_NAME + "*.so")))
    return hits[0] if hits else None


def _load_ext(fullname):
    path = _ext_path()
    if path is None:
        return None
    loader = importlib.machinery.ExtensionFileLoader(fullname, path)
    spec = importlib.util.spec_from_file_location(fullname, path, loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


_dedup_mod = None


def load_ref_dedup():
    """The reference's compiled `set_redundant_lpj_to_low_CPU(new_states, new_lpj, old_states)`
    (numpy int8/uint8 views + float32/float64 lpj, in place), or None if not built."""
    global _dedup_mod
    if _dedup_mod is None:
        _dedup_mod = _load_ext(_EXT_NAME) or False
    return getattr(_dedup_mod, "set_redundant_lpj_to_low_CPU", None) if _dedup_mod else None


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "tvo")) and _ext_path() is not None


def import_reference():
    """Return the reference's top-level `tvo` package (build container only)."""
    if "tvo" in sys.modules and getattr(sys.modules["tvo"], "__file__", "").startswith(REFERENCE_ROOT):
        return sys.modules["tvo"]
    if not reference_available():
        raise RuntimeError("reference not available (no /root/reference or oracle/_ref not built)")
    for name in ("h5py", "munch"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                stub = types.ModuleType(name)
                if name == "munch":
                    class Munch(dict):
                        __getattr__ = dict.__getitem__
                        __setattr__ = dict.__setitem__
                    stub.Munch = Munch
                else:
                    stub.File = stub.Dataset = stub.Group = object
                sys.modules[name] = stub
    full = "tvo.variational." + _EXT_NAME
    if full not in sys.modules:
        sys.modules[full] = _load_ext(full)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import tvo  # noqa: E402
    return tvo
