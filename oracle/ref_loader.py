"""Imports the unmodified reference modules from `oracle/_ref/*.refbin` (built by oracle/build_ref.py).

Test / measurement infrastructure only (tests/, __graft_entry__.smoke(), bench.py's CPU legs)."""
import contextlib
import importlib.machinery
import importlib.util
import io
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_PREFIX = "sct_reference_"


def available() -> bool:
    return all(os.path.exists(os.path.join(_DIR, n + ".refbin")) for n in ("model", "train", "data_augmentation"))


def _load(name):
    """The reference modules import each other by bare name (`from data_augmentation import ...`), so they are
    registered under their own names for the duration of the import, then kept under a private prefix."""
    key = _PREFIX + name
    if key in sys.modules:
        return sys.modules[key]
    path = os.path.join(_DIR, name + ".refbin")
    loader = importlib.machinery.SourcelessFileLoader(name, path)
    spec = importlib.util.spec_from_loader(name, loader)
    mod = importlib.util.module_from_spec(spec)
    prev = sys.modules.get(name)
    sys.modules[name] = mod
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            loader.exec_module(mod)
    finally:
        if prev is not None:
            sys.modules[name] = prev
        else:
            sys.modules.pop(name, None)
    sys.modules[key] = mod
    return mod


def load():
    """Returns (model_module, train_module) of the unmodified reference."""
    if not available():
        raise RuntimeError("oracle/_ref is not built (python oracle/build_ref.py, needs /root/reference)")
    da = _load("data_augmentation")
    sys.modules.setdefault("data_augmentation", da)  # train.py: `from data_augmentation import SmartContractAugmenter`
    return _load("model"), _load("train")
