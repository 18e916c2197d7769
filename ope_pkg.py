"""Import helper: the package directory is named `object-pose-estimation_b200` (as the layout contract asks),
which is not a valid Python identifier, so it is registered under the module name `ope_b200`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(_ROOT, "object-pose-estimation_b200")


def load():
    if "ope_b200" in sys.modules:
        return sys.modules["ope_b200"]
    spec = importlib.util.spec_from_file_location(
        "ope_b200", os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ope_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
