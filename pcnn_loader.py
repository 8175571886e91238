"""Import helper: the package directory is `parallel-cnn_b200/` (hyphen, per the build spec), which Python cannot
import by name; load() registers it as the module `parallel_cnn_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "parallel-cnn_b200")
MODULE = "parallel_cnn_b200"


def load():
    if MODULE in sys.modules:
        return sys.modules[MODULE]
    spec = importlib.util.spec_from_file_location(MODULE, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[MODULE] = mod
    spec.loader.exec_module(mod)
    return mod
