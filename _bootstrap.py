"""Import helper: the package directory is named `biogpt.cpp_b200` (not a legal dotted
module name), so it is loaded under the alias `biogpt_cpp_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "biogpt.cpp_b200")
ALIAS = "biogpt_cpp_b200"


def load_pkg():
    if ALIAS in sys.modules:
        return sys.modules[ALIAS]
    spec = importlib.util.spec_from_file_location(
        ALIAS, os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[ALIAS] = mod
    spec.loader.exec_module(mod)
    return mod
