"""Loader for the package directory ``vsr-tlaplus_b200/`` (its hyphenated name is fixed by the build
contract and is not importable by name): registers it as module ``vsr_tlaplus_b200``."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "vsr-tlaplus_b200")


def load():
    name = "vsr_tlaplus_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
