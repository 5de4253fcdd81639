"""Imports the package directory `vbmatrixfactorization.jl_b200/` under the module name `vbmf_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "vbmatrixfactorization.jl_b200")


def load():
    if "vbmf_b200" in sys.modules:
        return sys.modules["vbmf_b200"]
    spec = importlib.util.spec_from_file_location("vbmf_b200", os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["vbmf_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def build(force=False):
    spec = importlib.util.spec_from_file_location("vbmf_b200_build", os.path.join(PKG_DIR, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force)
