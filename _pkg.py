"""Registers the hyphenated package directory ``dl-poly_b200/`` as the importable module ``dl_poly_b200``."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "dl-poly_b200")


def load():
    if "dl_poly_b200" in sys.modules:
        return sys.modules["dl_poly_b200"]
    spec = importlib.util.spec_from_file_location(
        "dl_poly_b200", os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["dl_poly_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
