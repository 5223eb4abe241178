"""Import shim: exposes the package directory ``video-diffusion-speedrun_b200/`` (not a valid Python
identifier) as the importable package ``vds_b200``.  ``import vds_b200`` needs only the repo root on
``sys.path``."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "video-diffusion-speedrun_b200")


def _load():
    spec = importlib.util.spec_from_file_location(
        "vds_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR]
    )
    mod = importlib.util.module_from_spec(spec)
    sys.modules["vds_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


_load()
