"""Importable alias for the package directory ``anime-illust-image-searcher_b200/``
(its name is not a Python identifier).  ``import ais_b200`` gives that package."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "anime-illust-image-searcher_b200")
_spec = _ilu.spec_from_file_location(
    "ais_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["ais_b200"] = _mod
_spec.loader.exec_module(_mod)
