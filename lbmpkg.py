"""lbmpkg.py -- import shim: the product package lives in `cfd-codes_b200/` (hyphenated, so not a
Python identifier).  `from lbmpkg import lbm` gives that package as module `cfd_codes_b200`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG_DIR = os.path.join(_ROOT, "cfd-codes_b200")
_NAME = "cfd_codes_b200"

if _NAME in sys.modules:
    lbm = sys.modules[_NAME]
else:
    _spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    lbm = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = lbm
    _spec.loader.exec_module(lbm)
