"""Import shim: the product package lives in the directory ``isca-2025-lia_b200/`` (the
name the build contract asks for), which is not a valid Python identifier.  Importing
``lia_b200`` loads that directory as the package ``lia_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "isca-2025-lia_b200")
_spec = importlib.util.spec_from_file_location(
    "lia_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["lia_b200"] = _mod
_spec.loader.exec_module(_mod)
