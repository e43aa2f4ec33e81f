"""Import alias: the product package lives in `monolithicfemvlfs.jl_b200/` (a directory name
Python cannot import directly); `import vlfs_b200` loads it as a regular package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "monolithicfemvlfs.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "vlfs_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["vlfs_b200"] = _mod
_spec.loader.exec_module(_mod)
