"""Import alias: `import cse_b200` loads the package in `contextual-speech-extraction_b200/`
(the directory name mirrors the reference repo and is not a legal Python identifier)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "contextual-speech-extraction_b200")
_spec = importlib.util.spec_from_file_location(
    "cse_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cse_b200"] = _mod
_spec.loader.exec_module(_mod)
