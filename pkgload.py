"""Imports the package directory `poly-commit_b200/` (the hyphen is the name the task fixes) under the
importable module name `poly_commit_b200`."""
import importlib.util
import os
import sys

_NAME = "poly_commit_b200"


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "poly-commit_b200")
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
