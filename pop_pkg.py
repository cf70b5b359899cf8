"""Loader for the product package, whose directory name (`pop2-cesm_b200/`) is not a valid Python
identifier: it is imported as module `pop2_cesm_b200`."""
import importlib.util
import os
import sys

_NAME = "pop2_cesm_b200"
ROOT = os.path.dirname(os.path.abspath(__file__))


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    d = os.path.join(ROOT, "pop2-cesm_b200")
    spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    m = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = m
    spec.loader.exec_module(m)
    return m
