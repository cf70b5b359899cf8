"""pop2-cesm_b200: B200-native (CUDA sm_100a, fp64) implementation of the POP2 baroclinic +
barotropic time-step hot path behind a C ABI (include/pop_b200.h).

Python here is only the host-side mirror used by tests and bench.py: `config` (the pop_config
struct), `synthetic` (seeded inputs), `api` (ctypes binding of csrc/libpop_b200.so) and `build`.
The compute path is the CUDA library; importing `api` fails loudly when it is missing.
"""
from . import config, synthetic  # noqa: F401

__all__ = ["config", "synthetic", "api", "build"]


def __getattr__(name):
    if name in ("api", "build"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
