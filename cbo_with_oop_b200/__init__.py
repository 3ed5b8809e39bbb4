"""cbo_with_oop_b200 -- B200-native (sm_100a) acquisition sweep for Causal Bayesian Optimisation.

Only what the hot path needs lives here: `csrc/` (hand-written CUDA kernels + the C ABI of
include/cbo_b200.h), the ctypes binding (`_lib`), the device-memory owner / driver (`engine`), the multi-GPU
partitioner (`partition`) and the observational-GP state builder (`obs_gp`).  The reference-shaped Python
API (CBO, DoCalculus, GaussianProcessFactory, ...) is the top-level `src` package, which calls into this one.
"""
from .partition import SetSize, partition  # noqa: F401

__all__ = ["SetSize", "partition", "SetProblem", "SweepEngine", "SweepOutput"]


def __getattr__(name):
    # engine imports torch and the CUDA library lazily so that `import cbo_with_oop_b200` stays cheap
    if name in ("SetProblem", "SweepEngine", "SweepOutput"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
