"""Drop-in `src` package: the Python API of ChampiB/CBO_with_OOP (CBO, DoCalculus, GaussianProcessFactory, Monitor,
DataLoader, ArgumentParser, the graph classes and the helpers in src.utils_functions) re-implemented from scratch
on top of the B200 acquisition sweep in cbo_with_oop_b200 -- no GPy / emukit / paramz.  `runCBO.py` of the reference
imports exactly these names."""
