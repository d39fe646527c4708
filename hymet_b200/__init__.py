"""hymet_b200 -- B200-native `mash screen` for HYMET's candidate-selection stage.

Only what the path needs: csrc/ (hand-written sm_100a kernels + the C ABI of
include/hymet_screen.h), a ctypes binding, the host mirror of the mash CLI, the .msh
writer and the synthetic workload generators.  Importing the package does not import
torch; `hymet_b200.dist` (multi-GPU glue) and bench.py do.
"""
__version__ = "0.1.0"
