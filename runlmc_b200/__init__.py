"""runlmc_b200: B200-native (sm_100a, fp64) SKI-LMC covariance products and
probe-MINRES gradients behind the Python API of vlad17/runlmc.

Sub-packages mirror the reference layout for the hot path only:
  linalg/  approx/  lmc/  util/  kern/   (+ fused.py: the device operator)
All arithmetic runs in runlmc_b200/_lib/liblmc_b200.so (CUDA); there is no CPU
fallback."""
__version__ = '0.1.0'
