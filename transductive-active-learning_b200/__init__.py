"""B200-native discrete-domain GP hot path.

This directory is meant to be put on `sys.path`: it provides

  lib/     the drop-in mirror of the reference's `lib.model`, `lib.algorithms`,
           `lib.gp`, `lib.noise`, `lib.utils` call surface for the hot path
  talgp/   the device engine and the ctypes binding of include/tal_b200.h
  csrc/    the sm_100a CUDA kernels and the C ABI (built into libtal_b200.so)
"""
