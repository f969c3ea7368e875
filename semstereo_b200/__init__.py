"""semstereo_b200 — B200-native (sm_100a) disparity hot path of SemStereo behind a C-ABI.

    ops         validated wrappers over the C-ABI kernels (include/semstereo_b200.h)
    hotpath     DisparityHotPath: the fused path of SemStereo.forward:273-324, state_dict-compatible
    submodule   signed operator surface (names of the reference's models/submodule.py)
    submodule_  unsigned operator surface (names of models/submodule_.py)
    params      parameter inventory + seeded synthetic inputs
    dist        one-process-per-GPU batch sharding and the NCCL output gather

Importing the package does not load CUDA; the first kernel call builds/loads libsemstereo_b200.so and raises
if that is impossible (there is no CPU or torch fallback).
"""
__version__ = "0.1.0"
