"""B200-native explicit 3-D linear-elastic FE time stepping (synchronization-avoiding).

Host side (Python) of the hot path of desResLab/Synchronization-avoiding-algorithms:
mesh / partition maps / sparse assembly feed an opaque device plan (`plan.StepPlan`) whose
time-step kernels are hand-written sm_100a CUDA behind the C ABI of `include/saa_fem.h`.
Import as `saa_b200` (see saa_b200.py at the repository root).
"""
__version__ = "0.1.0"
