"""particle_fm_b200 -- B200-native (sm_100a) implementation of particle_fm's hot path:
ODE sampling and flow-matching training through the EPiC vector-field network.

Python here is host code only; the arithmetic runs in libpfm_b200.so (hand-written CUDA) behind the
C ABI of include/pfm_b200.h.  There is no CPU fallback."""
__version__ = "0.1.0"
