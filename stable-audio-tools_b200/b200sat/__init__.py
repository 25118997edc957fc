"""b200sat — B200-native (sm_100a) implementation of the Stable Audio latent-diffusion hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); every hot op is a hand-written CUDA kernel in
libb200sat.so reached through the C ABI of include/b200sat.h.
"""
from ._lib import lib, B200SatError, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
