"""hgb200 -- host-side binding of libhgb200.so (hand-written sm_100a kernels, C ABI in include/hg_api.h).

PyTorch is used for device memory, streams and torch.distributed only.  There is NO CPU fallback:
every op raises if the shared library or a CUDA device is missing.
"""
from ._lib import lib, HgError, lib_path  # noqa: F401
from . import ops  # noqa: F401
