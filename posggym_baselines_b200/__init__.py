"""posg-b200: the batched posggym grid-world model hot path of posggym-baselines
(model.step / sample_initial_state / observation / reward, POMCP leaf rollouts,
belief reinvigoration, vec-env collection) as hand-written sm_100a CUDA kernels behind
the reference's Python API.  PyTorch is used for device memory, streams and
torch.distributed only; all arithmetic is in libposg_b200.so (C-ABI: include/posg_b200.h).

There is NO CPU fallback: calling a compute entry point without the built library or
without a CUDA device raises ``PosgB200Error``.
"""
from posggym_baselines_b200._lib import PosgB200Error, lib, library_path  # noqa: F401

__version__ = "0.1.0"
