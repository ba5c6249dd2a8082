"""reflexiv_b200 -- host-side mirror of Reflexiv's counter / run path on top of libreflexiv_cuda.

The compute lives in ``reflexiv_b200/csrc`` (hand-written sm_100a CUDA behind the C ABI of
``include/reflexiv_cuda.h``).  This package only binds that ABI (ctypes), mirrors the reference's
parameter block and pipeline entry points, and provides the torch.distributed plumbing for sharded runs.
There is no CPU implementation here: every entry point fails if the CUDA library or a GPU is missing.
"""
from ._lib import RfxError, load_library, lib_path  # noqa: F401
from .params import DefaultParam, Parameter, ParameterOfCounter  # noqa: F401
from .pipeline import ReflexivContext, Pipelines  # noqa: F401

__all__ = ["RfxError", "load_library", "lib_path", "DefaultParam", "Parameter", "ParameterOfCounter",
           "ReflexivContext", "Pipelines"]
