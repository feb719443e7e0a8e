"""mcarray_b200: B200 (sm_100a) implementation of mcarray's frame-based multichannel hot path.

The product is libmcarray_b200.so (C ABI: include/mcarray_b200.h) plus the C++ host classes in include/mcarray/.
This Python package is the thin ctypes mirror used by the tests and bench.py."""
from . import _capi as capi  # noqa: F401
from .processors import (DelayAndSumFan, FastBinauralMasking, FilterAndSumFan, FreqGCCBinauralLocalisation, MultibandBinarualLocalisation, Processor,  # noqa: F401
                         SourceLocalisation, SourceSeparationAndLocalisation, SrpPhat, TdoaEstimator, set_default_device)
from . import sharding  # noqa: F401,E402
