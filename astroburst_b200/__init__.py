"""astroburst_b200 -- Blackwell (sm_100a) backend for the AstroBurst per-pixel hot path.

The product is `lib/libastroburst_cuda.so` (hand-written CUDA behind the C ABI declared in
include/astroburst_cuda.h).  This package is the thin host-side mirror of the reference's
`core::*` functions used by the tests and the benchmark: same names, argument meaning and
error behaviour as the Rust functions they replace.  There is no CPU fallback: importing
works anywhere (so the build can be checked without a GPU) but every compute call needs the
CUDA library and a device, and fails loudly otherwise.
"""
from ._lib import AstroBurstError, Context, default_context, lib, library_path  # noqa: F401
from .core import (  # noqa: F401
    StackConfig,
    StackResult,
    sigma_clip_combine,
    stack_fits_units,
    stack_images,
    stack_images_device,
)

from . import alignment, calibration, fits, imaging, preview  # noqa: F401,E402

__version__ = "0.1.0"
