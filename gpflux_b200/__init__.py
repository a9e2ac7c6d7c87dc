"""gpflux_b200: GPflux's per-layer sparse-variational GP hot path on B200 (sm_100a).

Hand-written CUDA behind a C ABI (``include/gpfx.h`` -> ``gpflux_b200/lib/libgpflux_b200.so``),
with a host-side mirror of the reference's ``gpflux.layers.GPLayer`` / ``gpflux.models.DeepGP``
interface.  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import architectures, gpflow_compat, helpers, layers, models, ops  # noqa: F401
from .exceptions import GPLayerIncompatibilityException  # noqa: F401
