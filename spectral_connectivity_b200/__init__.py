"""B200-native multitaper spectral-connectivity engine.

Drop-in for the hot path of Eden-Kramer-Lab/spectral_connectivity: ``Multitaper`` and
``Connectivity`` keep the reference's API; the compute is hand-written sm_100a CUDA behind
the C ABI in ``include/sc_b200.h`` (``libsc_b200.so``).  There is no CPU fallback.
"""
from .connectivity import Connectivity, pinned_empty, unpack_upper  # noqa: F401
from .minimum_phase_decomposition import minimum_phase_decomposition  # noqa: F401
from .transforms import Multitaper  # noqa: F401
from ._dpss import dpss_windows  # noqa: F401
from .wrapper import multitaper_connectivity  # noqa: F401

__version__ = "0.1.0"
__all__ = ["Multitaper", "Connectivity", "minimum_phase_decomposition", "dpss_windows", "pinned_empty", "unpack_upper", "multitaper_connectivity"]
