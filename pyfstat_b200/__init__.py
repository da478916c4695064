"""pyfstat_b200 -- B200-native backend for PyFstat's transient-CW F-statistic map.

One hot path, built from scratch for sm_100a: per-SFT F-stat atoms -> (t0,tau) map F_mn for
rectangular and exponential windows, its max/argmax and the lnBtSG marginalisation, behind
PyFstat's own ``tCWFstatMapVersion`` registry::

    import pyfstat, pyfstat_b200
    pyfstat_b200.register()                       # adds "b200" to tcw.fstatmap_versions
    search = pyfstat.TransientGridSearch(..., tCWFstatMapVersion="b200")

Python here is a thin ctypes layer over ``libtcw_b200.so`` (``include/tcw_b200.h``); there is
no CPU fallback.
"""

from .atoms import ATOM_DTYPE, AtomBatch, batch_from_detector_lists, from_multi_fstat_atoms, synth_atoms
from .backend import (
    BACKEND_NAME,
    b200_compute_transient_fstat_map,
    backend_available,
    fstat_map_class,
    get_handle,
    register,
    unregister,
)
from . import semicoherent
from .mcmc import TransientWalkerPool, transient_detstat_batch
from .window import (
    TRANSIENT_EXPONENTIAL,
    TRANSIENT_LAST,
    TRANSIENT_NONE,
    TRANSIENT_RECTANGULAR,
    TransientWindowRange,
    canonical_window,
)

__all__ = [
    "ATOM_DTYPE", "AtomBatch", "batch_from_detector_lists", "from_multi_fstat_atoms", "synth_atoms",
    "BACKEND_NAME", "b200_compute_transient_fstat_map", "backend_available", "fstat_map_class",
    "get_handle", "register", "unregister", "semicoherent", "TransientWalkerPool", "transient_detstat_batch",
    "TRANSIENT_NONE", "TRANSIENT_RECTANGULAR", "TRANSIENT_EXPONENTIAL", "TRANSIENT_LAST",
    "TransientWindowRange", "canonical_window",
]
