"""Per-step batch evaluation of the transient detection statistic for MCMC (BASELINE config 5).

``MCMCTransientSearch`` evaluates, per walker and per step, ONE cell: ``ComputeFstat`` is built
with a window type but no bands (``mcmc_based_searches.py:3443-3466``), and each ``_logl`` call
(``:3511-3516``) sets ``windowRange.t0 = int(tstart)``, ``windowRange.tau = int(tend - tstart)``
(``core.py:1447-1449``) => a 1x1 map whose single F is returned as ``maxTwoF`` (or as
``lnBtSG = ln 70 + F``).  A sampler step therefore is a batch of templates, each with its own
window -- which is what ``tcw_map_batch_windows`` computes in one call.

As for the grid driver, the atoms of each walker's Doppler point come from the caller.
"""

from __future__ import annotations

import numpy as np

from . import _lib
from .backend import default_flags, get_handle
from .window import WINDOW_TYPES, TransientWindowRange


def transient_detstat_batch(batch, tstarts, tends, transientWindowType="rect", BtSG=False, maxStartTime=None,
                            TAtom_step=None, device=-1, flags=None):
    """Detection statistic of every walker of one sampler step.

    Mirrors ``MCMCTransientSearch._logl`` + ``ComputeFstat.get_transient_detstats``:
    ``-inf`` where ``tend > maxStartTime`` (mcmc_based_searches.py:3513-3514), otherwise
    ``2 * F(t0=int(tstart), tau=int(tend - tstart))`` -- or ``lnBtSG`` with ``BtSG=True``.

    Parameters: ``batch`` -- :class:`AtomBatch`, one template per walker; ``tstarts``, ``tends`` --
    arrays of GPS seconds.  Returns ``(detstat, records)``.
    """
    tstarts = np.asarray(tstarts, dtype=float)
    tends = np.asarray(tends, dtype=float)
    if len(tstarts) != batch.T or len(tends) != batch.T:
        raise ValueError("need one (tstart, tend) per template")
    wtype = WINDOW_TYPES[transientWindowType]
    step = int(TAtom_step or batch.TAtom)  # dt0 = dtau = Tsft by default (core.py:843-844); irrelevant for 1x1
    ok = np.ones(batch.T, dtype=bool) if maxStartTime is None else tends <= maxStartTime
    detstat = np.full(batch.T, -np.inf)
    if not ok.any():
        return detstat, np.zeros(0, dtype=_lib.RESULT_DTYPE)
    idx = np.flatnonzero(ok)
    sub = batch if ok.all() else type(batch)(batch.atoms[idx], batch.n_atoms[idx], batch.TAtom)
    # one transientWindowRange_t row per walker: t0 = int(tstart), tau = int(tend - tstart) (core.py:1447-1449)
    t0 = tstarts[idx].astype(np.int64)
    tau = (tends[idx] - tstarts[idx]).astype(np.int64)
    if t0.min() < 0 or tau.min() < 0 or t0.max() > 0xFFFFFFFF or tau.max() > 0xFFFFFFFF:
        raise ValueError("window start times / durations do not fit UINT4")
    wins = np.zeros((len(idx), 7), dtype=np.uint32)
    wins[:, 0], wins[:, 1], wins[:, 3], wins[:, 4], wins[:, 6] = wtype, t0, step, tau, step
    if flags is None:
        flags = default_flags()
    flags |= _lib.WANT_BTSG if BtSG else 0
    # a walker can land on a single-atom window; pycuda semantics (fallback) rather than abort
    rec, _ = get_handle(device).map_batch_windows(sub, wins, flags | _lib.ALLOW_DEGENERATE)
    detstat[idx] = rec["lnBtSG"] if BtSG else 2.0 * rec["maxF"].astype(np.float64)
    return detstat, rec
