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


class TransientWalkerPool:
    """A ``pool`` for ptemcee's ``Sampler(..., pool=...)`` that turns one sampler step into ONE GPU call.

    ptemcee evaluates a step's proposals with ``list(pool.map(evaluator, thetas))`` where
    ``evaluator`` is its ``LikePriorEvaluator`` -- attributes ``logl, logp, loglargs, logpargs,
    loglkwargs, logpkwargs``; ``evaluator(theta) -> (logl, logp)``, prior first, the likelihood only
    where the prior is finite (``logl`` reported as 0 where the prior is ``-inf``).  PyFstat builds
    the sampler with ``logl=self._logl`` and no pool (``mcmc_based_searches.py:734-744``), and
    ``MCMCTransientSearch._logl`` (``:3511-3516``) is, per walker::

        in_theta = self._set_point_for_evaluation(theta)          # Doppler point + tstart / tend
        if in_theta["tend"] > self.maxStartTime: return -inf
        return search.get_det_stat(**in_theta) * likelihooddetstatmultiplier + likelihoodcoef

    i.e. one 1x1 transient map per walker.  ``map`` reproduces exactly that, but evaluates the maps
    of all walkers of the step together through :func:`transient_detstat_batch`.

    Parameters
    ----------
    mcmc_search:
        the ``MCMCTransientSearch`` (duck-typed: ``_set_point_for_evaluation(theta) -> dict`` with
        ``tstart`` / ``tend``, ``maxStartTime``, ``likelihooddetstatmultiplier``, ``likelihoodcoef``,
        ``transientWindowType``, ``BtSG``).
    atoms_for_points:
        ``callable(list of point dicts) -> AtomBatch``, one template per point: the F-stat atoms of
        each walker's Doppler point.  In PyFstat they come from ``lalpulsar.ComputeFstat``
        (``core.py:1359-1365``, CPU, outside this repository's scope).
    detstat_batch:
        the batched evaluator; defaults to :func:`transient_detstat_batch` (the CUDA path).  Only
        the tests inject another one.
    """

    def __init__(self, mcmc_search, atoms_for_points, *, device=-1, flags=None, detstat_batch=None):
        self.search = mcmc_search
        self.atoms_for_points = atoms_for_points
        self.device = device
        self.flags = flags
        self.detstat_batch = detstat_batch or transient_detstat_batch
        self.n_steps = 0
        self.n_batched = 0

    @staticmethod
    def _is_evaluator(fn) -> bool:
        return all(hasattr(fn, a) for a in ("logl", "logp", "loglargs", "logpargs"))

    def map(self, fn, thetas):
        thetas = [np.asarray(t) for t in thetas]
        if not self._is_evaluator(fn):
            # not a likelihood/prior evaluator: behave like any pool, call by call
            return [fn(t) for t in thetas]
        s = self.search
        lkw, pkw = getattr(fn, "loglkwargs", {}) or {}, getattr(fn, "logpkwargs", {}) or {}
        lp = np.array([fn.logp(t, *fn.logpargs, **pkw) for t in thetas], dtype=float)
        if np.isnan(lp).any():
            raise ValueError("Prior function returned NaN.")
        ll = np.zeros(len(thetas))  # ptemcee reports logl = 0 where the prior is -inf
        live = np.flatnonzero(lp != -np.inf)
        if getattr(fn.logl, "__func__", fn.logl) is not getattr(s._logl, "__func__", s._logl):
            # some other likelihood: nothing to batch, evaluate it as the evaluator would
            for i in live:
                ll[i] = fn.logl(thetas[i], *fn.loglargs, **lkw)
        elif len(live):
            points = [s._set_point_for_evaluation(thetas[i]) for i in live]
            tstart = np.array([p["tstart"] for p in points], dtype=float)
            tend = np.array([p["tend"] for p in points], dtype=float)
            ok = tend <= s.maxStartTime  # mcmc_based_searches.py:3513-3514
            ll[live[~ok]] = -np.inf
            if ok.any():
                batch = self.atoms_for_points([p for p, k in zip(points, ok) if k])
                if batch.T != int(ok.sum()):
                    raise ValueError("atoms_for_points returned the wrong number of templates")
                detstat, _ = self.detstat_batch(batch, tstart[ok], tend[ok], s.transientWindowType or "rect",
                                                BtSG=bool(getattr(s, "BtSG", False)), device=self.device,
                                                flags=self.flags)
                ll[live[ok]] = detstat * s.likelihooddetstatmultiplier + s.likelihoodcoef
                self.n_batched += int(ok.sum())
        if np.isnan(ll).any():
            raise ValueError("Log likelihood function returned NaN.")
        self.n_steps += 1
        return [(float(a), float(b)) for a, b in zip(ll, lp)]

    # the rest of the multiprocessing.Pool surface samplers may touch
    def close(self):
        pass

    def join(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False
