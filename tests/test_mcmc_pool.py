"""SURVEY 8(f)-1, second half: the ptemcee ``pool``-style mapper for ``MCMCTransientSearch``.

ptemcee is not installed here, so the sampler side is a stub that does what
``ptemcee.Sampler._evaluate`` does with a pool: ``list(pool.map(LikePriorEvaluator(...), thetas))``
with the evaluator semantics of ptemcee (prior first; ``logl = 0`` where the prior is ``-inf``).
The search side is a stand-in for ``MCMCTransientSearch`` carrying exactly the members its ``_logl``
uses (``pyfstat/mcmc_based_searches.py:3479-3516``).  The batched pool must return what the serial
evaluator returns, walker by walker.
"""

import numpy as np
import pytest

from pyfstat_b200 import _lib as L
from pyfstat_b200.atoms import AtomBatch, synth_atoms
from pyfstat_b200.mcmc import TransientWalkerPool, transient_detstat_batch
from pyfstat_b200.window import TransientWindowRange

T0, TATOM, N = 10**9, 1800, 96


class LikePriorEvaluator:
    """ptemcee.sampler.LikePriorEvaluator, restated."""

    def __init__(self, logl, logp, loglargs=(), logpargs=(), loglkwargs=None, logpkwargs=None):
        self.logl, self.logp = logl, logp
        self.loglargs, self.logpargs = loglargs, logpargs
        self.loglkwargs, self.logpkwargs = loglkwargs or {}, logpkwargs or {}

    def __call__(self, x):
        lp = self.logp(x, *self.logpargs, **self.logpkwargs)
        if np.isnan(lp):
            raise ValueError("Prior function returned NaN.")
        if lp == float("-inf"):
            ll = 0
        else:
            ll = self.logl(x, *self.loglargs, **self.loglkwargs)
        return ll, lp


class StubSampler:
    """The part of ptemcee.Sampler that touches the pool."""

    def __init__(self, logl, logp, loglargs, logpargs, pool=None):
        self._likeprior = LikePriorEvaluator(logl, logp, loglargs, logpargs)
        self.pool = pool

    def evaluate(self, ps):
        mapf = map if self.pool is None else self.pool.map
        results = list(mapf(self._likeprior, ps.reshape((-1, ps.shape[-1]))))
        logl = np.fromiter((r[0] for r in results), float, count=len(results)).reshape(ps.shape[:-1])
        logp = np.fromiter((r[1] for r in results), float, count=len(results)).reshape(ps.shape[:-1])
        return logl, logp


class StubTransientSearch:
    """Members of MCMCTransientSearch that _logl / _logp / the pool use.  theta = (F0, tstart, duration)."""

    theta_keys = ["F0", "transient_tstart", "transient_duration"]
    transientWindowType = "rect"
    BtSG = False
    likelihooddetstatmultiplier = 0.5
    likelihoodcoef = np.log(70.0 / 10.0**4)

    def __init__(self, single_detstat, maxStartTime):
        self.single_detstat = single_detstat  # (point dict) -> detection statistic of ONE walker
        self.maxStartTime = maxStartTime
        self.search = object()

    def _set_point_for_evaluation(self, theta):  # mcmc_based_searches.py:3479-3509
        p = {"F0": theta[0], "F1": -1e-10, "F2": 0.0, "Alpha": 1.0, "Delta": 0.5}
        p["tstart"] = theta[1]
        p["tend"] = theta[1] + theta[2]
        return p

    def _logp(self, theta, *args):
        return 0.0 if 29.0 <= theta[0] <= 31.0 and theta[2] > 0 else -np.inf

    def _logl(self, theta, search):  # mcmc_based_searches.py:3511-3516
        in_theta = self._set_point_for_evaluation(theta)
        if in_theta["tend"] > self.maxStartTime:
            return -np.inf
        return self.single_detstat(in_theta) * self.likelihooddetstatmultiplier + self.likelihoodcoef


def atoms_of_point(p) -> AtomBatch:
    """Synthetic atoms keyed by the Doppler point (stands in for lalpulsar.ComputeFstat)."""
    return synth_atoms(1, N, ("H1", "L1"), seed=int(round((p["F0"] - 29.0) * 1000)), t0_data=T0, TAtom=TATOM)


def atoms_for_points(points) -> AtomBatch:
    parts = [atoms_of_point(p) for p in points]
    return AtomBatch(np.concatenate([b.atoms for b in parts]), np.concatenate([b.n_atoms for b in parts]), TATOM)


def walkers(ntemps=2, nwalkers=12, seed=3):
    rng = np.random.default_rng(seed)
    ps = np.empty((ntemps, nwalkers, 3))
    ps[..., 0] = rng.uniform(28.8, 31.2, (ntemps, nwalkers))            # some outside the prior
    ps[..., 1] = T0 + rng.uniform(0, 0.5 * N * TATOM, (ntemps, nwalkers))
    ps[..., 2] = rng.uniform(3 * TATOM, 0.7 * N * TATOM, (ntemps, nwalkers))  # some beyond maxStartTime
    return ps


def _check(pool_factory, single_detstat):
    search = StubTransientSearch(single_detstat, maxStartTime=T0 + N * TATOM)
    serial = StubSampler(search._logl, search._logp, (search.search,), (None,), pool=None)
    pool = pool_factory(search)
    batched = StubSampler(search._logl, search._logp, (search.search,), (None,), pool=pool)
    ps = walkers()
    ll0, lp0 = serial.evaluate(ps)
    ll1, lp1 = batched.evaluate(ps)
    assert np.array_equal(lp0, lp1)
    assert np.isneginf(lp0).any() and np.isneginf(ll0).any(), "the test must cover both -inf branches"
    assert np.array_equal(np.isneginf(ll0), np.isneginf(ll1))
    fin = np.isfinite(ll0)
    assert np.array_equal(ll0[fin], ll1[fin]), "batched step != walker-by-walker evaluation"
    assert pool.n_steps == 1 and pool.n_batched == int((np.isfinite(ll0) & ~np.isneginf(lp0)).sum())
    # a pool is also handed plain functions: behaves like map
    assert pool.map(lambda t: float(t[0]) * 2, ps.reshape(-1, 3)[:3]) == [float(t[0]) * 2 for t in ps.reshape(-1, 3)[:3]]
    # a NaN prior raises like ptemcee's evaluator
    bad = LikePriorEvaluator(search._logl, lambda t, *a: np.nan, (search.search,), (None,))
    with pytest.raises(ValueError):
        pool.map(bad, ps.reshape(-1, 3)[:2])
    with pool as p:
        assert p is pool
    pool.close()
    pool.join()


def test_pool_batches_a_sampler_step_host_logic(oracle):
    """CPU: the batched evaluator is replaced by the oracle (host logic only)."""

    def one_cell(batch, t, tstart, tend):
        w = TransientWindowRange(1, int(tstart), 0, TATOM, int(tend - tstart), 0, TATOM)
        return 2.0 * float(np.float32(oracle.compute_map(batch.template(t), TATOM, w, allow_degenerate=True)["maxF"]))

    def oracle_detstat_batch(batch, tstarts, tends, wtype, BtSG=False, device=-1, flags=None):
        assert wtype == "rect" and not BtSG
        return np.array([one_cell(batch, t, tstarts[t], tends[t]) for t in range(batch.T)]), None

    def single(p):
        return one_cell(atoms_of_point(p), 0, p["tstart"], p["tend"])

    _check(lambda s: TransientWalkerPool(s, atoms_for_points, detstat_batch=oracle_detstat_batch), single)


@pytest.mark.gpu
def test_pool_batches_a_sampler_step_on_gpu(gpu, monkeypatch):
    """GPU: one tcw_map_batch_windows call per step == one registered-callable call per walker."""
    from pyfstat_b200 import backend

    monkeypatch.setattr(backend, "get_handle", lambda device=-1: gpu)
    from pyfstat_b200 import mcmc

    monkeypatch.setattr(mcmc, "get_handle", lambda device=-1: gpu)

    def single(p):
        w = TransientWindowRange(1, int(p["tstart"]), 0, TATOM, int(p["tend"] - p["tstart"]), 0, TATOM)
        fm = backend.b200_compute_transient_fstat_map(atoms_of_point(p), w, False,
                                                      flags=L.ALLOW_DEGENERATE | L.FORCE_GENERIC)
        return 2.0 * fm.maxF

    launches0 = gpu.launch_count
    _check(lambda s: TransientWalkerPool(s, atoms_for_points), single)
    assert gpu.launch_count > launches0
    # and the helper itself on a 1x1-per-walker step, BtSG flavour: lnBtSG = ln 70 + F
    b = atoms_for_points([{"F0": 29.5 + 0.01 * k} for k in range(8)])
    ts = T0 + TATOM * np.arange(8.0)
    det, rec = transient_detstat_batch(b, ts, ts + 20 * TATOM, "exp", BtSG=True)
    assert np.allclose(det, np.log(70.0) + rec["maxF"].astype(np.float64), atol=1e-9)
