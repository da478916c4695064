"""Geometry of lalpulsar's ``XLALFastNegExp`` lookup table, MEASURED instead of recalled.

The reference's ``lal`` backend (``pyfstat/tcw_fstat_map_funcs.py:571-586``) evaluates every
exponential -- the exponential-window weights inside ``XLALComputeTransientFstatMap`` and the
terms of ``XLALComputeTransientBstat`` / the posteriors -- through a nearest-point lookup table
``XLALFastNegExp`` (``lalpulsar/lib/TransientCW_utils.c``, a third-party source that is not in
the PyFstat tree).  A wrong table step moves exponential-window ``F_mn`` by ~1e-3 relative, ten
times the parity bar, and two recollections of its constants are on file (``xmax = 20`` with
5120 steps, SURVEY A.4-1, and with 2000 steps).  So nothing is compiled in: the table geometry is
a run-time property of the device handle (``tcw_set_exp_lut``), and wherever lalpulsar is
importable -- i.e. in every real PyFstat installation, since the atoms come from it -- this
module measures the table from the library itself:

* ``xmax``: the largest argument with a non-zero result (bisection);
* the step: the positions of the first two jumps of the step function (``dx/2`` and ``3dx/2``);
* the entries: ``f(i * dx)`` for every ``i`` -- uploaded as they are, so even a table built by a
  different libm is reproduced bit for bit.

``measure_exp_lut`` works on any callable, which is how the CPU tests exercise it (against the
oracle's restatement with both geometries).
"""

from __future__ import annotations

import logging
import math

import numpy as np

logger = logging.getLogger(__name__)


def _bisect_last_true(pred, lo: float, hi: float, iters: int = 200) -> float:
    """Largest double x in [lo, hi) with pred(x), given pred(lo) and not pred(hi)."""
    for _ in range(iters):
        mid = 0.5 * (lo + hi)
        if mid == lo or mid == hi:
            break
        if pred(mid):
            lo = mid
        else:
            hi = mid
    return lo


def measure_exp_lut(f, max_length: int = 1 << 22, tol: float = 0.0):
    """``(xmax, length, table)`` of a nearest-point negative-exponential table ``f(x)``.

    ``tol``: absolute tolerance below which two values of ``f`` count as equal (0 for a direct
    view of the table; ~1e-12 when ``f`` is recovered through another computation).
    Raises ``ValueError`` if ``f`` does not behave like a nearest-point table (no cut-off, no
    steps, inconsistent step positions) -- the caller then keeps the configured default.
    """

    def same(a, b):
        return abs(a - b) <= tol

    f0 = float(f(0.0))
    if not same(f0, 1.0):
        raise ValueError(f"f(0) = {f0!r}, expected 1")
    # cut-off: f(x) == 0 for x > xmax
    hi = 1.0
    while not same(float(f(hi)), 0.0):
        hi *= 2.0
        if hi > 1e6:
            raise ValueError("no cut-off found: f(x) > 0 up to 1e6")
    xmax = _bisect_last_true(lambda x: not same(float(f(x)), 0.0), 0.0, hi)
    if abs(xmax - round(xmax)) < 1e-9 * max(1.0, xmax):
        xmax = float(round(xmax))
    # first two jumps of the step function: x = dx/2 and 3 dx/2
    j1 = _bisect_last_true(lambda x: same(float(f(x)), f0), 0.0, xmax)
    f1 = float(f(math.nextafter(j1, math.inf)))
    if same(f1, f0) or same(f1, 0.0):
        raise ValueError("no first step found")
    j2 = _bisect_last_true(lambda x: same(float(f(x)), f1), math.nextafter(j1, math.inf), xmax)
    dx = j2 - j1
    if not (dx > 0) or abs(2.0 * j1 - dx) > 1e-6 * dx:
        raise ValueError(f"step positions {j1!r}, {j2!r} are not dx/2, 3dx/2: not a nearest-point table")
    length = int(round(xmax / dx))
    if length < 1 or length > max_length:
        raise ValueError(f"implausible table length {length}")
    dx = xmax / length
    table = np.array([float(f(min(i * dx, xmax))) for i in range(length + 1)], dtype=np.float64)
    # consistency: nearest-point lookup over the whole range
    for i in (1, 2, length // 3, length // 2, length - 1):
        for off in (-0.49, 0.0, 0.49):
            x = (i + off) * dx
            if 0.0 <= x <= xmax and not same(float(f(x)), table[i]):
                raise ValueError(f"f({x!r}) != table[{i}]: not a nearest-point table with dx = {dx!r}")
    return xmax, length, table


def neg_exp_through_bstat(lalpulsar):
    """``f(x) = XLALFastNegExp(x)`` recovered WITHOUT a direct binding: lalpulsar's own
    ``ComputeTransientBstat`` (the call of the reference's ``lal`` path, tcw:578-580) on a 1 x 2 map
    ``F_mn = [[0, -x]]`` returns ``ln(70/2) + ln(1 + FastNegExp(x))``.  Values are good to ~1e-15
    absolute, enough to find the cut-off and the step positions; the table entries themselves are
    then taken as ``exp(-i dx)`` (canonical)."""
    fm = lalpulsar.CreateTransientFstatMap(1, 2)
    wr = lalpulsar.transientWindowRange_t()
    wr.type = lalpulsar.TRANSIENT_RECTANGULAR
    wr.t0, wr.t0Band, wr.dt0 = 0, 0, 1
    wr.tau, wr.tauBand, wr.dtau = 1, 1, 1

    def f(x):
        fm.F_mn.data[0, 0] = 0.0
        fm.F_mn.data[0, 1] = -float(x)
        fm.maxF = 0.0
        return math.expm1(float(lalpulsar.ComputeTransientBstat(wr, fm)) - math.log(35.0))

    return f


def probe_lalpulsar(module=None):
    """``(xmax, length, table)`` measured from lalpulsar: through ``lalpulsar.FastNegExp`` (the SWIG
    name of ``XLALFastNegExp``) when it is exported, else through ``ComputeTransientBstat``
    (:func:`neg_exp_through_bstat`; entries then canonical, ``table`` is ``None``).  ``None`` when
    lalpulsar is not importable or neither route works."""
    lalpulsar = module
    if lalpulsar is None:
        try:
            import lalpulsar  # noqa: PLC0415
        except Exception:  # noqa: BLE001 -- absent in the build container
            return None
    f = getattr(lalpulsar, "FastNegExp", None)
    try:
        if f is not None:
            return measure_exp_lut(f)
        xmax, length, _ = measure_exp_lut(neg_exp_through_bstat(lalpulsar), tol=1e-12)
        return xmax, length, None
    except Exception as e:  # noqa: BLE001
        logger.warning("could not measure lalpulsar's FastNegExp table (%s): keeping the configured geometry", e)
        return None


def parse_geometry(text: str):
    """``"xmax:length"`` -> ``(float, int)`` (``$PYFSTAT_B200_EXPLUT`` / ``$TCW_EXP_LUT``)."""
    a, b = text.split(":")
    xmax, length = float(a), int(b)
    if not (xmax > 0 and length >= 1):
        raise ValueError(f"bad exp-table geometry {text!r}")
    return xmax, length
