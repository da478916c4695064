"""Transient window ranges: lalpulsar's ``transientWindowRange_t`` as plain Python.

lalpulsar cannot be imported in the build container, so the enum values the reference uses
(``pyfstat/tcw_fstat_map_funcs.py:691-697, 742-743, 793-807``; ``pyfstat/core.py:828-840``)
are carried here as constants, and window ranges are accepted duck-typed (anything with
``.type, .t0, .t0Band, .dt0, .tau, .tauBand, .dtau``).
"""

from __future__ import annotations

from dataclasses import dataclass

TRANSIENT_NONE = 0
TRANSIENT_RECTANGULAR = 1
TRANSIENT_EXPONENTIAL = 2
TRANSIENT_LAST = 3
TRANSIENT_EXP_EFOLDING = 3  # pyCUDAkernels/cudaTransientFstatExpWindow.cu:22

WINDOW_TYPES = {"none": TRANSIENT_NONE, "rect": TRANSIENT_RECTANGULAR, "exp": TRANSIENT_EXPONENTIAL}

_U32 = 0xFFFFFFFF


@dataclass
class TransientWindowRange:
    """Stand-in for ``lalpulsar.transientWindowRange_t`` (all fields UINT4)."""

    type: int = TRANSIENT_NONE
    t0: int = 0
    t0Band: int = 0
    dt0: int = 0
    tau: int = 0
    tauBand: int = 0
    dtau: int = 0

    @classmethod
    def from_any(cls, w) -> "TransientWindowRange":
        """Copy of a duck-typed window range; never aliases or mutates the caller's object
        (the reference's pycuda path mutates it for TRANSIENT_NONE, tcw:742-749)."""
        vals = {}
        for f in ("type", "t0", "t0Band", "dt0", "tau", "tauBand", "dtau"):
            v = int(getattr(w, f))
            if v < 0 or v > _U32:
                raise ValueError(f"windowRange.{f}={v} does not fit UINT4")
            vals[f] = v
        return cls(**vals)

    def check_type(self):
        """ValueError as tcw:691-697."""
        if self.type >= TRANSIENT_LAST:
            raise ValueError(
                "Unknown window-type ({}) passed as input. Allowed are [0,{}].".format(
                    self.type, TRANSIENT_LAST - 1
                )
            )

    def dims(self):
        """``(N_t0Range, N_tauRange)`` (tcw:775-780); (1, 1) for TRANSIENT_NONE."""
        self.check_type()
        if self.type == TRANSIENT_NONE:
            return 1, 1
        if self.dt0 <= 0 or self.dtau <= 0:
            raise ValueError("windowRange.dt0 and .dtau must be positive")
        return self.t0Band // self.dt0 + 1, self.tauBand // self.dtau + 1


def canonical_window(window: str, t0_data: int, n_atoms: int, TAtom: int = 1800) -> TransientWindowRange:
    """The window recipe of the reference's examples and tests:
    ``t0Band = Tspan - 2*TAtom``, ``tau_min = 2*TAtom``, ``tauBand = Tspan``,
    ``dt0 = dtau = TAtom``  =>  ``N_t0 = N-1``, ``N_tau = N+1``
    (examples/transient_examples/PyFstat_example_short_transient_grid_search.py:81-83,
    tests/test_grid_based_searches.py:328-329, pyfstat/core.py:882)."""
    Tspan = n_atoms * TAtom
    return TransientWindowRange(
        type=WINDOW_TYPES[window],
        t0=t0_data,
        t0Band=Tspan - 2 * TAtom,
        dt0=TAtom,
        tau=2 * TAtom,
        tauBand=Tspan,
        dtau=TAtom,
    )
