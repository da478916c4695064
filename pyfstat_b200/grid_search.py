"""Batched counterpart of ``pyfstat.TransientGridSearch``.

The reference loops serially over its Cartesian grid, one ``get_det_stat`` and one map per
point (``pyfstat/grid_based_searches.py:1144-1201``); the registered single-template signature
cannot batch.  This driver keeps the reference's grid construction, output columns, detection
statistic choice and file format, but evaluates the transient maps ``batch_size`` templates at
a time through ``tcw_map_batch`` (and shards the grid over ranks when a ``torch.distributed``
process group is initialised).

What it does NOT do: produce the atoms.  In PyFstat they come from
``lalpulsar.ComputeFstat`` per Doppler point (``core.py:1359-1365``, CPU, out of this repo's
scope); here the caller supplies ``atoms_for_points(points) -> AtomBatch`` -- e.g. a loop over
``ComputeFstat.get_fullycoherent_twoF`` + an atoms copy in a real search, or synthetic atoms.
"""

from __future__ import annotations

import itertools
import logging
import time

import numpy as np

from . import _lib
from .batch import gather_records, map_again, shard_range, submit_batch, wait_batch
from .window import TRANSIENT_NONE, TransientWindowRange

logger = logging.getLogger(__name__)

FMT_DETSTAT = "%.9g"   # pyfstat/core.py:123-124
FMT_DOPPLER = "%.16g"  # pyfstat/core.py:126-127


def get_array_from_tuple(x):
    """Grid points of one search dimension, as ``GridSearch._get_array_from_tuple``
    (grid_based_searches.py:173-189): ``[value]`` or ``[min, max, step]`` (end point included)."""
    x = np.atleast_1d(x)
    if len(x) == 1:
        return np.array(x, dtype=float)
    if len(x) == 3:
        return np.linspace(x[0], x[1], num=int((x[1] - x[0]) / x[2]) + 1, endpoint=True)
    return np.array(x, dtype=float)


class BatchedTransientGridSearch:
    """Transient-CW grid search with the (t0,tau) maps evaluated in batches on the GPU.

    Parameters
    ----------
    atoms_for_points:
        ``callable(points) -> AtomBatch``; ``points`` is a structured array (fields =
        ``search_keys``) of the grid points of one batch, in grid order.
    search_ranges:
        ``{key: [value] | [min, max, step]}`` for every key in ``search_keys`` (the reference's
        ``F0s, F1s, F2s, Alphas, Deltas``).
    window:
        the transient window range (``lalpulsar.transientWindowRange_t`` duck type), i.e. what
        ``ComputeFstat`` builds from ``transientWindowType, t0Band, tauBand, dt0, dtau, tauMin``
        (``core.py:821-891``).
    BtSG:
        if true the detection statistic is ``lnBtSG`` and ``t0_MP, tau_MP`` are output as well
        (grid_based_searches.py:999-1002, 1040-1042).
    """

    search_keys = ["F0", "F1", "F2", "Alpha", "Delta"]  # grid_based_searches.py:35-36

    def __init__(self, atoms_for_points, search_ranges, window, BtSG=False, batch_size=256, device=-1,
                 search_keys=None, header=None):
        if search_keys is not None:
            self.search_keys = list(search_keys)
        self.atoms_for_points = atoms_for_points
        self.window = TransientWindowRange.from_any(window)
        self.window.check_type()
        self.BtSG = bool(BtSG)
        self.detstat = "lnBtSG" if self.BtSG else "maxTwoF"
        self.batch_size = int(batch_size)
        self.device = device
        self.output_file_header = list(header or [])
        self.coord_arrays = [get_array_from_tuple(search_ranges[k]) for k in self.search_keys]
        self.total_iterations = int(np.prod([len(c) for c in self.coord_arrays]))
        in_dtype = np.dtype({"names": self.search_keys, "formats": [float] * len(self.search_keys)})
        self.input_data = np.array(list(itertools.product(*self.coord_arrays)), dtype=in_dtype)
        # output columns in the reference's order (grid_based_searches.py:1023-1042)
        self.output_keys = self.search_keys + ["twoF", "maxTwoF"]
        if self.detstat != "maxTwoF":
            self.output_keys.append(self.detstat)
        self.output_keys += ["t0_ML", "tau_ML"]
        if self.BtSG:
            self.output_keys += ["t0_MP", "tau_MP"]
        self.data = None
        self.timingFstatMap = 0.0

    def _records_for_range(self, lo, hi):
        """Map records + full-coherent F for grid points [lo, hi), batch by batch."""
        recs, twoF = [], []
        starts = list(range(lo, hi, self.batch_size))

        def produce(a):
            b = min(a + self.batch_size, hi)
            batch = self.atoms_for_points(self.input_data[a:b])
            if batch.T != b - a:
                raise ValueError("atoms_for_points returned the wrong number of templates")
            return batch

        nxt = produce(starts[0]) if starts else None
        for k, a in enumerate(starts):
            batch = nxt
            t0 = time.time()
            ticket = submit_batch(batch, self.window, BtSG=self.BtSG, device=self.device)
            t_sub = time.time() - t0
            # the GPU works on this batch while the host produces the atoms of the next one
            # (in a real search: lalpulsar.ComputeFstat per Doppler point, core.py:1359-1365)
            nxt = produce(starts[k + 1]) if k + 1 < len(starts) else None
            t0 = time.time()
            r = wait_batch(ticket)
            # twoF over ALL the data = the 1x1 map of TRANSIENT_NONE (tcw:742-749); the atoms of the
            # batch are still resident on the device after the first call: no second upload
            full = map_again(TransientWindowRange(type=TRANSIENT_NONE), batch, device=self.device)
            self.timingFstatMap += t_sub + time.time() - t0
            recs.append(r)
            twoF.append(2.0 * full["maxF"].astype(np.float64))
        if not recs:
            return np.zeros(0, dtype=_lib.RESULT_DTYPE), np.zeros(0)
        return np.concatenate(recs), np.concatenate(twoF)

    def run(self, group=None):
        """Evaluate the grid.  With an initialised process group every rank computes a contiguous
        block of grid points and all ranks end up with the full table."""
        import torch.distributed as dist

        T = self.total_iterations
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        else:
            rank, world = 0, 1
        lo, hi = shard_range(T, rank, world)
        rec, twoF = self._records_for_range(lo, hi)
        ext = np.zeros(len(rec), dtype=np.dtype(_lib.RESULT_DTYPE.descr + [("twoF", "<f8")]))
        for name in _lib.RESULT_DTYPE.names:
            ext[name] = rec[name]
        ext["twoF"] = twoF
        ext = gather_records(ext, T, group)

        out_dtype = np.dtype({"names": self.output_keys, "formats": [float] * len(self.output_keys)})
        data = np.zeros(T, dtype=out_dtype)
        for k in self.search_keys:
            data[k] = self.input_data[k]
        data["twoF"] = ext["twoF"]
        data["maxTwoF"] = 2.0 * ext["maxF"].astype(np.float64)  # core.py:1460
        if self.BtSG:
            data["lnBtSG"] = ext["lnBtSG"]
            data["t0_MP"] = ext["t0_MP"]
            data["tau_MP"] = ext["tau_MP"]
        data["t0_ML"] = ext["t0_ML"]   # = windowRange.t0 + m*dt0 (grid_based_searches.py:1129)
        data["tau_ML"] = ext["tau_ML"]
        self.data = data
        self.records = ext
        logger.info("Total time spent computing transient F-stat maps: %.2f s", self.timingFstatMap)
        return data

    # ---- output, in the reference's text format (grid_based_searches.py:471-501, 1236-1252) ----
    def _get_savetxt_fmt_list(self):
        fmt = {k: FMT_DOPPLER for k in self.search_keys}
        fmt.update({"twoF": FMT_DETSTAT, "maxTwoF": FMT_DETSTAT, "lnBtSG": FMT_DETSTAT})
        fmt.update({"t0_ML": "%d", "tau_ML": "%d", "t0_MP": "%d", "tau_MP": "%d"})
        return [fmt[k] for k in self.output_keys]

    def save_array_to_disk(self, out_file):
        header = "\n".join(self.output_file_header + [" ".join(self.output_keys)])
        np.savetxt(out_file, np.nan_to_num(self.data), delimiter=" ", header=header,
                   fmt=self._get_savetxt_fmt_list())

    def get_max_det_stat(self):
        """Grid point with the loudest detection statistic (GridSearch.get_max_det_stat)."""
        idx = int(np.argmax(self.data[self.detstat]))
        return {k: self.data[k][idx] for k in self.output_keys}
