"""Error statistics of the tiled kernels against the CPU oracle (lal semantics) at BASELINE sizes
(development aid, run under gpurun).  Prints one JSON line per case."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import tcw_oracle as O  # noqa: E402
from pyfstat_b200 import _lib as L  # noqa: E402
from pyfstat_b200.atoms import synth_atoms  # noqa: E402
from pyfstat_b200.window import canonical_window  # noqa: E402

h = L.Handle(0)
for win, n, dets in (("rect", 1440, ("H1", "L1")), ("rect", 2880, ("H1", "L1")), ("rect", 1440, ("H1",)),
                     ("exp", 1440, ("H1", "L1")), ("exp", 720, ("H1",))):
    b = synth_atoms(1, n, dets, seed=171)
    w = canonical_window(win, 10**9, n)
    res, F = h.map_batch(b, w, L.WANT_FMN | L.WANT_BTSG)
    t0 = time.time()
    o = O.compute_map(b.template(0), b.TAtom, w)
    dt = time.time() - t0
    Fo = o["F_mn"]
    rel = np.abs(F[0] - Fo) / np.maximum(np.abs(Fo), 1e-30)
    print(json.dumps(dict(
        window=win, atoms=n, detectors="+".join(dets), cells=int(rel.size), oracle_s=round(dt, 2),
        rel_median=float(np.median(rel)), rel_p99=float(np.quantile(rel, 0.99)), rel_p9999=float(np.quantile(rel, 0.9999)),
        rel_max=float(rel.max()), n_gt_1e4=int((rel > 1e-4).sum()), n_gt_1e5=int((rel > 1e-5).sum()),
        argmax_equal=bool((int(res["m_ML"][0]), int(res["n_ML"][0])) == (o["m_ML"], o["n_ML"])),
        maxF_rel=float(abs(float(res["maxF"][0]) - o["maxF"]) / o["maxF"]),
        lnBtSG_abs=float(abs(float(res["lnBtSG"][0]) - o["lnBtSG"])),
    )))
h.close()
