"""Development probe (run under gpurun): the exponential-window recurrence / tensor-core path against the
oracle and the tiled direct sum, plus timings."""
import os
import statistics
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import tcw_oracle as O  # noqa: E402
from pyfstat_b200 import _lib as L  # noqa: E402
from pyfstat_b200.atoms import synth_atoms  # noqa: E402
from pyfstat_b200.window import canonical_window  # noqa: E402

h = L.Handle(0)
sizes = [int(a) for a in sys.argv[1:]] or [144, 480]
for n in sizes:
    for dets in (("H1", "L1"), ("H1",)):
        b = synth_atoms(2, n, dets, seed=11)
        w = canonical_window("exp", 10**9, n)
        for exact in (True, False):
            fl = L.WANT_FMN | L.WANT_BTSG | (L.EXP_EXACT if exact else 0)
            res, F = h.map_batch(b, w, fl)
            resd, Fd = h.map_batch(b, w, fl | L.EXP_DIRECT)
            t0 = time.time()
            o = O.compute_map(b.template(1), b.TAtom, w, exact_exp=exact)
            Fo = o["F_mn"]
            ok = Fo != 2.0
            rel = np.abs(F[1] - Fo) / np.abs(Fo)
            reld = np.abs(Fd[1] - Fo) / np.abs(Fo)
            print(f"n={n} dets={len(dets)} exact={exact}: path {int(res['path'][1])}/{int(resd['path'][1])} "
                  f"rec-vs-oracle max {rel[ok].max():.3e} p99.9 {np.quantile(rel[ok], 0.999):.3e} | direct-vs-oracle max {reld[ok].max():.3e} "
                  f"| fallback cells rec {int((F[1] == 2.0).sum())} oracle {int((Fo == 2.0).sum())} "
                  f"| argmax {(int(res['m_ML'][1]), int(res['n_ML'][1]))} vs {(o['m_ML'], o['n_ML'])} "
                  f"| lnBtSG {float(res['lnBtSG'][1]):.6f} vs {o['lnBtSG']:.6f} (oracle {time.time() - t0:.1f} s)", flush=True)
# timings
for n, T in ((1440, 128), (5760, 8)):
    b = synth_atoms(T, n, ("H1", "L1"), seed=3)
    w = canonical_window("exp", 10**9, n)
    h.upload(b)
    for name, fl in (("lut rec+tc", 0), ("exact rec", L.EXP_EXACT), ("lut direct", L.EXP_DIRECT)):
        ms = []
        for i in range(6):
            h.flush_l2()
            h.synchronize()
            h.timer_start()
            h.map_resident(w, L.WANT_BTSG | fl)
            h.timer_stop()
            if i >= 2:
                ms.append(h.last_stage_ms()["map"])
        print(f"N={n} T={T} {name}: map {statistics.mean(ms):.3f} ms (min {min(ms):.3f}) = {statistics.mean(ms) / T:.4f} ms/template", flush=True)
h.close()
