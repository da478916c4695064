"""Single-template call latency through the registered plugin callable (what PyFstat's
dispatcher invokes once per Doppler point), configs[0] and configs[1] shapes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pyfstat_b200
from pyfstat_b200.atoms import synth_atoms
from pyfstat_b200.window import canonical_window

for win, n, dets in (("rect", 1440, ("H1",)), ("exp", 1440, ("H1", "L1")), ("rect", 2880, ("H1", "L1"))):
    b = synth_atoms(1, n, dets, seed=1)
    w = canonical_window(win, 10**9, n)
    for btsg in (False, True):
        ts = []
        for i in range(30):
            t0 = time.perf_counter()
            fm = pyfstat_b200.b200_compute_transient_fstat_map(b, w, btsg)
            _ = fm.maxF, fm.get_maxF_idx()
            ts.append(time.perf_counter() - t0)
        t_fmn = time.perf_counter(); F = fm.F_mn; t_fmn = time.perf_counter() - t_fmn
        print(f"{win} N={n} {'+'.join(dets)} BtSG={btsg}: median {1e3*np.median(ts[5:]):.3f} ms/call, "
              f"min {1e3*min(ts[5:]):.3f} ms; lazy F_mn materialisation {1e3*t_fmn:.2f} ms ({F.nbytes/1e6:.1f} MB)")
