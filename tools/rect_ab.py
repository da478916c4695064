"""Rect-window kernel timing probe (development aid, run under gpurun): BASELINE shapes, the three
output modes, stage times from CUDA events.  Env: TCW_RECT_PERSIST=0/1 selects the kernel."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pyfstat_b200 import _lib as L  # noqa: E402
from pyfstat_b200.atoms import synth_atoms  # noqa: E402
from pyfstat_b200.window import canonical_window  # noqa: E402

h = L.Handle(0)
for n, dets, T in ((2880, ("H1", "L1"), 64), (1440, ("H1",), 256), (5760, ("H1", "L1"), 16)):
    b = synth_atoms(T, n, dets, seed=3)
    w = canonical_window("rect", 10**9, n)
    h.upload(b)
    for name, flags in (("fmn", L.WANT_FMN), ("btsg", L.WANT_BTSG), ("max", 0)):
        ms, st = [], []
        for i in range(13):
            h.flush_l2()
            h.synchronize()
            h.timer_start()
            h.map_resident(w, flags)
            t = h.timer_stop()
            if i >= 3:
                ms.append(t)
                st.append(h.last_stage_ms())
        m = {k: round(statistics.mean(s[k] for s in st), 4) for k in st[0]}
        cells = T * (n - 1) * (n + 1)
        print(f"persist={os.environ.get('TCW_RECT_PERSIST', '1')} N={n} {'+'.join(dets)} T={T} {name}: step {statistics.mean(ms):.4f} ms "
              f"{cells / statistics.mean(ms) * 1e3:.3e} cells/s stages {m}")
h.close()
