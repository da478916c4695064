"""Exponential-window kernel timing probe (development aid, run under gpurun).
$PYFSTAT_B200_LIB selects the library build."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pyfstat_b200 import _lib as L  # noqa: E402
from pyfstat_b200.atoms import synth_atoms  # noqa: E402
from pyfstat_b200.window import canonical_window  # noqa: E402

h = L.Handle(0)
for n, T in ((1440, 128), (5760, 4)):
    b = synth_atoms(T, n, ("H1", "L1"), seed=3)
    w = canonical_window("exp", 10**9, n)
    h.upload(b)
    ms = []
    for i in range(9):
        h.flush_l2()
        h.synchronize()
        h.timer_start()
        h.map_resident(w, L.WANT_BTSG)
        t = h.timer_stop()
        if i >= 3:
            ms.append(h.last_stage_ms()["map"])
    print(f"lib={os.path.basename(L.LIB_PATH)} N={n} T={T}: map {statistics.mean(ms):.3f} ms (min {min(ms):.3f})")
h.close()
