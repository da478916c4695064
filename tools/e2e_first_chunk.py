"""End-to-end time of one tcw_map_batch over host atoms (exp window) for first-upload-chunk sizes:
   python tools/e2e_first_chunk.py N T  (set TCW_UPLOAD_FIRST to override the first chunk)"""
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyfstat_b200 import _lib as L  # noqa: E402
from pyfstat_b200.atoms import synth_atoms  # noqa: E402
from pyfstat_b200.window import canonical_window  # noqa: E402

N, T = int(sys.argv[1]), int(sys.argv[2])
h = L.Handle(0)
b = synth_atoms(T, N, ("H1", "L1"), seed=3, pinned_alloc=L.pinned_atoms_alloc())
w = canonical_window("exp", 10**9, N)
ts = []
for i in range(13):
    t0 = time.perf_counter()
    h.map_batch(b, w, L.WANT_BTSG)
    if i >= 3:
        ts.append(time.perf_counter() - t0)
print("TCW_UPLOAD_FIRST=%s  N=%d T=%d  e2e median %.3f ms  min %.3f ms" % (os.environ.get("TCW_UPLOAD_FIRST", "default"), N, T,
      1e3 * statistics.median(ts), 1e3 * min(ts)))
h.close()
