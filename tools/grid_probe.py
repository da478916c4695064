"""BASELINE configs[2] shape through the batched grid-search driver (development aid, run under
gpurun): 60 d H1+L1 rect window with lnBtSG, N grid points, atoms pre-staged in pinned memory."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pyfstat_b200 import _lib as L  # noqa: E402
from pyfstat_b200.atoms import AtomBatch, synth_atoms  # noqa: E402
from pyfstat_b200.grid_search import BatchedTransientGridSearch  # noqa: E402
from pyfstat_b200.window import canonical_window  # noqa: E402

n, B = 2880, 256
n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 4
win = sys.argv[2] if len(sys.argv) > 2 else "rect"
b0 = synth_atoms(B, n, ("H1", "L1"), seed=3)
pin = L.PinnedBuffer(b0.atoms.nbytes)
arr = np.frombuffer(pin.array, dtype=b0.atoms.dtype, count=b0.atoms.size).reshape(b0.atoms.shape)
arr[...] = b0.atoms
batch = AtomBatch(arr, b0.n_atoms, b0.TAtom)  # the same pinned atoms stand in for every batch of grid points
ranges = {"F0": [30.0, 30.0 + 1e-6 * (B * n_batches - 1), 1e-6], "F1": [-1e-10], "F2": [0], "Alpha": [1.0], "Delta": [0.5]}
w = canonical_window(win, 10**9, n)
for rep in range(2):
    s = BatchedTransientGridSearch(lambda pts: batch[: len(pts)], ranges, w, BtSG=True, batch_size=B)
    t0 = time.perf_counter()
    s.run()
    dt = time.perf_counter() - t0
print(f"{win} 60 d H1+L1, {s.total_iterations} grid points in batches of {B}: {dt * 1e3:.1f} ms wall "
      f"({s.total_iterations / dt:.0f} templates/s, {s.total_iterations * (n - 1) * (n + 1) / dt:.3e} cells/s; "
      f"map calls {s.timingFstatMap * 1e3:.1f} ms)")
