"""Run one resident workload a few times (target for ncu captures under gpurun).

usage: python tools/prof_one.py {rect|exp} [N] [T] [flags: fmn,btsg] [reps]
"""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pyfstat_b200 import _lib as L  # noqa: E402
from pyfstat_b200.atoms import synth_atoms  # noqa: E402
from pyfstat_b200.window import canonical_window  # noqa: E402


def main():
    win = sys.argv[1] if len(sys.argv) > 1 else "rect"
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 2880
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    fl = sys.argv[4] if len(sys.argv) > 4 else ""
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
    flags = (L.WANT_FMN if "fmn" in fl else 0) | (L.WANT_BTSG if "btsg" in fl else 0)
    h = L.Handle(0)
    b = synth_atoms(T, N, ("H1", "L1"), seed=3)
    w = canonical_window(win, 10**9, N)
    h.upload(b)
    for _ in range(reps):
        h.timer_start()
        h.map_resident(w, flags)
        ms = h.timer_stop()
    cells = (N - 1) * (N + 1) * T
    print(f"{win} N={N} T={T} flags={fl!r}: {ms:.3f} ms {cells / ms * 1e3:.3e} cells/s {h.last_stage_ms()}")
    h.close()


if __name__ == "__main__":
    main()
