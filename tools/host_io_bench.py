"""Host-only: native BAM window loop (inflate + decode) and gathers on the C2 fixture, per thread count."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from npore_b200 import bamio
S, NP = bench.load_tables()
ref, reads = bench.make_workload(20260101, 1_000_000, 3000, 10000, NP)
bench.write_fixture_bam("/tmp/c2.bam", ref, reads)
print("threads:", os.cpu_count())
for nt in (0, 4, 8, 16, 32):
    for win in (16 << 20, 0):
        best = 1e9
        for rep in range(4):
            t = time.perf_counter()
            nb = bamio.NativeBam("/tmp/c2.bam", nt, window_bytes=win)
            tot = nb.n
            if win:
                while True:
                    n = nb.advance()
                    if not n:
                        break
                    tot += n
            dt = time.perf_counter() - t
            nb.close()
            best = min(best, dt)
        print(f"n_threads {nt:2d} window {win >> 20:2d} MB: all windows loaded in {1e3 * best:6.1f} ms ({tot} records)", flush=True)
nb = bamio.NativeBam("/tmp/c2.bam", 0, window_bytes=0)
sel = np.arange(nb.n)
for name, fn in (("gather text", lambda: nb.gather(sel, 0, want_codes=False, want_cigar=False)), ("gather nib", lambda: nb.gather_nib(sel, 0)),
                 ("gather cigar", lambda: nb.gather_cigar(sel, 0))):
    best = 1e9
    for rep in range(4):
        t = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t)
    print(f"{name}: {1e3 * best:.1f} ms for 3000 reads")
