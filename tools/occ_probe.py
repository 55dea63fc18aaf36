import sys, os
sys.path.insert(0, os.getcwd())
os.environ["NPORE_DEBUG"] = "1"
import numpy as np, bench
from npore_b200.engine import Realigner
S, NP = bench.load_tables()
for r in (10, 30, 60, 100):
    e = Realigner(S, NP, r=r)
    out = e.align_many([np.array([1,2,3,4,1,2,3,4],np.uint8)], [np.array([1,2,3,4,1,2,3,4],np.uint8)], ["8="])
    print(r, e.stats()["fwd_warps_per_sm"]); e.close()
