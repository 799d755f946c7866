"""Walker preprocessing (k_cc_small + k_cc_block) as a function of the CTAs of k_cc_small in the launch per SM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_embedding_b200 import _lib
for scale in (22, 26):
    for ctas in ("16", "8", "32", "64", "128", "256"):
        os.environ["GW_CC_CTAS_PER_SM"] = ctas
        best = 1e9
        for rep in range(2):
            h = _lib.GraphHandle.rmat(scale, 16 << scale, a=0.45, b=0.15, c=0.15, seed=1)
            best = min(best, h.prepare_walks())
            del h
        print("R-MAT-%d, %s CTAs per SM: %.2f ms" % (scale, ctas, best), flush=True)
