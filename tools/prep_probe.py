"""Why does the R-MAT-26 preprocessing take 0.49 s in tools/prep_bench.py and 1.0-1.1 s inside bench.py's sharded block?
Runs it (GW_TIMING=1 prints the phases) in a fresh process, after a torch allocation pattern like bench.py's, and twice."""
import os
import sys
import time

os.environ["GW_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from graph_embedding_b200 import _lib

mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
if mode == "torch":
    x = torch.empty(25 << 30, dtype=torch.uint8, device="cuda")      # bench.py holds / releases tensors of this size before
    del x
    torch.cuda.empty_cache()
if mode == "held":
    x = torch.empty(25 << 30, dtype=torch.uint8, device="cuda")
for rep in range(2):
    h = _lib.GraphHandle.rmat(26, 16 << 26, seed=1)
    t0 = time.perf_counter()
    ms = h.prepare_walks()
    print(mode, "rep", rep, "device_ms %.1f wall_ms %.1f" % (ms, (time.perf_counter() - t0) * 1e3), flush=True)
    del h
