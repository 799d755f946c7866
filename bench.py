#!/usr/bin/env python
"""bench.py — headline benchmark of the walk engine (driver contract in the task statement).

Default (N=1): BASELINE.json configs[2] — node2vec on synthetic R-MAT scale-22 (reference generator
probabilities .45/.15/.15/.25, 16*2^22 edge tuples), p=0.25 q=4, walk_length=80.  One "step" is one
walk iteration: one walk from every non-isolated vertex (= one pass of simulate_walks' inner loop,
node2vec.py:53-57); num_walks=10 is `--steps 10`.

  value  walk-steps/s, corpus produced into a resident HBM buffer from resident start nodes
  e2e    same metric through the host-buffer C-ABI call (gw_node2vec_walks): pinned host start
         nodes -> H2D -> kernel -> D2H of the whole corpus, all inside the timed region
  roofline  algorithmic bytes (68 + 32*S(d_prev) per step, measured on the corpus
         actually produced) / CUDA-event kernel time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle's python port of the reference walker, 1 core, bounded sample

`--workload simrank` measures TopSim queries/s on a Barabasi-Albert graph (configs[4] shape).
`--impl reference` times the CPU port on all host cores (rank 0 only); its line names the graph it really walked.
The `sharded` block (always at N > 1, `--sharded on` at N = 1) runs BASELINE configs[3] and configs[4] through the
C-ABI's own multi-GPU entry points (gw_comm_init + gw_node2vec_walks_sharded / gw_simrank_topk_sharded): one pass of
the R-MAT scale-26 start list split over the ranks, and 1 M BA-10M queries with the NCCL gather of the top-k tiles.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="node2vec", choices=["node2vec", "simrank"])
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--edge-factor", type=int, default=16)
    ap.add_argument("--rmat-abc", default="0.45,0.15,0.15", help="R-MAT quadrant probabilities a,b,c (d = 1-a-b-c); "
                    "default = the reference generator's (RMATGraphGenerator.java:179-182); Graph500: 0.57,0.19,0.19")
    ap.add_argument("--p", type=float, default=0.25)
    ap.add_argument("--q", type=float, default=4.0)
    ap.add_argument("--walk-length", type=int, default=80)
    ap.add_argument("--ba-nodes", type=int, default=10_000_000)
    ap.add_argument("--ba-m", type=int, default=8)
    ap.add_argument("--queries-per-step", type=int, default=8192)
    ap.add_argument("--sample", type=int, default=10000)
    ap.add_argument("--sr-step", type=int, default=5)
    ap.add_argument("--topk", type=int, default=20)
    ap.add_argument("--estimator", default="mc", choices=["mc", "hybrid"],
                    help="simrank workload: mc = SingleRandomWalk (headline), hybrid = TopSim_singleSample path tree")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the TopSim block that rides along the node2vec line")
    ap.add_argument("--sharded", default="auto", choices=["auto", "on", "off"],
                    help="BASELINE configs 4/5 through gw_*_sharded (auto: on for N > 1 and for the default N = 1 run)")
    ap.add_argument("--shard-scale", type=int, default=26, help="R-MAT scale of the sharded node2vec pass (configs[3]: 26)")
    ap.add_argument("--shard-queries", type=int, default=1_000_000, help="queries of the sharded TopSim run (configs[4]: 1 M)")
    ap.add_argument("--no-e2e-variants", action="store_true", help="skip the pageable / forced hand-off e2e variants")
    ap.add_argument("--embeddings", default="auto", choices=["auto", "on", "off"],
                    help="the device consumer of the corpus (SURVEY 8(f)4): walks -> vocabulary -> skip-gram on the headline graph "
                         "(auto: on for the default N = 1 run)")
    ap.add_argument("--dimensions", type=int, default=128)
    ap.add_argument("--window-size", type=int, default=10)
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)                                  # the timed region is tens of ms: sample every few ms

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def ncu_traffic(kernel, workload, units=None):
    """dram__bytes_read+write per launch from the committed ncu capture (profiles/traffic.json) of exactly this kernel
    instantiation on exactly this workload; None when no capture matches (a changed kernel never inherits a stale
    figure).  Scaled to `units` per launch when the capture used another batch size."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        for e in t.get("captures", []):
            if e["kernel"] == kernel and e["workload"] == workload:
                v = float(e["dram_bytes_per_launch"])
                if units and e.get("units_per_launch"):
                    v = v * units / e["units_per_launch"]
                return v
    except Exception:
        pass
    return None


def gather_ceiling(array_bytes):
    """Measured ceiling of dependent random accesses/s for an array of this size (profiles/gather_ceiling.json,
    tools/gather_bench.cu): the bound a one-access-per-step walker actually runs against on this part."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "gather_ceiling.json")))
        xs, ys = t["array_MiB"], t["G_loads_per_s"]
        mib = array_bytes / float(1 << 20)
        if mib <= xs[0]:
            return ys[0]
        for i in range(1, len(xs)):
            if mib <= xs[i]:
                f = (np.log(mib) - np.log(xs[i - 1])) / (np.log(xs[i]) - np.log(xs[i - 1]))
                return ys[i - 1] + f * (ys[i] - ys[i - 1])
        return ys[-1]
    except Exception:
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# CPU arms (oracle port).  These are the ONLY places bench.py touches oracle/.
# ---------------------------------------------------------------------------------------------
def _rmat_edges_numpy(scale, n_tuples, a, b, c, seed):
    rs = np.random.RandomState(seed)
    frm = np.zeros(n_tuples, dtype=np.int64)
    to = np.zeros(n_tuples, dtype=np.int64)
    for _ in range(scale):
        x = rs.rand(n_tuples)
        fb = ((x >= a) & (x < a + b)) | (x >= a + b + c)
        tb = x >= a + b
        frm = (frm << 1) | fb
        to = (to << 1) | tb
    keep = frm != to
    return frm[keep], to[keep]


_REF = {}          # reference-arm state shared with forked workers: graph + alias tables (built once, inherited copy-on-write)
REF_SCALE = 10     # the port walks an R-MAT of the benchmark's generator shape at this scale (sum deg^2 must be materialised)


def _cpu_node2vec_worker(args):
    """One process: python port of simulate_walks' inner loop on its shard of start nodes, for `budget` seconds."""
    starts, seed, budget = args
    from oracle import n2v_oracle as O
    g, an, ae, L = _REF["g"], _REF["an"], _REF["ae"], _REF["L"]
    rng = np.random.RandomState(seed)
    t0 = time.perf_counter()
    steps = 0
    i = 0
    while time.perf_counter() - t0 < budget:
        s = int(starts[i % len(starts)])
        i += 1
        u = rng.rand(2 * (L - 1))
        w, _ = O.walk_replay(g, an, ae, L, s, u, 0)
        steps += len(w) - 1
    return steps, time.perf_counter() - t0


def cpu_node2vec_prepare(p, q, L):
    """Graph + preprocess_transition_probs of the port (node2vec.py:83-113), timed on one core as the reference runs it."""
    from oracle import n2v_oracle as O
    s, d = _rmat_edges_numpy(REF_SCALE, 16 << REF_SCALE, 0.45, 0.15, 0.15, 1)
    g = O.build_simple_graph(s, d, np.ones(len(s)), directed=False)
    t0 = time.perf_counter()
    an = O.alias_nodes_flat(g)
    ae = O.alias_edges_flat(g, p, q)
    prep_s = time.perf_counter() - t0
    _REF.update(g=g, an=an, ae=ae, L=L)
    deg = np.diff(g["row_ptr"])
    entries = int(deg.sum() + (deg[g["col_idx"]]).sum())          # sum deg (alias_nodes) + sum deg^2 (alias_edges)
    return {"starts": np.nonzero(deg > 0)[0], "preprocess_s": prep_s, "alias_entries": entries,
            "nodes": int((deg > 0).sum()), "directed_entries": int(deg.sum())}


def cpu_node2vec_sample(prep, seconds, procs, pool=None, seed0=1):
    """One bounded sample: every process walks for `seconds`; returns (walk-steps, wall seconds)."""
    starts = prep["starts"]
    if procs == 1 or pool is None:
        res = [_cpu_node2vec_worker((starts, seed0, seconds))]
    else:
        res = pool.map(_cpu_node2vec_worker, [(starts[i::procs], seed0 * 1000 + i + 1, seconds) for i in range(procs)])
    return sum(r[0] for r in res), max(r[1] for r in res)


def cpu_node2vec(p, q, L, seconds, procs):
    """Reference-structured walker (materialised alias tables, per-step python loop, two uniform draws per step) on a
    down-scaled R-MAT of the benchmark's generator shape: the reference cannot preprocess sum(deg^2) at scale-22, and
    its steps/s does not depend on graph size.  -> (steps/s, description, preprocess record)"""
    prep = cpu_node2vec_prepare(p, q, L)
    pool = None
    if procs > 1:
        import multiprocessing as mp
        pool = mp.get_context("fork").Pool(procs)
    try:
        steps, wall = cpu_node2vec_sample(prep, seconds, procs, pool)
    finally:
        if pool:
            pool.close()
    sample = ("python port of node2vec_walk/alias_draw (oracle/n2v_oracle.py) on R-MAT scale-%d, p=%g q=%g L=%d, "
              "%d walk-steps in %.1fs on %d process(es); preprocess_transition_probs timed apart: %d alias entries in %.2fs on 1 core"
              % (REF_SCALE, p, q, L, steps, wall, procs, prep["alias_entries"], prep["preprocess_s"]))
    return steps / wall, sample, prep


def _ba_edges_py(n, m, seed):
    rs = np.random.RandomState(seed)
    rep, src, dst = [], [], []
    for i in range(m):
        for j in range(i + 1, m):
            src.append(i); dst.append(j); rep += [i, j]
    for v in range(m, n):
        tg = set()
        while len(tg) < m:
            tg.add(rep[rs.randint(len(rep))])
        for t in tg:
            src.append(v); dst.append(t); rep += [v, t]
    return np.array(src), np.array(dst)


_BA_CACHE = {}


def cpu_simrank(sample, step, k, seconds, procs, n=100_000, m=8):
    """C restatement of SingleRandomWalk.walk + FixedMaxPQ top-k on a down-scaled BA graph; `procs` threads as
    SingleRandomWalkApproxMultiThreads.java:165-179 runs them (ctypes releases the GIL)."""
    from oracle import simrank_oracle as S
    if (n, m) not in _BA_CACHE:
        src, dst = _ba_edges_py(n, m, 1)
        _BA_CACHE[(n, m)] = S.build_multigraph(src, dst, n)
    g = _BA_CACHE[(n, m)]

    def worker(seed, out):
        st = S.java_seed(seed)
        rq = np.random.RandomState(seed)
        t0 = time.perf_counter()
        nqd = 0
        while time.perf_counter() - t0 < seconds:
            v = int(rq.randint(n))
            row, _, st = S.single_random_walk_row(g, v, sample, step, 0.6, st)
            S.fixedmaxpq_topk(row, k)
            nqd += 1
        out.append((nqd, time.perf_counter() - t0))
    outs = []
    ths = [threading.Thread(target=worker, args=(i + 1, outs)) for i in range(procs)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    nqd = sum(o[0] for o in outs)
    wall = max(o[1] for o in outs)
    desc = ("C restatement of SingleRandomWalk.walk + FixedMaxPQ (oracle/simrank_oracle.c, not a JVM) on BA n=%d m=%d, "
            "SAMPLE=%d STEP=%d k=%d, %d queries in %.1fs on %d thread(s)" % (n, m, sample, step, k, nqd, wall, procs))
    return nqd / wall, desc


def run_reference(args):
    """The reference's own CPU algorithm on the box's host cores.  Each of the W + K steps is one bounded sample (all
    cores walk for the same few seconds); the line's `steps` are the samples really taken and `config.workload` names
    the graph really walked.  The reference's modules themselves are not on the GPU box (/root/reference does not
    travel), so the arm runs the oracle port that tests/golden pins bit for bit against them: kind = "port"."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    t_all = time.perf_counter()
    K, W = max(1, args.steps), max(0, args.warmup)
    budget = float(min(args.cpu_seconds, max(1.0, 100.0 / (K + W))))       # the whole run ends within ~2 minutes
    metric, unit = metric_unit(args)
    extra = {}
    if args.workload == "node2vec":
        prep = cpu_node2vec_prepare(args.p, args.q, args.walk_length)
        import multiprocessing as mp
        pool = mp.get_context("fork").Pool(cores) if cores > 1 else None
        try:
            for i in range(W):
                cpu_node2vec_sample(prep, budget, cores, pool, seed0=i + 1)
            done, wall, step_ms = 0, 0.0, []
            for i in range(K):
                t0 = time.perf_counter()
                st, w = cpu_node2vec_sample(prep, budget, cores, pool, seed0=W + i + 1)
                done += st
                wall += w
                step_ms.append((time.perf_counter() - t0) * 1e3)
        finally:
            if pool:
                pool.close()
        v = done / wall
        workload = ("node2vec on synthetic R-MAT scale-%d (%d*2^%d tuples, a,b,c,d=.45/.15/.15/.25; %d nodes, %d directed entries), "
                    "p=%g q=%g, walk_length=%d, one step = %.1f s of walks from the shuffled node list on every host core"
                    % (REF_SCALE, 16, REF_SCALE, prep["nodes"], prep["directed_entries"], args.p, args.q, args.walk_length, budget))
        sample = ("python port of node2vec_walk/alias_draw (oracle/n2v_oracle.py; pinned bit-exact against node2vec/src/node2vec.py by "
                  "tests/golden), %d walk-steps in %.1fs on %d process(es)" % (done, wall, cores))
        extra["preprocess"] = {"what": "preprocess_transition_probs (node2vec.py:83-113) of the port on the same graph, 1 core",
                               "seconds": prep["preprocess_s"], "alias_entries": prep["alias_entries"],
                               "entries_per_s": prep["alias_entries"] / prep["preprocess_s"]}
        extra["stands_for"] = ("BASELINE configs[2] (R-MAT scale-22): the reference materialises sum(deg^2) = 1.03e10 alias entries "
                               "(123 GB) before its first walk and cannot run there; its steps/s does not depend on graph size "
                               "(BASELINE.md section 2)")
        if not args.no_secondary:
            sb = min(budget * 2, 8.0)
            v1, d1 = cpu_simrank(args.sample, args.sr_step, args.topk, sb, 1)
            vn, dn = cpu_simrank(args.sample, args.sr_step, args.topk, sb, cores)
            extra["secondary"] = {"metric": "TopSim SimRank queries/sec", "unit": "queries/s", "value": vn,
                                  "cpu_baseline": {"value": vn, "unit": "queries/s", "cores": cores, "kind": "port", "sample": dn},
                                  "one_thread": {"value": v1, "sample": d1},
                                  "config": {"workload": "TopSim SimRank top-%d on synthetic Barabasi-Albert n=100000 m=8, c=0.6 STEP=%d "
                                                         "SAMPLE=%d (stands for the BA n=1e7 shape: steps/query do not depend on n)"
                                                         % (args.topk, args.sr_step, args.sample)}}
    else:
        for i in range(W):
            cpu_simrank(args.sample, args.sr_step, args.topk, budget, cores)
        done, wall, step_ms = 0.0, 0.0, []
        sample = ""
        for i in range(K):
            t0 = time.perf_counter()
            vq, sample = cpu_simrank(args.sample, args.sr_step, args.topk, budget, cores)
            done += vq * budget
            wall += budget
            step_ms.append((time.perf_counter() - t0) * 1e3)
        v = done / wall
        workload = ("TopSim SimRank top-%d on synthetic Barabasi-Albert n=100000 m=8, c=0.6 STEP=%d SAMPLE=%d, one step = %.1f s of "
                    "queries on every host core" % (args.topk, args.sr_step, args.sample, budget))
    line = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": float(np.mean(step_ms)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "cache": "n/a (host)", "sharding": "start nodes / queries split over the host cores"},
            "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": None}
    line.update(extra)
    line["wall_s"] = time.perf_counter() - t_all
    emit(line)


def metric_unit(args):
    if args.workload == "node2vec":
        return "node2vec walk-steps/sec", "walk-steps/s"
    return "TopSim SimRank queries/sec", "queries/s"


def config_of(args):
    if args.workload == "node2vec":
        abc = "a,b,c,d=.45/.15/.15/.25" if args.rmat_abc == "0.45,0.15,0.15" else "a,b,c=" + args.rmat_abc
        return {"workload": "node2vec on synthetic R-MAT scale-%d (%d*2^%d tuples, %s), p=%g q=%g, "
                            "walk_length=%d, one step = one walk per non-isolated vertex" %
                            (args.scale, args.edge_factor, args.scale, abc, args.p, args.q, args.walk_length),
                "cache": "inputs larger than L2 (col_idx %.0f MB, corpus %.0f MB per step)" %
                         (4.0 * 2 * args.edge_factor * (1 << args.scale) / 1e6,
                          4.0 * args.walk_length * (1 << args.scale) / 1e6),
                "sharding": "graph replicated per GPU, disjoint walk ids per rank, no data-path collective"}
    return {"workload": "TopSim SimRank top-%d on synthetic Barabasi-Albert n=%d m=%d, c=0.6 STEP=%d SAMPLE=%d, "
                        "one step = %d queries%s" % (args.topk, args.ba_nodes, args.ba_m, args.sr_step, args.sample,
                                                     args.queries_per_step,
                                                     "" if getattr(args, "estimator", "mc") == "mc" else ", estimator TopSim_singleSample (path tree)"),
            "cache": "inputs larger than L2 (col_idx %.0f MB)" % (4.0 * 2 * args.ba_m * args.ba_nodes / 1e6),
            "sharding": "graph replicated per GPU, disjoint query slices per rank, no data-path collective"}


# ---------------------------------------------------------------------------------------------
_JSON_OUT = None                    # the process's real stdout once main() has pointed fd 1 at stderr


def emit(line):
    """The ONE JSON line of this run, on the real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_OUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_OUT, data)


def main():
    global _JSON_OUT
    args = parse()
    # stdout carries the JSON line and nothing else: NCCL (torch's and gw_comm's) announces its version on fd 1 whenever
    # NCCL_DEBUG is set (the image sets VERSION), native libraries may print too -- all of that goes to stderr
    sys.stdout.flush()
    _JSON_OUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
        return

    import copy
    import torch
    import torch.distributed as dist
    from graph_embedding_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    _lib.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = measure(args, rank, world, local)
    if args.workload == "node2vec" and not args.no_secondary:
        # the metric names BOTH walk-steps/s and TopSim queries/s: the second one rides along
        a2 = copy.copy(args)
        a2.workload = "simrank"
        a2.steps, a2.warmup = min(args.steps, 3), min(max(args.warmup, 1), 3)
        l2 = measure(a2, rank, world, local)
        if rank == 0:
            line["secondary"] = {k: l2[k] for k in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "config",
                                                    "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks")
                                 if k in l2}
            line["gpu_launches"] += l2["gpu_launches"]
    run_sharded = args.sharded == "on" or (args.sharded == "auto" and args.workload == "node2vec" and
                                              (world > 1 or (args.scale == 22 and not args.no_secondary)))
    if run_sharded:
        sh = measure_sharded(args, rank, world, local)
        if rank == 0:
            line["sharded"] = sh            # its own gpu_launches inside; the line's count stays that of the timed headline regions
    if args.embeddings == "on" or (args.embeddings == "auto" and args.workload == "node2vec" and world == 1 and
                                   args.scale == 22 and not args.no_secondary):
        try:                                 # an add-on block (one rank, no collective): its failure must not take the headline line with it
            line["embeddings"] = measure_embeddings(args, local)
        except Exception as ex:              # noqa: BLE001 -- reported in the line, and on stderr
            import traceback
            traceback.print_exc()
            line["embeddings"] = {"error": repr(ex)}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def measure_embeddings(args, local):
    """The consumer of the corpus on the device (node2vec/src/main.py:92-101, gensim Word2Vec sg=1 negative=5): ONE
    pass of walks from every non-isolated vertex, vocabulary scan, one epoch of skip-gram -- gw_node2vec_embeddings, the
    corpus regenerated from its seed and never moved.  Reported beside what the same pass costs when the corpus has to
    be handed to a host trainer first (the e2e of the headline)."""
    import torch
    from graph_embedding_b200 import _lib
    peak, peak_src = measured_peak()
    ra, rb, rc = [float(x) for x in args.rmat_abc.split(",")]
    g = _lib.GraphHandle.rmat(args.scale, args.edge_factor << args.scale, a=ra, b=rb, c=rc, seed=1)
    g.prepare_walks()
    starts = np.random.RandomState(7).permutation(g.nonisolated())[None, :]
    L, dim, neg = args.walk_length, args.dimensions, 5
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    vec, cnt, sec = _lib.node2vec_embeddings(g, args.p, args.q, L, 1, starts, dimensions=dim, window=args.window_size, iter=1,
                                             negative=neg, sample=1e-3, seed=11)
    wall = time.perf_counter() - t0
    words = int(cnt.sum())
    # pairs of the epoch: counted by the kernel (gw_sgns_info) -- re-run the training leg on a held model to read it
    m = _lib.SkipGram(g, dim, seed=11)
    d_starts = torch.from_numpy(starts[0]).cuda()
    d_w = torch.empty((starts.shape[1], L), dtype=torch.int32, device="cuda")
    g.walks_dev(args.p, args.q, L, d_starts.data_ptr(), starts.shape[1], d_w.data_ptr(), seed=11, walk_id_base=0)
    m.count_dev(d_w.data_ptr(), starts.shape[1], L)
    m.finalize_vocab(sample=1e-3, negative=neg)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    m.train_dev(d_w.data_ptr(), starts.shape[1], L, window=args.window_size, total_words=float(words))
    ev[1].record()
    torch.cuda.synchronize()
    train_ms = ev[0].elapsed_time(ev[1])
    pairs = m.info()["trained_pairs"]
    row = 4.0 * dim
    # per pair: syn0[word2] read + written and `negative` syn1neg rows read + written; per centre word: its own syn1neg row
    # once (it stays in registers across the window)
    alg = pairs * (2.0 * row + neg * 2.0 * row) + words * 2.0 * row
    finite = bool(np.isfinite(vec).all())
    # what the embedding learnt after this ONE pass: P(cos(u, v) of an edge > cos(u, w) of the same u with a random vertex w),
    # vectors centred on their mean, 50 k samples (0.5 = nothing).  One pass is 1 % of the reference's schedule (num_walks 10 x
    # iter 10); the uncentred global cosine is still dominated by the common early direction (profiles/r2_sg_auc_probe.txt:
    # 0.61 after one pass, 0.95 after three by this score)
    c = g.csr(weights=False, node_ids=False, first_seen=False)
    rs = np.random.RandomState(0)
    e = rs.randint(0, g.nnz, size=50000)
    eu = np.searchsorted(c["row_ptr"], e, side="right") - 1
    ecol = c["col_idx"][e]
    rw = rs.choice(starts[0], 50000)
    vc = vec - vec[starts[0]].mean(0)

    def cos(a, b):
        x, y = vc[a], vc[b]
        return (x * y).sum(1) / np.maximum(np.linalg.norm(x, axis=1) * np.linalg.norm(y, axis=1), 1e-20)
    auc = float((cos(eu, ecol) > cos(eu, rw)).mean())
    del c, m, d_w, g
    torch.cuda.empty_cache()
    return {"api": "gw_node2vec_embeddings (walks -> vocabulary scan -> skip-gram with negative sampling, corpus never leaves the device)",
            "workload": "R-MAT scale-%d, p=%g q=%g, ONE pass of walks (L=%d), dimensions=%d window=%d negative=%d sample=1e-3, one epoch"
                        % (args.scale, args.p, args.q, L, dim, args.window_size, neg),
            "words": words, "trained_pairs": int(pairs), "seconds": {"wall": wall, **{k: float(v) for k, v in sec.items()}},
            "words_per_s": words / sec["training"], "pairs_per_s": pairs / (train_ms * 1e-3),
            "roofline": {"bound": "hbm", "kernel": "k_sgns_pipe<%d>" % (dim // 32), "bytes_per_unit": alg / max(pairs, 1), "units_per_launch": int(pairs),
                         "launch_ms": train_ms, "achieved": alg / (train_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (train_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                         "traffic": ncu_traffic("k_sgns_pipe<%d>" % (dim // 32), "rmat%d/ef%d/abc=%s/p=%g/q=%g/L=%d/dim=%d/window=%d/negative=%d/sample=0.001"
                                                % (args.scale, args.edge_factor, args.rmat_abc, args.p, args.q, L, dim, args.window_size, neg), pairs),
                         "model": "rows touched per (word, word2) pair x %d B, read and written (L2 hits of hot rows not discounted)" % int(row)},
            "vectors_finite": finite, "edge_auc_after_one_pass": auc,
            "edge_auc_definition": "P(cos(u,v) of an edge > cos(u,w), w a random vertex), mean-centred vectors, 50 k samples; three passes reach 0.95 (profiles/r2_sg_auc_probe.txt)"}


def measure_sharded(args, rank, world, local):
    """BASELINE configs[3] and configs[4] through the C ABI's own multi-GPU entry points (gw_comm_init +
    gw_node2vec_walks_sharded / gw_simrank_topk_sharded): the start list / query list is SPLIT over the ranks
    (gw_shard_range), the graph is replicated, the walk corpus stays sharded (gather = 0: 21 GB per pass fit no single
    host buffer sensibly) and the top-k tiles are gathered over NCCL.  Strong scaling: the same total work at every N;
    `one_gpu_ms` is the same pass walked by ONE GPU in this very run (every rank does it, max over ranks)."""
    import torch
    import torch.distributed as dist
    from graph_embedding_b200 import _lib
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream().cuda_stream
    launches0 = _lib.kernel_launches()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # gw_comm bootstrap as any launcher does it: rank 0's 128-byte NCCL id travels over the existing process group
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid = torch.tensor(list(_lib.Comm.unique_id()), dtype=torch.uint8, device=dev)
    if world > 1:
        dist.broadcast(uid, 0)
    comm = _lib.Comm(rank, world, bytes(uid.cpu().numpy().tobytes()), local)
    out = {"api": "gw_comm_init + gw_node2vec_walks_sharded(gather=0) / gw_simrank_topk_sharded (include/graphwalk.h)",
           "n_gpus": world, "scaling": "strong"}

    # ---------------- configs[3]: node2vec, R-MAT scale-26 (default), p=4 q=0.5, ONE pass of the start list ----------------
    L, p, q = args.walk_length, 4.0, 0.5
    t0 = time.perf_counter()
    g = _lib.GraphHandle.rmat(args.shard_scale, args.edge_factor << args.shard_scale, seed=1)
    build_s = time.perf_counter() - t0
    prep_ms = g.prepare_walks()
    starts = np.random.RandomState(99).permutation(g.nonisolated())         # random.shuffle(nodes), node2vec.py:51 (same on every rank)
    n = len(starts)
    lo, hi = _lib.shard_range(n, rank, world)
    mine = hi - lo
    d_starts = torch.from_numpy(starts).to(dev)
    d_out = torch.empty((max(mine, 1), L), dtype=torch.int32, device=dev)

    def walk_slice(a, b):
        g.walks_dev(p, q, L, d_starts.data_ptr() + 8 * a, b - a, d_out.data_ptr(), seed=42, walk_id_base=a, stream=stream)

    walk_slice(lo, min(hi, lo + (1 << 20)))                                   # warm-up: Bloom filter, code, clocks
    barrier()
    reps = 3
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        walk_slice(lo, hi)
    ev[1].record()
    barrier()
    shard_ms = allmax(ev[0].elapsed_time(ev[1]) / reps)
    steps_mine, _ = g.byte_model_dev(d_out.data_ptr(), mine, L, True, stream=stream)
    steps_all = allsum(steps_mine)
    one_ms = shard_ms
    if world > 1:                                                             # the same pass on ONE GPU: slices walked back to back
        barrier()
        ev[0].record()
        for r in range(world):
            a, b = _lib.shard_range(n, r, world)
            walk_slice(a, b)
        ev[1].record()
        barrier()
        one_ms = allmax(ev[0].elapsed_time(ev[1]))
    # end to end through the sharded C-ABI call: host start list in (all of it, on every rank), this rank's rows out
    h_out = np.empty((n, L), dtype=np.int32)                                  # rows of other ranks are never touched (no pages committed)
    comm.walks(g, p, q, L, starts, seed=42, gather=0, out=h_out)              # first call: page faults of the fresh buffer, ring allocation
    barrier()
    t0 = time.perf_counter()
    comm.walks(g, p, q, L, starts, seed=42, gather=0, out=h_out)
    e2e_s = allmax(time.perf_counter() - t0)
    ho = g.last_handoff()
    k = min(mine, 4096)
    if k:                                                                     # the rows that arrived == the rows a device-resident run produces
        walk_slice(lo, lo + k)
        torch.cuda.synchronize()
    same = bool((torch.from_numpy(h_out[lo:lo + k]).to(dev) == d_out[:k]).all().item()) if k else True
    out["node2vec_rmat%d" % args.shard_scale] = {
        "workload": "node2vec on synthetic R-MAT scale-%d (%d*2^%d tuples, a,b,c,d=.45/.15/.15/.25), p=%g q=%g, walk_length=%d, ONE pass: "
                    "one walk per non-isolated vertex, start list split over %d GPU(s)" % (args.shard_scale, args.edge_factor, args.shard_scale, p, q, L, world),
        "graph": {"nodes": g.n, "non_isolated": n, "directed_entries": g.nnz, "max_degree": g.max_degree},
        "graph_build_s": round(build_s, 3), "walk_preprocess_ms": round(prep_ms, 3),
        "walk_steps": int(steps_all), "device_ms": shard_ms, "value": steps_all / (shard_ms * 1e-3), "unit": "walk-steps/s",
        "one_gpu_ms": one_ms, "strong_scaling_efficiency": one_ms / (world * shard_ms),
        "e2e": {"seconds": e2e_s, "value": steps_all / e2e_s, "unit": "walk-steps/s", "handoff": ho,
                "h2d_bytes": n * 8, "d2h_bytes_per_rank": mine * L * 4, "corpus_matches_device_run": same,
                "limiter": "host link: every rank moves its own %d MB block through PCIe and the host's copy threads; "
                           "no NCCL traffic (gather = 0)" % (mine * L * 4 // (1 << 20))}}
    del d_out, d_starts, h_out, g
    torch.cuda.empty_cache()

    # ---------------- configs[4]: TopSim top-20, BA n=1e7 m=8, 1 M queries, tiles gathered over NCCL ----------------
    t0 = time.perf_counter()
    b = _lib.GraphHandle.barabasi_albert(args.ba_nodes, args.ba_m, seed=1)
    build_s = time.perf_counter() - t0
    nq = min(args.shard_queries, b.n)
    queries = np.random.RandomState(2).choice(b.n, size=nq, replace=False).astype(np.int64)
    comm.simrank_topk(b, queries[:8192 * world], 0.6, args.sr_step, args.sample, args.topk, seed=7)      # warm-up
    ids = np.zeros((nq, args.topk), dtype=np.int32)                            # result buffers committed before the clock starts
    sc = np.zeros((nq, args.topk), dtype=np.float64)
    runs = []
    for _ in range(2):                                                        # as for one_gpu_ms below: two runs, the faster one counts
        barrier()
        t0 = time.perf_counter()
        comm.simrank_topk(b, queries, 0.6, args.sr_step, args.sample, args.topk, seed=7, out=(ids, sc))
        e2e_r = allmax(time.perf_counter() - t0)
        c_ms, g_ms = comm.last_times()
        runs.append((allmax(c_ms), allmax(g_ms), e2e_r))
    comp_ms, gath_ms, e2e_s = min(runs, key=lambda r: r[0] + r[1])
    one_ms = comp_ms
    if world > 1:                                                             # the same 1 M queries on ONE GPU (every rank, max)
        d_q = torch.from_numpy(queries).to(dev)
        d_ids = torch.empty((nq, args.topk), dtype=torch.int32, device=dev)
        d_sc = torch.empty((nq, args.topk), dtype=torch.float64, device=dev)
        one_ms = None
        for _ in range(2):                                                    # first run sizes the scratch for 1 M queries
            barrier()
            ev[0].record()
            b.simrank_topk_dev(d_q.data_ptr(), nq, 0.6, args.sr_step, args.sample, args.topk, d_ids.data_ptr(), d_sc.data_ptr(),
                               seed=7, query_id_base=0, stream=stream)
            ev[1].record()
            barrier()
            t = allmax(ev[0].elapsed_time(ev[1]))
            one_ms = t if one_ms is None else min(one_ms, t)
        lo_q, hi_q = _lib.shard_range(nq, rank, world)
        k = min(2048, hi_q - lo_q)
        same = bool(np.array_equal(d_ids[lo_q:lo_q + k].cpu().numpy(), ids[lo_q:lo_q + k]))
    else:
        same = True
    out["topsim_ba%d" % args.ba_nodes] = {
        "workload": "TopSim SimRank top-%d on synthetic Barabasi-Albert n=%d m=%d, c=0.6 STEP=%d SAMPLE=%d, %d queries drawn without "
                    "replacement, split over %d GPU(s), top-k tiles gathered to every rank" % (args.topk, b.n, args.ba_m, args.sr_step, args.sample, nq, world),
        "graph_build_s": round(build_s, 3), "queries": nq,
        "device_ms": comp_ms, "device_ms_runs": [r[0] + r[1] for r in runs], "gather_ms": gath_ms, "gather_share": gath_ms / max(comp_ms + gath_ms, 1e-9),
        "gather_bytes_per_rank": nq * args.topk * 12,
        "value": nq / ((comp_ms + gath_ms) * 1e-3), "unit": "queries/s",
        "one_gpu_ms": one_ms, "strong_scaling_efficiency": one_ms / (world * (comp_ms + gath_ms)),
        "e2e": {"seconds": e2e_s, "value": nq / e2e_s, "unit": "queries/s", "h2d_bytes": nq * 8, "d2h_bytes_per_rank": nq * args.topk * 12},
        "gathered_equals_one_gpu": same}
    out["gpu_launches"] = int(_lib.kernel_launches() - launches0)
    comm.close()
    del b
    torch.cuda.empty_cache()
    return out


def measure(args, rank, world, local):
    import torch
    import torch.distributed as dist
    from graph_embedding_b200 import _lib
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream().cuda_stream
    peak, peak_src = measured_peak()
    launches0 = _lib.kernel_launches()
    metric, unit = metric_unit(args)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    extra = {}
    if args.workload == "node2vec":
        L = args.walk_length
        _lib.GraphHandle.rmat(8, 16 << 8, seed=1)              # CUDA context, module load: not part of the graph build
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ra, rb, rc = [float(x) for x in args.rmat_abc.split(",")]
        g = _lib.GraphHandle.rmat(args.scale, args.edge_factor << args.scale, a=ra, b=rb, c=rc, seed=1)
        starts_np = g.nonisolated()
        extra["graph_build_s"] = round(time.perf_counter() - t0, 3)
        extra["walk_preprocess_ms"] = round(g.prepare_walks(), 3)      # per-edge common-neighbour counts, one-off
        nw = len(starts_np)
        extra["graph"] = {"nodes": g.n, "non_isolated": nw, "directed_entries": g.nnz, "max_degree": g.max_degree}
        rs = np.random.RandomState(1234 + rank)
        d_out = torch.empty((nw, L), dtype=torch.int32, device=dev)
        perms = []
        for s in range(args.warmup + args.steps):
            perms.append(torch.from_numpy(rs.permutation(starts_np)).to(dev))    # random.shuffle(nodes), :51
        units_per_step = None
        second_order = not (args.p == 1.0 and args.q == 1.0)

        def step(i):
            g.walks_dev(args.p, args.q, L, perms[i].data_ptr(), nw, d_out.data_ptr(), seed=42,
                        walk_id_base=(rank * 1000 + i) * nw, stream=stream)

        for i in range(args.warmup):
            step(i)
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        launches_t0 = _lib.kernel_launches()
        ev[0].record()
        for i in range(args.steps):
            step(args.warmup + i)
            ev[i + 1].record()
        barrier()
        clocks = sampler.finish()
        gpu_launches = _lib.kernel_launches() - launches_t0
        ms_total = ev[0].elapsed_time(ev[-1])
        per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
        steps_exec, sum_s = g.byte_model_dev(d_out.data_ptr(), nw, L, second_order, stream=stream)
        units_per_step = steps_exec
        kernel_ms = float(np.mean(per_launch_ms))
        survey_bytes = 68.0 * steps_exec + 32.0 * sum_s          # 68 + 32*S(d_prev) per step
        mixture = os.environ.get("GW_WALKER") != "rejection"
        if mixture:
            # the production walker needs fewer bytes than the binary-search model: count ITS algorithmic
            # HBM bytes by re-running the last timed step's walks in counting mode
            last = args.warmup + args.steps - 1
            tr = g.walk_traffic_dev(args.p, args.q, L, perms[last].data_ptr(), nw, seed=42,
                                    walk_id_base=(rank * 1000 + last) * nw, stream=stream)
            assert tr["steps"] == steps_exec, (tr, steps_exec)
            # atom model (main): a random access moves one 64-byte HBM atom (the walkers load with L2::64B; ncu:
            # 61 B of dram__bytes_read per access, profiles/r1_gather_flavours_ncu.csv) + streamed rows + 4 B store
            alg_bytes = 64.0 * tr["random_accesses"] + tr["streamed_bytes"] + 4.0 * steps_exec
            sector_bytes = 32.0 * tr["random_accesses"] + tr["streamed_bytes"] + 4.0 * steps_exec
            kname = "k_walk_cn<VEC8=1,COUNT=0,MINB=5,HUB=%d,RIDX=%d>" % (
                int(g.max_degree > 2048), int(g.max_degree < 65536 and os.environ.get("GW_CN_RIDX") != "0"))
        else:
            alg_bytes, tr, sector_bytes = survey_bytes, None, survey_bytes
            kname = "k_walk_free<false,false>"
        wkey = "rmat%d/ef%d/abc=%s/p=%g/q=%g/L=%d" % (args.scale, args.edge_factor, args.rmat_abc, args.p, args.q, L)
        roof = {"bound": "hbm", "achieved": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "traffic": ncu_traffic(kname, wkey), "kernel": kname, "peak_source": peak_src,
                "bytes_per_unit": alg_bytes / steps_exec, "units_per_launch": steps_exec, "launch_ms": kernel_ms,
                "model": ("mixture walker, atom model: 64 B per random access (one {nbr,cnt|reverse index,offset,degree} entry per "
                          "non-return step; one Bloom word and, on a positive, S(d_prev) more per adjacency test of a context "
                          "that has common neighbours) + streamed "
                          "rows of intersections + 4 B store per step" if mixture else "SURVEY 8(d): 68 + 32*S(d_prev) per step"),
                "sector_model": {"bytes_per_unit": sector_bytes / steps_exec,
                                 "frac": sector_bytes / (kernel_ms * 1e-3) / 1e9 / peak,
                                 "note": "same accesses counted as 32-byte sectors (what the kernel consumes)"},
                "survey_model": {"bytes_per_unit": survey_bytes / steps_exec, "mean_search_sectors": sum_s / steps_exec,
                                 "achieved": survey_bytes / (kernel_ms * 1e-3) / 1e9,
                                 "frac": survey_bytes / (kernel_ms * 1e-3) / 1e9 / peak,
                                 "roofline_steps_per_s": peak * 1e9 / (survey_bytes / steps_exec),
                                 "note": "byte model of a binary-search walker (SURVEY 8d); > 1 means the sampler needs "
                                         "fewer bytes than that model, not that work is skipped"}}
        if tr:
            rate = tr["random_accesses"] / (kernel_ms * 1e-3) / 1e9
            ceil = gather_ceiling(16.0 * g.nnz)
            roof["random_accesses_per_step"] = tr["random_accesses"] / steps_exec
            roof["access_rate"] = {"achieved_G_per_s": rate, "ceiling_G_per_s": ceil, "frac": (rate / ceil) if ceil else None,
                                   "note": "dependent random loads/s over an array of nbr4's size, measured by tools/gather_bench.cu "
                                           "on this pool (profiles/gather_ceiling.json): translation-bound, independent of bytes per "
                                           "access -- the ceiling the byte roofline cannot see"}
            roof["intersections_per_step"] = tr["intersections"] / steps_exec
            roof["extra_proposals_per_step"] = tr["extra_proposals"] / steps_exec
        roof["frac"] = roof["achieved"] / peak

        # the reference workload is num_walks = 10 passes after ONE preprocessing: the honest whole-job rates
        ten = 10.0 * steps_exec
        extra["value_incl_preprocess"] = {
            "num_walks": 10, "unit": unit,
            "value": ten / ((10.0 * kernel_ms + extra["walk_preprocess_ms"]) * 1e-3),
            "value_incl_graph_build": ten / ((10.0 * kernel_ms + extra["walk_preprocess_ms"]) * 1e-3 + extra["graph_build_s"]),
            "note": "10 passes of the timed kernel + the one-off common-neighbour counts (stand-in for "
                    "preprocess_transition_probs, node2vec.py:99-108) [+ generating and building the graph]"}

        # ---- e2e: host buffers through the blocking C-ABI entry point ----
        e2e = None
        if not args.no_e2e:
            import ctypes
            h_starts = torch.from_numpy(rs.permutation(starts_np)).pin_memory()
            h_out = torch.empty((nw, L), dtype=torch.int32).pin_memory()
            L_ = _lib.load()

            def e2e_run(out_ptr, reps, seed0):
                def one(i):
                    _lib.check(L_.gw_node2vec_walks(g.h, args.p, args.q, L, ctypes.cast(h_starts.data_ptr(), _lib.c_i64p),
                                                    nw, 43, (rank * 1000 + seed0 + i) * nw,
                                                    ctypes.cast(out_ptr, _lib.c_i32p), None))
                one(0)
                barrier()
                t0 = time.perf_counter()
                for i in range(reps):
                    one(1 + i)
                torch.cuda.synchronize()
                tt = torch.tensor([time.perf_counter() - t0], device=dev)
                if world > 1:
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                return world * reps * steps_exec / float(tt.item())

            n_e2e = max(3, min(args.steps, 5))
            v = e2e_run(h_out.data_ptr(), n_e2e, 0)
            ho = g.last_handoff()
            e2e = {"value": v, "unit": unit, "h2d_bytes_per_step": nw * 8,
                   "d2h_bytes_per_step": nw * L * (3 if ho["mode"] == "packed" else 4) + (nw * 4 if ho["mode"] == "packed" else 0),
                   "steps": n_e2e, "api": "gw_node2vec_walks (pinned host start nodes in, pinned host corpus out)",
                   "handoff": ho}
            if not args.no_e2e_variants:
                variants = {}
                h_page = np.empty((nw, L), dtype=np.int32)              # what _lib.py / a JVM heap array hand over
                for name, env, ptr_ in (("direct_pinned", "direct", h_out.data_ptr()), ("ring_pinned", "ring", h_out.data_ptr()),
                                        ("packed_pinned", "packed", h_out.data_ptr()), ("ring_pageable", "ring", h_page.ctypes.data),
                                        ("packed_pageable", "packed", h_page.ctypes.data), ("default_pageable", None, h_page.ctypes.data)):
                    if env is None:
                        os.environ.pop("GW_E2E", None)
                    else:
                        os.environ["GW_E2E"] = env
                    variants[name] = {"value": e2e_run(ptr_, 2, 100), "handoff": g.last_handoff()["mode"]}
                os.environ.pop("GW_E2E", None)
                e2e["variants"] = variants
                if world == 1:
                    # the Python drop-in as a user calls it: node2vec.Graph(...).simulate_walks(1, L, as_array=True) --
                    # random.shuffle of the node list, id mapping and the pageable numpy corpus included
                    from graph_embedding_b200 import node2vec as n2v
                    G = n2v.Graph(n2v.EdgeListGraph(g), False, args.p, args.q)
                    import contextlib
                    import io
                    with contextlib.redirect_stdout(io.StringIO()):
                        G.simulate_walks(1, L, as_array=True)
                        t0 = time.perf_counter()
                        G.simulate_walks(1, L, as_array=True)
                        dt = time.perf_counter() - t0
                    e2e["python_drop_in"] = {"value": (g.n - (g.n - nw)) * (L - 1) / dt, "unit": unit, "seconds": dt,
                                             "api": "node2vec.Graph.simulate_walks(1, %d, as_array=True): random.shuffle of %d nodes + "
                                                    "walks + id mapping, pageable numpy corpus" % (L, g.n)}
    else:
        t0 = time.perf_counter()
        g = _lib.GraphHandle.barabasi_albert(args.ba_nodes, args.ba_m, seed=1)
        extra["graph_build_s"] = round(time.perf_counter() - t0, 3)
        extra["graph"] = {"nodes": g.n, "directed_entries": g.nnz, "max_degree": g.max_degree}
        nq = args.queries_per_step
        sr_mode = _lib.GW_SIMRANK_HYBRID if args.estimator == "hybrid" else _lib.GW_SIMRANK_MC
        rs = np.random.RandomState(2)
        total_q = nq * (args.warmup + args.steps) * world
        allq = rs.choice(g.n, size=min(total_q, g.n), replace=False).astype(np.int64)
        myq = allq[rank::world]
        d_q = torch.from_numpy(myq).to(dev)
        d_ids = torch.empty((nq, args.topk), dtype=torch.int32, device=dev)
        d_sc = torch.empty((nq, args.topk), dtype=torch.float64, device=dev)

        def step(i):
            g.simrank_topk_dev(d_q.data_ptr() + 8 * nq * i, nq, 0.6, args.sr_step, args.sample, args.topk,
                               d_ids.data_ptr(), d_sc.data_ptr(), mode=sr_mode, seed=7,
                               query_id_base=(rank * 100000 + i) * nq, stream=stream)
        for i in range(args.warmup):
            step(i)
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        launches_t0 = _lib.kernel_launches()
        ev[0].record()
        for i in range(args.steps):
            step(args.warmup + i)
            ev[i + 1].record()
        barrier()
        clocks = sampler.finish()
        gpu_launches = _lib.kernel_launches() - launches_t0
        ms_total = ev[0].elapsed_time(ev[-1])
        per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
        units_per_step = nq
        walk_steps = g.simrank_last_steps()
        extra["slow_path_queries_last_step"] = g.simrank_last_slow_queries()
        kernel_ms = float(np.mean(per_launch_ms))
        alg_bytes = walk_steps * 64.0 + nq * args.topk * 12.0
        kname = ("k_simrank_log<%d>" if args.estimator == "mc" else "k_topsim_hybrid<%d,true>") % args.sr_step
        rate = walk_steps / (kernel_ms * 1e-3) / 1e9
        ceil = gather_ceiling(16.0 * g.nnz)
        roof = {"bound": "hbm", "achieved": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "traffic": ncu_traffic(kname, "ba%d/m%d/sample=%d/step=%d/k=%d" % (args.ba_nodes, args.ba_m, args.sample, args.sr_step, args.topk), nq),
                "kernel": kname, "peak_source": peak_src,
                "bytes_per_unit": alg_bytes / nq, "units_per_launch": nq, "launch_ms": kernel_ms,
                "model": "SURVEY 8(d): 64 B per walk step (here ONE random 16-byte entry = one 64-byte HBM atom) + 12 B per result slot",
                "walk_steps_per_s": walk_steps / (kernel_ms * 1e-3),
                "access_rate": {"achieved_G_per_s": rate, "ceiling_G_per_s": ceil, "frac": (rate / ceil) if ceil else None,
                                "note": "one dependent random load per walk step; ceiling = DRAM-missing dependent loads/s measured by "
                                        "tools/gather_bench.cu for an array of nbr4's size in a many-wave launch (profiles/gather_ceiling.json; "
                                        "equal work per SM ends at 39.6 G/s because the SMs come in three throughput classes, profiles/r2_gather_waves.txt).  The Monte-Carlo "
                                        "kernel's first two steps of every walk start at the query vertex and hit L1/L2 (~20 % of its loads), "
                                        "so its rate can exceed the ceiling of misses"}}
        roof["frac"] = roof["achieved"] / peak
        e2e = None
        if not args.no_e2e:
            hq = myq[:nq].copy()
            g.simrank_topk(hq, 0.6, args.sr_step, args.sample, args.topk, mode=sr_mode, seed=8)
            barrier()
            n_e2e = max(3, min(args.steps, 5))
            t0 = time.perf_counter()
            for i in range(n_e2e):
                g.simrank_topk(hq, 0.6, args.sr_step, args.sample, args.topk, mode=sr_mode, seed=9 + i)
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e = {"value": world * n_e2e * nq / float(tt.item()), "unit": unit, "h2d_bytes_per_step": nq * 8,
                   "d2h_bytes_per_step": nq * args.topk * 12, "steps": n_e2e,
                   "api": "gw_simrank_topk (host queries in, host top-k out)"}

    # max over ranks of the device-timed region
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * args.steps * units_per_step / (ms_total * 1e-3)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if args.workload == "node2vec":
            v, sample, _ = cpu_node2vec(args.p, args.q, args.walk_length, args.cpu_seconds, 1)
        else:
            v, sample = cpu_simrank(args.sample, args.sr_step, args.topk, args.cpu_seconds, 1)
        cpu = {"value": v, "unit": unit, "cores": 1, "kind": "port", "sample": sample}

    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "arith": ("vertex ids / offsets int32; mixture-component masses fp32 with 24-bit uniforms (fp64 masses and a 32-bit "
                      "uniform from degree 4096 up), neighbour index = umulhi(32-bit uniform, degree)" if args.workload == "node2vec" else
                      "vertex ids int32; increments C^i*deg/deg/SAMPLE in fp32, accumulated as 32.32 fixed point (exact integer adds), "
                      "scores leave as fp64"),
            "config": config_of(args), "roofline": roof, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(gpu_launches), "clocks": clocks, "per_launch_ms": [round(x, 3) for x in per_launch_ms]}
    line.update(extra)
    del g
    torch.cuda.empty_cache()
    return line


if __name__ == "__main__":
    main()
