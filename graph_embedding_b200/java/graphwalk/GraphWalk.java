package graphwalk;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.foreign.ValueLayout;
import java.lang.invoke.MethodHandle;

/**
 * Panama FFM (JDK 22+) binding of libgraphwalk.so — include/graphwalk.h.
 *
 * NOT COMPILED IN THIS REPOSITORY'S IMAGE: the build container has no JDK (SURVEY.md §8c), so
 * this file is the binding a maintainer adds to DeepSim/TopSimAll/src; the same C ABI is
 * exercised end to end from Python (graph_embedding_b200/_lib.py, tests/).  Every downcall
 * mirrors one prototype of graphwalk.h; status != 0 raises with gw_last_error().
 */
public final class GraphWalk {
    public static final int MODE_SIMPLE = 0, MODE_MULTI = 1;
    public static final int SIMRANK_MC = 0, SIMRANK_HYBRID = 1;

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB =
        SymbolLookup.libraryLookup(System.getProperty("graphwalk.lib", "libgraphwalk.so"), Arena.global());

    private static MethodHandle h(String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), fd);
    }

    private static final ValueLayout.OfInt I = ValueLayout.JAVA_INT;
    private static final ValueLayout.OfLong J = ValueLayout.JAVA_LONG;
    private static final ValueLayout.OfDouble D = ValueLayout.JAVA_DOUBLE;
    private static final ValueLayout P = ValueLayout.ADDRESS;

    private static final MethodHandle LAST_ERROR = h("gw_last_error", FunctionDescriptor.of(P));
    private static final MethodHandle SET_DEVICE = h("gw_set_device", FunctionDescriptor.of(I, I));
    private static final MethodHandle LOAD = h("gw_graph_load_edgelist",
        FunctionDescriptor.of(I, P, P, I, I, I, J, P));
    private static final MethodHandle FREE = h("gw_graph_free", FunctionDescriptor.of(I, P));
    private static final MethodHandle INFO = h("gw_graph_info", FunctionDescriptor.of(I, P, P, P, P, P, P));
    private static final MethodHandle CSR = h("gw_graph_csr", FunctionDescriptor.of(I, P, P, P, P, P, P));
    private static final MethodHandle TOPK = h("gw_simrank_topk",
        FunctionDescriptor.of(I, P, P, J, D, I, I, I, I, J, J, P, P));
    private static final MethodHandle ROWS = h("gw_simrank_rows",
        FunctionDescriptor.of(I, P, P, J, D, I, I, I, J, J, P));
    private static final MethodHandle EXACT = h("gw_simrank_exact",
        FunctionDescriptor.of(I, P, D, I, P, J, P));
    private static final MethodHandle ROWS_JAVARNG = h("gw_simrank_rows_javarng",
        FunctionDescriptor.of(I, P, P, J, D, I, I, P, P));

    private static final ValueLayout.OfFloat F = ValueLayout.JAVA_FLOAT;
    private static final MethodHandle TOPSIM_JAVARNG = h("gw_topsim_rows_javarng",
        FunctionDescriptor.of(I, P, P, J, D, I, I, I, J, P, P));
    private static final MethodHandle CACHE_JAVARNG = h("gw_simrank_cache_javarng",
        FunctionDescriptor.of(I, P, P, J, D, I, I, I, I, J, P, P, P, P));
    private static final MethodHandle DW_PATHS = h("gw_double_walk_paths",
        FunctionDescriptor.of(I, P, P, J, I, I, J, P, P));
    private static final MethodHandle DW_SIMS = h("gw_double_walk_sims",
        FunctionDescriptor.of(I, P, P, J, I, I, D, P, J, I, P));
    private static final MethodHandle MASS = h("gw_topsim_mass",
        FunctionDescriptor.of(I, P, P, J, D, I, J, J, J, P, P));
    private static final MethodHandle MASS_SIMS = h("gw_topsim_mass_sims",
        FunctionDescriptor.of(I, P, P, J, I, D, P, P, J, I, P));
    private static final MethodHandle COMM_ID = h("gw_comm_unique_id", FunctionDescriptor.of(I, P));
    private static final MethodHandle COMM_INIT = h("gw_comm_init", FunctionDescriptor.of(I, I, I, P, I, P));
    private static final MethodHandle COMM_FREE = h("gw_comm_free", FunctionDescriptor.of(I, P));
    private static final MethodHandle TOPK_SHARDED = h("gw_simrank_topk_sharded",
        FunctionDescriptor.of(I, P, P, P, J, D, I, I, I, I, J, P, P));
    private static final MethodHandle WALKS_SHARDED = h("gw_node2vec_walks_sharded",
        FunctionDescriptor.of(I, P, P, D, D, I, P, J, J, I, P, P));

    private GraphWalk() {}

    static void check(int rc) {
        if (rc == 0) return;
        try {
            MemorySegment msg = ((MemorySegment) LAST_ERROR.invokeExact()).reinterpret(1024);
            throw new IllegalStateException("libgraphwalk error " + rc + ": " + msg.getString(0));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException("libgraphwalk error " + rc, t);
        }
    }

    public static void setDevice(int device) {
        try { check((int) SET_DEVICE.invokeExact(device)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new AssertionError(t); }
    }

    /** gw_graph_load_edgelist(path, delimiter, weighted=0, directed=0, MULTI, V) -> handle */
    public static MemorySegment loadMultigraph(String path, String separator, long vCount) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(P);
            check((int) LOAD.invokeExact(a.allocateFrom(path), a.allocateFrom(separator), 0, 0, MODE_MULTI, vCount, out));
            return out.get(P, 0);
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new AssertionError(t); }
    }

    public static void free(MemorySegment g) {
        try { check((int) FREE.invokeExact(g)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new AssertionError(t); }
    }

    /** returns {n_nodes, n_entries} */
    public static long[] info(MemorySegment g) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment n = a.allocate(J), nnz = a.allocate(J);
            check((int) INFO.invokeExact(g, n, nnz, MemorySegment.NULL, MemorySegment.NULL, MemorySegment.NULL));
            return new long[] {n.get(J, 0), nnz.get(J, 0)};
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new AssertionError(t); }
    }

    /** CSR copy: rowPtr[n+1], col[nnz] (file order inside each row, duplicates kept). */
    public static void csr(MemorySegment g, long[] rowPtr, int[] col) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment rp = a.allocate(J, rowPtr.length), c = a.allocate(I, Math.max(col.length, 1));
            check((int) CSR.invokeExact(g, rp, c, MemorySegment.NULL, MemorySegment.NULL, MemorySegment.NULL));
            MemorySegment.copy(rp, J, 0, rowPtr, 0, rowPtr.length);
            MemorySegment.copy(c, I, 0, col, 0, col.length);
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new AssertionError(t); }
    }

    /** gw_simrank_topk: ids[nq*k], scores[nq*k] (score descending, id ascending, -1 padding). */
    public static void simrankTopk(MemorySegment g, long[] queries, double c, int step, int sample, int k, int mode,
                                   long seed, int[] outIds, double[] outScores) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment q = a.allocateFrom(J, queries);
            MemorySegment ids = a.allocate(I, (long) queries.length * k), sc = a.allocate(D, (long) queries.length * k);
            check((int) TOPK.invokeExact(g, q, (long) queries.length, c, step, sample, k, mode, seed, 0L, ids, sc));
            MemorySegment.copy(ids, I, 0, outIds, 0, outIds.length);
            MemorySegment.copy(sc, D, 0, outScores, 0, outScores.length);
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new AssertionError(t); }
    }

    /**
     * gw_simrank_rows_javarng: replay mode.  rngState[i] is the 48-bit java.util.Random state in front of query i
     * ((seed ^ 0x5DEECE66DL) & ((1L << 48) - 1) for a fresh Random(seed)); it is updated to the state after
     * the query, so a JVM run with a seeded Graph.rand is reproduced bit for bit.
     */
    public static double[][] simrankRowsJavaRng(MemorySegment g, long[] queries, int vCount, double c, int step, int sample,
                                                long[] rngState) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment q = a.allocateFrom(J, queries);
            MemorySegment st = a.allocateFrom(J, rngState);
            MemorySegment out = a.allocate(D, (long) queries.length * vCount);
            check((int) ROWS_JAVARNG.invokeExact(g, q, (long) queries.length, c, step, sample, st, out));
            MemorySegment.copy(st, J, 0, rngState, 0, rngState.length);
            double[][] sim = new double[queries.length][vCount];
            for (int r = 0; r < queries.length; r++) MemorySegment.copy(out, D, (long) r * vCount * 8, sim[r], 0, vCount);
            return sim;
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new AssertionError(t); }
    }

    /** gw_simrank_rows: dense rows sim[query][*] (the reference's double[][] result). */
    public static double[][] simrankRows(MemorySegment g, long[] queries, int vCount, double c, int step, int sample,
                                         int mode, long seed) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment q = a.allocateFrom(J, queries);
            MemorySegment out = a.allocate(D, (long) queries.length * vCount);
            check((int) ROWS.invokeExact(g, q, (long) queries.length, c, step, sample, mode, seed, 0L, out));
            double[][] sim = new double[queries.length][vCount];
            for (int r = 0; r < queries.length; r++) MemorySegment.copy(out, D, (long) r * vCount * 8, sim[r], 0, vCount);
            return sim;
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new AssertionError(t); }
    }

    /** gw_simrank_exact: all rows of the exact iteration (SimRank.java). */
    public static double[][] simrankExact(MemorySegment g, int vCount, double c, int iters) {
        try (Arena a = Arena.ofConfined()) {
            long[] rows = new long[vCount];
            for (int i = 0; i < vCount; i++) rows[i] = i;
            MemorySegment r = a.allocateFrom(J, rows);
            MemorySegment out = a.allocate(D, (long) vCount * vCount);
            check((int) EXACT.invokeExact(g, c, iters, r, (long) vCount, out));
            double[][] sim = new double[vCount][vCount];
            for (int i = 0; i < vCount; i++) MemorySegment.copy(out, D, (long) i * vCount * 8, sim[i], 0, vCount);
            return sim;
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new AssertionError(t); }
    }

    /** gw_topsim_rows_javarng: replay of TopSim_singleSample (mode 0) / TopSim_Enumerate (mode 1); rows x SAMPLE. */
    public static double[][] topsimRowsJavaRng(MemorySegment g, long[] queries, int vCount, double c, int step, int sample,
                                               int mode, long maxPaths, long[] rngState) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment q = a.allocateFrom(J, queries), st = a.allocateFrom(J, rngState);
            MemorySegment out = a.allocate(D, (long) queries.length * vCount);
            check((int) TOPSIM_JAVARNG.invokeExact(g, q, (long) queries.length, c, step, sample, mode, maxPaths, st, out));
            MemorySegment.copy(st, J, 0, rngState, 0, rngState.length);
            double[][] sim = new double[queries.length][vCount];
            for (int r = 0; r < queries.length; r++) MemorySegment.copy(out, D, (long) r * vCount * 8, sim[r], 0, vCount);
            return sim;
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new AssertionError(t); }
    }

    /**
     * gw_simrank_cache_javarng: replay of SingleRandomWalk_M (mode 0) / TopSim_singleSample_M (mode 1).  keys/vals
     * [nq * capacity] receive heap slots 1..capacity of every query's FixedCacheMap, sizes[nq] the live entries.
     */
    public static void simrankCacheJavaRng(MemorySegment g, long[] queries, double c, int step, int sample, int mode,
                                           int capacity, long maxPaths, long[] rngState, int[] keys, float[] vals, int[] sizes) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment q = a.allocateFrom(J, queries), st = a.allocateFrom(J, rngState);
            MemorySegment k = a.allocate(I, keys.length), v = a.allocate(F, vals.length), z = a.allocate(I, sizes.length);
            check((int) CACHE_JAVARNG.invokeExact(g, q, (long) queries.length, c, step, sample, mode, capacity, maxPaths, st, k, v, z));
            MemorySegment.copy(st, J, 0, rngState, 0, rngState.length);
            MemorySegment.copy(k, I, 0, keys, 0, keys.length);
            MemorySegment.copy(v, F, 0, vals, 0, vals.length);
            MemorySegment.copy(z, I, 0, sizes, 0, sizes.length);
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new AssertionError(t); }
    }

    /** gw_double_walk_paths: paths[v][i][step] flattened; rngState == null -> Philox(seed), else java.util.Random replay. */
    public static int[] doubleWalkPaths(MemorySegment g, long[] vertices, int sample, int step, long seed, long[] rngState) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment v = a.allocateFrom(J, vertices);
            MemorySegment st = rngState == null ? MemorySegment.NULL : a.allocateFrom(J, rngState);
            long total = (long) vertices.length * sample * step;
            MemorySegment out = a.allocate(I, total);
            check((int) DW_PATHS.invokeExact(g, v, (long) vertices.length, sample, step, seed, st, out));
            if (rngState != null) MemorySegment.copy(st, J, 0, rngState, 0, rngState.length);
            return out.toArray(I);
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new AssertionError(t); }
    }

    /** gw_double_walk_sims: sim[row][*] of DoubleRandomWalk.getSim over a path set of nv vertices. */
    public static double[][] doubleWalkSims(MemorySegment g, int[] paths, int nv, int sample, int step, double c, long[] rows,
                                            boolean exactOrder) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment p = a.allocateFrom(I, paths), r = a.allocateFrom(J, rows);
            MemorySegment out = a.allocate(D, (long) rows.length * nv);
            check((int) DW_SIMS.invokeExact(g, p, (long) nv, sample, step, c, r, (long) rows.length, exactOrder ? 1 : 0, out));
            double[][] sim = new double[rows.length][nv];
            for (int i = 0; i < rows.length; i++) MemorySegment.copy(out, D, (long) i * nv * 8, sim[i], 0, nv);
            return sim;
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new AssertionError(t); }
    }

    /** gw_topsim_mass: path-mass trees of TopSim_doubleSample / TopSim_Dev, [ns][n][step+1] flattened (-1 = unset). */
    public static double[] topsimMass(MemorySegment g, long[] sources, int vCount, double weight, int step, long maxPaths,
                                      long seed, long callIdBase, long[] rngState) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment s = a.allocateFrom(J, sources);
            MemorySegment st = rngState == null ? MemorySegment.NULL : a.allocateFrom(J, rngState);
            MemorySegment out = a.allocate(D, (long) sources.length * vCount * (step + 1));
            check((int) MASS.invokeExact(g, s, (long) sources.length, weight, step, maxPaths, seed, callIdBase, st, out));
            if (rngState != null) MemorySegment.copy(st, J, 0, rngState, 0, rngState.length);
            return out.toArray(D);
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new AssertionError(t); }
    }

    /** gw_topsim_mass_sims: getSim for (a, b) index pairs into a mass set of ns trees. */
    public static double[] topsimMassSims(MemorySegment g, double[] mass, long ns, int step, double c, long[] pairA, long[] pairB,
                                          boolean exactOrder) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment m = a.allocateFrom(D, mass), pa = a.allocateFrom(J, pairA), pb = a.allocateFrom(J, pairB);
            MemorySegment out = a.allocate(D, pairA.length);
            check((int) MASS_SIMS.invokeExact(g, m, ns, step, c, pa, pb, (long) pairA.length, exactOrder ? 1 : 0, out));
            return out.toArray(D);
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new AssertionError(t); }
    }

    /** gw_comm_unique_id: 128 bytes created on rank 0 and shipped to the other ranks by the launcher. */
    public static byte[] commUniqueId() {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment id = a.allocate(128);
            check((int) COMM_ID.invokeExact(id));
            return id.toArray(ValueLayout.JAVA_BYTE);
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new AssertionError(t); }
    }

    /** gw_comm_init: one JVM per GPU; returns the communicator handle (free with commFree). */
    public static MemorySegment commInit(int rank, int nranks, byte[] uniqueId, int device) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment id = a.allocateFrom(ValueLayout.JAVA_BYTE, uniqueId), out = a.allocate(P);
            check((int) COMM_INIT.invokeExact(rank, nranks, id, device, out));
            return out.get(P, 0);
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new AssertionError(t); }
    }

    public static void commFree(MemorySegment comm) {
        try { check((int) COMM_FREE.invokeExact(comm)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new AssertionError(t); }
    }

    /** gw_simrank_topk_sharded: every rank passes the same queries and receives the whole [nq*k] result. */
    public static void simrankTopkSharded(MemorySegment g, MemorySegment comm, long[] queries, double c, int step, int sample,
                                          int k, int mode, long seed, int[] outIds, double[] outScores) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment q = a.allocateFrom(J, queries);
            MemorySegment ids = a.allocate(I, (long) queries.length * k), sc = a.allocate(D, (long) queries.length * k);
            check((int) TOPK_SHARDED.invokeExact(g, comm, q, (long) queries.length, c, step, sample, k, mode, seed, ids, sc));
            MemorySegment.copy(ids, I, 0, outIds, 0, outIds.length);
            MemorySegment.copy(sc, D, 0, outScores, 0, outScores.length);
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new AssertionError(t); }
    }
}
