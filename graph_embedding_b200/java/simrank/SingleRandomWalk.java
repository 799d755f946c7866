package simrank;

import conf.MyConfiguration;
import graphwalk.GraphWalk;
import structures.Graph;

/**
 * Drop-in for DeepSim/TopSimAll/src/simrank/SingleRandomWalk.java: same constructor,
 * compute() and getResult(); the SAMPLE x 2*STEP walks, first-meeting accumulation and the
 * division by SAMPLE run in the fused walk-and-meet kernel (gw_simrank_rows / gw_simrank_topk).
 * benchmark/Test_u_u_SingleRandomWalk_Sample.java compiles against it unchanged.
 * Untested in this repository's image (no JDK) — see INTEGRATION.md.
 */
public class SingleRandomWalk {
    protected final int topk = MyConfiguration.TOPK;
    protected int STEP = 1;
    protected int COUNT;
    protected Graph g;
    protected double[][] sim;
    public static int SAMPLE = 10000;
    protected int mode = GraphWalk.SIMRANK_MC;
    protected long seed = System.nanoTime();               // the reference RNG is unseeded (Graph.java:17)

    public SingleRandomWalk(Graph g, int sample, int step) {
        this.STEP = step;
        SAMPLE = sample;
        this.g = g;
        this.COUNT = g.getVCount();
    }

    public void compute() {
        long[] q = new long[COUNT];
        for (int i = 0; i < COUNT; i++) q[i] = i;
        sim = GraphWalk.simrankRows(g.nativeHandle(), q, COUNT, MyConfiguration.C, STEP, SAMPLE, mode, seed);
    }

    /** Per-query top-k straight from the device, without the V x V matrix. */
    public void topk(long[] queries, int k, int[] outIds, double[] outScores) {
        GraphWalk.simrankTopk(g.nativeHandle(), queries, MyConfiguration.C, STEP, SAMPLE, k, mode, seed, outIds, outScores);
    }

    public double[][] getResult() { return sim; }
}
