package simrank;

import conf.MyConfiguration;
import graphwalk.GraphWalk;
import structures.Graph;

/**
 * Drop-in for DeepSim/TopSimAll/src/simrank/DoubleRandomWalk.java: same constructor, compute() and getResult().
 * samplePaths (:50-65) and the SAMPLE^2 pair scan of getSim (:77-91) run on the device (gw_double_walk_paths,
 * gw_double_walk_sims).  benchmark/Test_u_u_doubleRandomWalk_Sample.java compiles against it unchanged.
 * Untested in this repository's image (no JDK) -- see INTEGRATION.md.
 */
public class DoubleRandomWalk {
    protected final int topk = MyConfiguration.TOPK;
    protected int STEP = 3;
    protected int COUNT;
    protected Graph g;
    protected int[] paths;                                  // [COUNT][SAMPLE][STEP] flattened
    protected double[][] sim;
    public static int SAMPLE = 200;
    protected long seed = System.nanoTime();               // the reference RNG is unseeded (Graph.java:17)

    public DoubleRandomWalk(Graph g, int sample, int step) {
        SAMPLE = sample;
        this.STEP = step;
        this.g = g;
        this.COUNT = g.getVCount();
    }

    public void compute() {
        long[] v = new long[COUNT];
        for (int i = 0; i < COUNT; i++) v[i] = i;
        paths = GraphWalk.doubleWalkPaths(g.nativeHandle(), v, SAMPLE, STEP, seed, null);
        sim = GraphWalk.doubleWalkSims(g.nativeHandle(), paths, COUNT, SAMPLE, STEP, MyConfiguration.C, v, false);
    }

    public double[][] getResult() { return sim; }
}
