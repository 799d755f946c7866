package structures;

import java.io.IOException;
import java.lang.foreign.MemorySegment;
import java.util.AbstractList;
import java.util.List;

import conf.MyConfiguration;
import graphwalk.GraphWalk;

/**
 * Drop-in for DeepSim/TopSimAll/src/structures/Graph.java (same constructor, degree, neighbors,
 * getVCount, getECount): the adjacency lives in B200 HBM, built by gw_graph_load_edgelist in
 * MULTI mode (V slots, both directions per line, duplicates and file order kept).
 * Untested in this repository's image (no JDK) — see INTEGRATION.md.
 */
public class Graph {
    public int vCount;
    public int eCount;
    final MemorySegment handle;
    private long[] rowPtr;
    private int[] col;

    public Graph(String graphPath, int V) throws IOException {
        try {
            this.handle = GraphWalk.loadMultigraph(graphPath, MyConfiguration.SEPARATOR, V);
        } catch (IllegalStateException e) {
            throw new IOException(e.getMessage(), e);          // the reference propagates IOException
        }
        long[] info = GraphWalk.info(handle);
        this.vCount = (int) info[0];
        this.eCount = (int) (info[1] / 2);
    }

    public MemorySegment nativeHandle() { return handle; }

    private void host() {
        if (rowPtr != null) return;
        rowPtr = new long[vCount + 1];
        col = new int[2 * eCount];
        GraphWalk.csr(handle, rowPtr, col);
    }

    public int degree(int v) { host(); return (int) (rowPtr[v + 1] - rowPtr[v]); }

    public List<Integer> neighbors(int v) {
        host();
        final int base = (int) rowPtr[v], d = degree(v);
        return new AbstractList<Integer>() {
            @Override public Integer get(int i) { return col[base + i]; }
            @Override public int size() { return d; }
        };
    }

    public int getVCount() { return vCount; }
    public int getECount() { return eCount; }
}
