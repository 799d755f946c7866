/* graphwalk.h — C ABI of libgraphwalk.so, the B200 (sm_100a) random-walk engine.
 *
 * This is the drop-in boundary for the ONE hot path of Junshuai-Song/Graph-Embedding:
 *   node2vec second-order walk generation   (reference: node2vec/src/node2vec.py)
 *   DeepSim/TopSim Monte-Carlo SimRank       (reference: DeepSim/TopSimAll/src/simrank/*.java)
 * The reference has no FFI layer of its own; every entry point below names the
 * reference function it replaces.  Host bindings: Python ctypes (graph_embedding_b200/_lib.py),
 * Java Panama FFM / JNI (graph_embedding_b200/java/).
 *
 * Conventions
 *   - plain C types only; every function returns 0 (GW_OK) or a negative GW_E_* code and
 *     gw_last_error() returns a thread-local message for the last failure on this thread;
 *   - handles are opaque; output buffers are caller-owned and sized by the caller (sizes come
 *     from gw_graph_info / gw_alias_edges_size); the library never frees caller memory;
 *   - functions without the _dev suffix take HOST pointers and block until the result is in
 *     host memory (H2D copy, kernels, D2H copy inside the call);
 *   - functions with the _dev suffix take DEVICE pointers on the graph's device plus a
 *     cudaStream_t (passed as void*, NULL = legacy default stream) and only enqueue work;
 *   - vertices are addressed by DENSE index 0..n-1 = rank of the original id in ascending
 *     order (SIMPLE mode) or the original id itself (MULTI mode / generators);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     GW_E_CUDA.
 */
#ifndef GRAPHWALK_H
#define GRAPHWALK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GW_OK 0
#define GW_E_INVALID (-1)   /* bad argument */
#define GW_E_CUDA (-2)      /* CUDA runtime error / no device */
#define GW_E_IO (-3)        /* file could not be read / parsed */
#define GW_E_TOO_LARGE (-4) /* result would exceed the library's / caller's budget */
#define GW_E_STATE (-5)     /* call order violated (e.g. tables not built) */
#define GW_E_KEY (-6)       /* unknown vertex / edge (the reference raises KeyError) */

/* graph construction modes */
#define GW_MODE_SIMPLE 0 /* networkx semantics of node2vec/src/main.py:76-89: ids ranked, duplicates
                            collapse, sorted adjacency */
#define GW_MODE_MULTI 1  /* structures/Graph.java:28-57: V slots, every line appends both directions,
                            duplicates and file order kept */

/* graph flags reported by gw_graph_info */
#define GW_F_DIRECTED 1
#define GW_F_WEIGHTED 2
#define GW_F_MULTI 4

/* SimRank estimator modes */
#define GW_SIMRANK_MC 0     /* simrank/SingleRandomWalk.java:53-106 (scores / SAMPLE) */
#define GW_SIMRANK_HYBRID 1 /* simrank/TopSim_singleSample.java:62-203 (scores x SAMPLE, as the reference) */
#define GW_SIMRANK_MC_F64 2 /* gw_simrank_rows only: the walks of GW_SIMRANK_MC (same seed => same paths) with every
                               increment evaluated and added in fp64 as SingleRandomWalk.java:89 -- the arithmetic
                               reference that bounds the production kernels' fp32 / 32.32 fixed-point accumulation */

typedef struct gw_graph gw_graph;

/* ---- library / device ------------------------------------------------------------------- */
int gw_version(void);
const char *gw_last_error(void);
int gw_device_count(int *count);
/* Binds the calling thread (and graphs created afterwards) to a CUDA device. */
int gw_set_device(int device);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t gw_kernel_launches(void);

/* ---- graph loading  (node2vec/src/main.py:76-89 read_graph ; structures/Graph.java:28-57) ---- */
/* Builds the device CSR from host edge arrays.  src/dst: original ids, w: NULL => weight 1
 * (unweighted).  n_slots: MULTI mode = vertex count V of Graph(path, V); SIMPLE mode = -1 (rank
 * the ids that occur) or n > max id to keep ids as dense indices (generators). */
int gw_graph_from_edges(const int64_t *src, const int64_t *dst, const double *w, int64_t m,
                        int directed, int mode, int64_t n_slots, gw_graph **out);
/* Parses "u<delim>v[<delim>w]" lines like networkx.read_edgelist (SIMPLE; '#' comments, lines
 * with < 2 fields skipped; delimiter NULL/"" = any whitespace) or like Graph(String,int)
 * (MULTI; String.split(delim)), then calls gw_graph_from_edges.  '.gz' is not handled here. */
int gw_graph_load_edgelist(const char *path, const char *delimiter, int weighted, int directed,
                           int mode, int64_t n_slots, gw_graph **out);
/* Synthetic inputs, generated on the device (benchmark shapes of BASELINE.json; the quadrant
 * probabilities default to utils/graphTools/RMATGraphGenerator.java:179-182 = .45/.15/.15/.25).
 * R-MAT: 2^scale ids, n_tuples edge tuples, symmetrised, self loops and duplicates dropped. */
int gw_graph_rmat(int scale, int64_t n_tuples, double a, double b, double c, uint64_t seed,
                  gw_graph **out);
/* Barabasi-Albert: n vertices, each new vertex attaches to m distinct targets ~ degree;
 * seed graph = m-clique.  Undirected, simple. */
int gw_graph_barabasi_albert(int64_t n, int m, uint64_t seed, gw_graph **out);
int gw_graph_free(gw_graph *g);

int gw_graph_info(const gw_graph *g, int64_t *n_nodes, int64_t *n_entries, int32_t *flags,
                  int32_t *max_degree, int32_t *device);
/* Copies the CSR to host.  Any pointer may be NULL.  row_ptr[n+1], col_idx[nnz], weights[nnz],
 * node_ids[n] (original ids, ascending), first_seen[n] (dense indices in list(G.nodes()) order,
 * node2vec.py:47; identity for generated graphs). */
int gw_graph_csr(const gw_graph *g, int64_t *row_ptr, int32_t *col_idx, double *weights,
                 int64_t *node_ids, int64_t *first_seen);
/* Device views (borrowed; valid until gw_graph_free): meta[n] = {uint32 offset, uint32 degree},
 * col[nnz] int32. */
int gw_graph_device_views(const gw_graph *g, const void **meta, const int32_t **col);
/* Dense indices of vertices with degree > 0, ascending (the reference only knows nodes that
 * occur in an edge).  out may be NULL to query the count. */
int gw_graph_nonisolated(const gw_graph *g, int64_t *out, int64_t *count);

/* ---- alias tables  (node2vec.py:116-147 alias_setup, :83-113 preprocess_transition_probs) ---- */
/* One table: bit-exact alias_setup(probs) on the device (J as int32, q as fp64). */
int gw_alias_setup(const double *probs, int64_t K, int32_t *J, double *q);
/* alias_nodes for every vertex, flattened at CSR offsets (J[nnz], q[nnz]).  Builds and keeps
 * the device copy; J/q may be NULL to only build. */
int gw_alias_nodes(gw_graph *g, int32_t *J, double *q);
/* Total entries of alias_edges = sum over directed CSR entries (u->v) of deg(v). */
int gw_alias_edges_size(const gw_graph *g, int64_t *total);
/* alias_edges for every directed CSR entry e=(u->v), table e at off[e]..off[e+1] (off[nnz+1]).
 * Builds and keeps the device copy for replay; fails with GW_E_TOO_LARGE above budget_bytes
 * (0 = 3/4 of free device memory).  off/J/q may be NULL to only build. */
int gw_alias_edges(gw_graph *g, double p, double q, int64_t budget_bytes, int64_t *off,
                   int32_t *J, double *qv);

/* ---- node2vec walks  (node2vec.py:13-59 node2vec_walk / simulate_walks) ---- */
/* Free-running walks, one per entry of starts[] (dense indices), Philox4x32-10 keyed by
 * (seed, walk_id_base + i): the corpus does not depend on how starts are split over calls,
 * ranks or GPUs.  out_walks[n_starts*walk_length] int32 dense indices, -1 padded after a dead
 * end (node2vec.py:36-37); out_lens[n_starts] may be NULL. */
int gw_node2vec_walks(gw_graph *g, double p, double q, int32_t walk_length, const int64_t *starts,
                      int64_t n_starts, uint64_t seed, uint64_t walk_id_base, int32_t *out_walks,
                      int32_t *out_lens);
int gw_node2vec_walks_dev(gw_graph *g, double p, double q, int32_t walk_length,
                          const int64_t *d_starts, int64_t n_starts, uint64_t seed,
                          uint64_t walk_id_base, int32_t *d_out_walks, int32_t *d_out_lens,
                          void *stream);
/* How the last gw_node2vec_walks call on this graph moved its corpus to the host: mode 1 = direct DMA
 * into the caller's page-locked buffer, 2 = pinned ring drained by the library's copy threads (pageable buffers),
 * 3 = the same ring with ids crossing PCIe as 3 bytes (graphs of <= 2^24 vertices); copy_threads = size of the pool.
 * Either pointer may be NULL. */
int gw_graph_last_handoff(const gw_graph *g, int32_t *mode, int32_t *copy_threads);
/* Host half of the packed hand-off, usable on its own (no device involved): widens a corpus of n_walks rows of
 * walk_length little-endian 3-byte ids into int32; positions >= lens[w] become -1 (lens may be NULL: every row is
 * full).  threads <= 0 = the library's default copy-thread count. */
int gw_corpus_unpack24(const void *packed, const int32_t *lens, int64_t n_walks, int32_t walk_length, int32_t threads,
                       int32_t *out_walks);
/* simulate_walks' `random.shuffle(nodes)` (node2vec.py:51) at native speed ON THE INTERPRETER'S OWN GENERATOR: mt624 /
 * mt_index are the 624 state words and the position of CPython's `random.getstate()` (MT19937), updated in place; items[n]
 * is permuted exactly as `random.shuffle` (CPython 3.2+: j = _randbelow(i + 1) by getrandbits with rejection, i = n-1 .. 1)
 * would, and the state left behind is the one the Python loop would leave, so the reference's global-RNG "seed interface"
 * keeps its meaning.  Host only. */
int gw_py_random_shuffle(uint32_t *mt624, int32_t *mt_index, int64_t *items, int64_t n);
/* Optional: runs the walker's one-off preprocessing now (per-edge common-neighbour counts, the
 * scalable stand-in for preprocess_transition_probs' alias_edges, node2vec.py:99-108) instead of
 * lazily inside the first walk call; build_ms (may be NULL) receives its device time. */
int gw_graph_prepare_walks(gw_graph *g, double *build_ms);
/* The preprocessing's result, for inspection and tests: counts[nnz] = |N(u) & N(v)| of every directed CSR entry
 * (u -> v) -- what decides the masses of get_alias_edge's three weight classes (node2vec.py:69-76) -- and, when
 * reverse_index != NULL, the position of u inside the sorted row of v (-1 when the graph has rows >= 65536 entries
 * and the walker keeps no reverse index).  Undirected, unweighted, loop-free SIMPLE graphs only (GW_E_STATE otherwise). */
int gw_graph_common_counts(gw_graph *g, int32_t *counts, int32_t *reverse_index);
/* Replay: consumes the reference's own np.random.rand() stream (two fp64 draws per executed
 * step, node2vec.py:156-160) and start order; needs gw_alias_nodes + gw_alias_edges built with
 * the same p,q.  draw_offset[n_starts+1] may be NULL when no walk can hit a dead end
 * (then walk i starts at 2*(walk_length-1)*i). */
int gw_node2vec_walks_replay(gw_graph *g, int32_t walk_length, const int64_t *starts,
                             int64_t n_starts, const double *uniforms, int64_t n_uniforms,
                             const int64_t *draw_offset, int32_t *out_walks, int32_t *out_lens);
/* Byte model of the production (mixture) walker, measured by re-running the SAME walks (same
 * seed / walk ids) in a counting mode that stores no corpus: stats5 = {steps, random accesses
 * ({nbr,cnt,offset,degree} loads + search sectors), streamed row bytes, warp-cooperative
 * intersections, extra proposals}  (DESIGN.md §4). */
int gw_node2vec_walk_traffic_dev(gw_graph *g, double p, double q, int32_t walk_length,
                                 const int64_t *d_starts, int64_t n_starts, uint64_t seed,
                                 uint64_t walk_id_base, int64_t *stats5, void *stream);
/* Byte model of SURVEY.md §8(d) evaluated on a device-resident corpus: number of executed
 * steps and sum over steps of S(d_prev) (0 for first steps). */
int gw_walks_byte_model_dev(const gw_graph *g, const int32_t *d_walks, int64_t n_walks,
                            int32_t walk_length, int second_order, int64_t *steps,
                            int64_t *sum_search_sectors, void *stream);

/* ---- SimRank  (SingleRandomWalk.java, TopSim_singleSample.java, Print.java/FixedMaxPQ.java) ---- */
/* Per query: SAMPLE reverse walks of length 2*step, first-meeting accumulation with decay c,
 * then top-k (score descending, id ascending; the source itself excluded, zero scores padded
 * with id -1).  out_ids[nq*k] int32, out_scores[nq*k] fp64.  Philox keyed by
 * (seed, query_id_base + i). */
int gw_simrank_topk(gw_graph *g, const int64_t *queries, int64_t nq, double c, int32_t step,
                    int32_t sample, int32_t k, int32_t mode, uint64_t seed,
                    uint64_t query_id_base, int32_t *out_ids, double *out_scores);
int gw_simrank_topk_dev(gw_graph *g, const int64_t *d_queries, int64_t nq, double c, int32_t step,
                        int32_t sample, int32_t k, int32_t mode, uint64_t seed,
                        uint64_t query_id_base, int32_t *d_out_ids, double *d_out_scores,
                        void *stream);
/* Dense rows of the same estimator (getResult() of the reference): out[nq*n] fp64. */
int gw_simrank_rows(gw_graph *g, const int64_t *queries, int64_t nq, double c, int32_t step,
                    int32_t sample, int32_t mode, uint64_t seed, uint64_t query_id_base,
                    double *out_dense);
/* Replay mode of SingleRandomWalk (SingleRandomWalk.java:39-106 with structures/Graph.java:69-73): every
 * query is walked SEQUENTIALLY by one thread with java.util.Random itself (48-bit LCG, nextInt(bound)
 * with its rejection loop), starting from rng_state[i] (the scrambled 48-bit seed), and accumulated
 * in fp64 in the reference's operation order: out[nq*n] is what the Java code leaves in sim[v][*]
 * when its static Random is in that state, and rng_state[i] is updated to the state after the query
 * (feed it to the next query to replay a whole compute()).  Parity path, not the fast path. */
int gw_simrank_rows_javarng(gw_graph *g, const int64_t *queries, int64_t nq, double c, int32_t step,
                            int32_t sample, uint64_t *rng_state, double *out_dense);
/* Replay mode of the path-tree estimators: mode 0 = TopSim_singleSample.walk (TopSim_singleSample.java:62-203:
 * enumerate while weight >= degree, else ceil(weight) children drawn with java.util.Random), mode 1 =
 * TopSim_Enumerate.walk (TopSim_Enumerate.java:61-184: always enumerate, no random draw).  One thread per
 * query runs the reference's queue in its order with fp64 in its operation order; scores are x SAMPLE as in
 * the reference (:189).  max_paths bounds the queue of one level (GW_E_TOO_LARGE when exceeded). */
int gw_topsim_rows_javarng(gw_graph *g, const int64_t *queries, int64_t nq, double c, int32_t step,
                           int32_t sample, int32_t mode, int64_t max_paths, uint64_t *rng_state,
                           double *out_dense);
/* Replay mode of the bounded-cache estimators: mode 0 = SingleRandomWalk_M.walk (SingleRandomWalk_M.java:59-94),
 * mode 1 = TopSim_singleSample_M.walk (TopSim_singleSample_M.java:61-239).  Same walks / path tree and
 * java.util.Random stream as the two entry points above; each increment (/ SAMPLE in both classes) is cast to float
 * and put() into the query's FixedCacheMap (lxctools/FixedCacheMap.java:32-50: accumulate a present key and sink it,
 * else append and swim while not full, else replace the minimum when strictly greater) of `capacity` = topk * M
 * slots, 1..32767 (the reference stores heap slots as Short, :17).  Output is what getResult()[v] holds: per query
 * out_sizes[i] live entries, out_keys/out_vals[i*capacity + s] = heap slot s+1 (heap order, NOT sorted; iterate with
 * delMin as Print.printByOrder(FixedCacheMap[], ...) does, utils/Print.java:94-123).  max_paths: mode 1 only. */
int gw_simrank_cache_javarng(gw_graph *g, const int64_t *queries, int64_t nq, double c, int32_t step,
                             int32_t sample, int32_t mode, int32_t capacity, int64_t max_paths,
                             uint64_t *rng_state, int32_t *out_keys, float *out_vals, int32_t *out_sizes);
/* DoubleRandomWalk.samplePaths (simrank/DoubleRandomWalk.java:50-65): `sample` walks of `step` steps from each of
 * the nv vertices; out_paths is int32 [nv][sample][step] as the reference's paths[][][] (a dead end stores -1 and
 * leaves the later slots 0).  rng_state == NULL: Philox keyed by (seed, vertex), counter (sample index, step);
 * rng_state != NULL: replay -- vertex i is walked by one thread with java.util.Random from rng_state[i], which is
 * updated to the state after its last draw (chain them to replay a seeded JVM). */
int gw_double_walk_paths(gw_graph *g, const int64_t *vertices, int64_t nv, int32_t sample, int32_t step,
                         uint64_t seed, uint64_t *rng_state, int32_t *out_paths);
/* DoubleRandomWalk.getSim / computeSims (:67-91) over a path set [nv][sample][step]: out[nrows*nv] holds, for every
 * requested row r (an index into the path set) and every column w != r, the score of the pair (min, max) as the
 * reference evaluates it; the diagonal is 0.  exact_order != 0: fp64 adds in the reference's order (bit-exact);
 * 0: integer first-meeting counts combined once (production; same value up to fp64 rounding order). */
int gw_double_walk_sims(gw_graph *g, const int32_t *paths, int64_t nv, int32_t sample, int32_t step, double c,
                        const int64_t *rows, int64_t nrows, int32_t exact_order, double *out_dense);
/* TopSim_doubleSample.sample / TopSim_Dev.sample + computePath (simrank/TopSim_doubleSample.java:66-178,
 * simrank/TopSim_Dev.java:104-226): ns independent path-mass trees (sources may repeat), each started with `weight`
 * (SAMPLE) and `step` levels deep; out_mass is fp64 [ns][n][step+1] as the reference's paths[src][target][level]
 * (-1 = not reached; level 0 unused; a target reached by several paths of a level keeps the LAST one's weight).
 * rng_state == NULL: Philox keyed by (seed, call_id_base + i); else java.util.Random replay from rng_state[i]
 * (updated).  max_paths bounds one level of one tree (GW_E_TOO_LARGE when exceeded). */
int gw_topsim_mass(gw_graph *g, const int64_t *sources, int64_t ns, double weight, int32_t step,
                   int64_t max_paths, uint64_t seed, uint64_t call_id_base, uint64_t *rng_state,
                   double *out_mass);
/* getSim of both classes (TopSim_doubleSample.java:189-199, TopSim_Dev.java:233-244) for npairs (a, b) index pairs
 * into a mass set [ns][n][step+1]: sum over targets and levels of c^level * mass_a * mass_b where both are set.
 * exact_order != 0: the reference's loop and fp64 operation order (bit-exact); 0: warp-parallel reduction. */
int gw_topsim_mass_sims(gw_graph *g, const double *mass, int64_t ns, int32_t step, double c,
                        const int64_t *pair_a, const int64_t *pair_b, int64_t npairs, int32_t exact_order,
                        double *out);
/* Total walk steps executed by the last gw_simrank_* call on this graph. */
int gw_simrank_last_steps(const gw_graph *g, int64_t *steps);
/* Queries of the last gw_simrank_topk* call that were finished by the exact hash-table kernel
 * instead of the log-structured fast path (threshold at the single-hit level; results identical). */
int gw_simrank_last_slow_queries(const gw_graph *g, int64_t *count);
/* The argument checks every gw_simrank_topk* / gw_simrank_rows call starts with (graph kind, decay in (0,1), step in
 * 1..10, sample >= 1, k in 1..128, known mode), without touching the device: GW_OK or GW_E_INVALID + message. */
int gw_simrank_check_args(const gw_graph *g, double c, int32_t step, int32_t sample, int32_t k, int32_t mode);
/* Device-side status of the last gw_simrank_topk* call on this graph (synchronises the device): 0, or the code of an
 * accumulator / path-buffer overflow, in which case the results of that call are truncated.  The host-buffer entry
 * points report this themselves (GW_E_STATE); callers of the _dev entry point check here after their stream has
 * finished.  The next call restores the accumulators on the device either way. */
int gw_simrank_last_error(const gw_graph *g, int32_t *code);
/* Exact SimRank (simrank/SimRank.java:36-77): iters Jacobi sweeps of S <- c P S P^T, diag = 1,
 * diag zeroed at the end; returns the requested rows out[nrows*n].  O(n^2) device memory. */
int gw_simrank_exact(gw_graph *g, double c, int32_t iters, const int64_t *rows, int64_t nrows,
                     double *out_dense);

/* ---- skip-gram on the walk corpus  (node2vec/src/main.py:92-101 learn_embeddings = gensim 0.13.3 Word2Vec(sg=1, hs=0,
 * negative=5, sample=1e-3, alpha=0.025, min_alpha=0.0001)) ----
 * The consumer of the corpus, on the device: the 13-215 GB hand-off that bounds the walk API end to end never happens.
 * A model holds syn0 (the embeddings, initialised to (uniform - 0.5) / dimensions) and syn1neg (zeros) for n_words
 * words (the n vertices of a graph), the word counts of the vocabulary scan, gensim's subsampling thresholds and the negative-sampling table.
 * dimensions: 32, 64, 128 or 256.  Words are dense vertex indices; -1 pads a walk.  Training runs one warp per walk,
 * walks concurrently, rows updated without locks (Hogwild, as gensim's worker threads); the random streams are Philox
 * keyed by (seed, sentence id), so a run is reproducible in its draws though not in the order of racing updates. */
typedef struct gw_sgns gw_sgns;
int gw_sgns_create(int64_t n_words, int32_t dimensions, uint64_t seed, int32_t device, gw_sgns **out);
int gw_sgns_free(gw_sgns *m);
/* build_vocab's scan: adds the words of d_walks[n_walks*walk_length] (device) to the model's counts. */
int gw_sgns_count_dev(gw_sgns *m, const int32_t *d_walks, int64_t n_walks, int32_t walk_length, void *stream);
/* scale_vocab + make_cum_table from the counts scanned so far: subsampling thresholds for `sample` (0 = keep all) and
 * the table negatives are drawn from (proportional to count^0.75).  negative: draws per pair, 1..64. */
int gw_sgns_finalize_vocab(gw_sgns *m, double sample, int32_t negative);
/* train_batch_sg over d_walks: learning rate of walk s = alpha - (alpha - min_alpha) * (words_before + s*walk_length) /
 * total_words; sentence ids sentence_id_base + s key the random streams; subsample = 0 switches the draw off;
 * sequential != 0 runs ONE warp over the walks in order (deterministic: the mode the tests compare with the oracle). */
int gw_sgns_train_dev(gw_sgns *m, const int32_t *d_walks, int64_t n_walks, int32_t walk_length, int32_t window, double alpha,
                      double min_alpha, double words_before, double total_words, uint64_t sentence_id_base,
                      int32_t subsample, int32_t sequential, void *stream);
int gw_sgns_info(const gw_sgns *m, int64_t *n, int32_t *dimensions, double *total_words, int64_t *trained_pairs);
/* Copies to host (any pointer may be NULL): syn0[n*dimensions] (the embeddings), syn1neg[n*dimensions], counts[n]. */
int gw_sgns_vectors(const gw_sgns *m, float *out_syn0, float *out_syn1neg, int64_t *out_counts);
int gw_sgns_set_vectors(gw_sgns *m, const float *syn0, const float *syn1neg);
/* main.py:104-114 from simulate_walks to the embeddings in one call, the corpus never materialised: starts holds
 * num_walks shuffled start orders of n_starts vertices each (node2vec.py:51), pass w is regenerated from the seed
 * (walk ids w*n_starts + i) for the vocabulary scan and for each of the `iter` epochs.  out_vectors[n*dimensions],
 * out_counts[n] (may be NULL), out_seconds3 (may be NULL) = device seconds spent in {walks, vocabulary scan, training}. */
int gw_node2vec_embeddings(gw_graph *g, double p, double q, int32_t walk_length, int32_t num_walks, const int64_t *starts,
                           int64_t n_starts, int32_t dimensions, int32_t window, int32_t iter, int32_t negative, double sample,
                           double alpha, double min_alpha, uint64_t seed, float *out_vectors, int64_t *out_counts,
                           double *out_seconds3);

/* ---- multi-GPU: one process per GPU, graph replicated, units sharded, results gathered over NCCL ----
 * The path shards into independent units (node2vec.py:53-57: one walk per start node; SingleRandomWalk.java:39-45:
 * one row per query), so there is no data-path collective: rank r of nranks handles the contiguous slice
 * gw_shard_range(n, r, nranks) with GLOBAL walk / query ids in the RNG counters, and the only exchange is the
 * gather of the result blocks.  NCCL is loaded at run time (libnccl.so.2); single-GPU callers never need it.
 * Bootstrap like any NCCL program: rank 0 calls gw_comm_unique_id and hands the 128 bytes to the other ranks
 * (file, socket, JVM/Python launcher), then every rank calls gw_comm_init with its device. */
typedef struct gw_comm gw_comm;
int gw_comm_unique_id(void *id128);
int gw_comm_init(int32_t rank, int32_t nranks, const void *id128, int32_t device, gw_comm **out);
int gw_comm_info(const gw_comm *c, int32_t *rank, int32_t *nranks, int32_t *device);
int gw_comm_free(gw_comm *c);
int gw_shard_range(int64_t n, int32_t rank, int32_t nranks, int64_t *lo, int64_t *hi);
/* Device time of the last gw_*_sharded call on this communicator, split into this rank's own slice (kernels) and the
 * NCCL exchange of the result blocks (0 for gather = 0: nothing is exchanged).  Either pointer may be NULL. */
int gw_comm_last_times(const gw_comm *c, double *compute_ms, double *gather_ms);
/* Every rank passes the SAME starts[n_starts]; rank r walks its slice with walk ids = global positions.
 * gather == 1: all ranks receive the whole corpus out_walks[n_starts*walk_length] (+ out_lens) -- identical to one
 * gw_node2vec_walks call on one GPU; gather == 2: only rank 0 receives it (grouped ncclSend/ncclRecv; out_walks may be
 * NULL elsewhere); gather == 0 (corpora beyond one GPU / host, e.g. R-MAT-26's 215 GB): nothing is exchanged, only this
 * rank's rows of out_walks / out_lens are written, at their global offsets, through the same chunked hand-off pipeline
 * as gw_node2vec_walks.  A collective call: per-rank failures (a start node outside the graph in one slice, an
 * allocation) are agreed on over NCCL and returned by every rank. */
int gw_node2vec_walks_sharded(gw_graph *g, gw_comm *c, double p, double q, int32_t walk_length,
                              const int64_t *starts, int64_t n_starts, uint64_t seed, int32_t gather,
                              int32_t *out_walks, int32_t *out_lens);
/* Every rank passes the SAME queries[nq]; all ranks receive out_ids / out_scores [nq*k], identical to one
 * gw_simrank_topk call on one GPU. */
int gw_simrank_topk_sharded(gw_graph *g, gw_comm *c, const int64_t *queries, int64_t nq, double c_decay,
                            int32_t step, int32_t sample, int32_t k, int32_t mode, uint64_t seed,
                            int32_t *out_ids, double *out_scores);

#ifdef __cplusplus
}
#endif
#endif /* GRAPHWALK_H */
