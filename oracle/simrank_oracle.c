/* CPU oracle for the DeepSim/TopSim SimRank path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's Java (DeepSim/TopSimAll/src):
 *   structures/Graph.java:28-73          multigraph adjacency, randNeighbor
 *   simrank/SingleRandomWalk.java:28-106 MC single-walk estimator
 *   simrank/TopSim_singleSample.java:35-218  hybrid enumerate-or-sample path tree
 *   simrank/TopSim_Enumerate.java:61-184 full enumeration (deterministic expectation)
 *   simrank/SimRank.java:21-77           naive exact iteration
 *   lxctools/FixedMaxPQ.java:30-39,72-76 + Pair.java:78-80   top-k (java.util.PriorityQueue)
 *   lxctools/FixedCacheMap.java:26-113   bounded evict-the-minimum cache (the `_M` estimators)
 *   simrank/SingleRandomWalk_M.java:28-103, simrank/TopSim_singleSample_M.java:33-239
 *   simrank/DoubleRandomWalk.java:25-91  two independent walk sets per pair
 *   simrank/TopSim_doubleSample.java:66-200, simrank/TopSim_Dev.java:104-244   path-mass trees and their products
 *   java.util.Random (JDK, not in the repo): 48-bit LCG, nextInt(bound)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * leg may load this library; the product (libgraphwalk.so) never links or calls it.
 *
 * Parity status: the exact-SimRank routine is PINNED by the repo's only golden vector
 * (IsoMap_LE/data/0_333_5038_simrank_navie_top10.txt.sim.txt, C=0.8, 30 sweeps; see
 * tests/test_oracle_simrank.py).  The Monte-Carlo estimators are UNPINNED at the RNG
 * boundary: no JVM exists in the build container and the reference seeds nothing
 * (Graph.java:17), so they are anchored on (i) the pinned exact routine, whose STEP-sweep
 * result is their expectation, and (ii) the independent restatement of the formula in
 * simrank/random_test/RandomWalkTest.java:177-219.
 *
 * FixedCacheMap, the _M estimators, DoubleRandomWalk, TopSim_doubleSample and TopSim_Dev: UNPINNED at the RNG
 * boundary for the same reason; anchored on known answers traced by hand from the Java source
 * (tests/test_oracle_simrank.py: FixedCacheMap.main()'s put sequence, path-mass trees with the last-writer-wins
 * rule of computePath, getSim products) and on the dense classes they share their walks with.
 *
 * The adjacency arrives as CSR (row_ptr int64[V+1], col int32[nnz]) built by the Python
 * side in FILE ORDER per vertex (Graph.addEdge appends, duplicates kept).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------- java.util.Random ---------------- */
typedef struct { uint64_t seed; } jrand;
#define JR_MULT 0x5DEECE66DULL
#define JR_MASK ((1ULL << 48) - 1)

void jr_seed(jrand *r, int64_t s) { r->seed = ((uint64_t)s ^ JR_MULT) & JR_MASK; }
static int32_t jr_next(jrand *r, int bits) {
    r->seed = (r->seed * JR_MULT + 0xBULL) & JR_MASK;
    return (int32_t)((int64_t)r->seed >> (48 - bits));
}
int32_t jr_next_int(jrand *r, int32_t bound) {
    int32_t v = jr_next(r, 31);
    int32_t m = bound - 1;
    if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)v) >> 31);
    int32_t u = v;
    /* while (u - (r = u % bound) + m < 0) with int32 wrap-around */
    for (;;) {
        v = u % bound;
        int32_t t = (int32_t)((uint32_t)u - (uint32_t)v + (uint32_t)m);
        if (t >= 0) break;
        u = jr_next(r, 31);
    }
    return v;
}
/* test hook: fill out[n] with nextInt(bound) from seed */
void jr_fill(int64_t seed, int32_t bound, int32_t n, int32_t *out) {
    jrand r; jr_seed(&r, seed);
    for (int i = 0; i < n; i++) out[i] = jr_next_int(&r, bound);
}

/* ---------------- Graph.java ---------------- */
typedef struct { int64_t V; const int64_t *rp; const int32_t *col; } graph;
static inline int deg(const graph *g, int v) { return (int)(g->rp[v + 1] - g->rp[v]); }
static inline int rand_neighbor(const graph *g, jrand *r, int v) {   /* Graph.java:69-73 */
    int d = deg(g, v);
    if (d == 0) return -1;
    return g->col[g->rp[v] + jr_next_int(r, d)];
}

/* SingleRandomWalk.isFirstMeet (SingleRandomWalk.java:100-106), srcIndex = 0 */
static int is_first_meet(const int32_t *path, int dst) {
    int internal = dst / 2;
    for (int i = 0; i < internal; i++)
        if (path[i] == path[dst - i]) return 0;
    return 1;
}


/* ---------------- lxctools/FixedCacheMap.java ----------------
 * 1-based binary min-heap on float values + key -> heap index map.  The reference keeps the map in a
 * HashMap<Integer, Short> (:17; "(short) N" :40 wraps above 32767 slots -- capacities here stay below);
 * the restatement keeps it as a dense pos[V] array (0 = absent), which has the same contents. */
typedef struct { int nmax, n; int32_t *keys; float *vals; int32_t *pos; } fcm;
static int fcm_greater(const fcm *m, int a, int b) { return m->vals[a] > m->vals[b]; }     /* :80-82 */
static void fcm_exch(fcm *m, int a, int b) {                                             /* :84-96 */
    m->pos[m->keys[a]] = b;
    m->pos[m->keys[b]] = a;
    int32_t tk = m->keys[a]; m->keys[a] = m->keys[b]; m->keys[b] = tk;
    float tv = m->vals[a]; m->vals[a] = m->vals[b]; m->vals[b] = tv;
}
static void fcm_sink(fcm *m, int i) {                                                    /* :61-69 */
    while (2 * i <= m->n) {
        int j = 2 * i;
        if (j < m->n && fcm_greater(m, j, j + 1)) j++;
        if (!fcm_greater(m, i, j)) break;
        fcm_exch(m, i, j);
        i = j;
    }
}
static void fcm_swim(fcm *m, int i) {                                                    /* :73-78 */
    while (i > 1 && fcm_greater(m, i / 2, i)) { fcm_exch(m, i, i / 2); i = i / 2; }
}
static void fcm_put(fcm *m, int32_t key, float value) {                                  /* :32-50 */
    int idx = m->pos[key];
    if (idx != 0) {
        m->vals[idx] += value;
        fcm_sink(m, idx);
    } else if (m->n < m->nmax) {
        m->n++;
        m->keys[m->n] = key;
        m->vals[m->n] = value;
        m->pos[key] = m->n;
        fcm_swim(m, m->n);
    } else if (value > m->vals[1]) {
        m->pos[m->keys[1]] = 0;
        m->keys[1] = key;
        m->vals[1] = value;
        m->pos[key] = 1;
        fcm_sink(m, 1);
    }
}
/* delMin (:102-108) until empty: the iteration order of `for (Pair p : cacheMap)` -- ascending by value,
 * ties in heap order; destructive.  heap arrays are 1-based with n live slots; out arrays get n entries. */
void or_fcm_drain(int32_t n, int32_t *keys /* n+1 */, float *vals /* n+1 */, int32_t *out_keys, float *out_vals) {
    int32_t maxkey = 0;
    for (int i = 1; i <= n; i++) if (keys[i] > maxkey) maxkey = keys[i];
    int32_t *pos = (int32_t *)calloc((size_t)maxkey + 1, sizeof(int32_t));
    for (int i = 1; i <= n; i++) pos[keys[i]] = i;
    fcm m = {n, n, keys, vals, pos};
    int k = 0;
    while (m.n > 0) {
        out_keys[k] = m.keys[1]; out_vals[k] = m.vals[1]; k++;
        m.pos[m.keys[1]] = 0;
        fcm_exch(&m, 1, m.n--);
        fcm_sink(&m, 1);
    }
    free(pos);
}
/* test hook: a put() sequence into one cache; returns the live size, heap arrays (1-based) in hk/hv */
int32_t or_fcm_puts(int32_t nmax, int32_t nput, const int32_t *pk, const float *pv, int32_t key_space,
                    int32_t *hk /* nmax+1 */, float *hv /* nmax+1 */) {
    int32_t *pos = (int32_t *)calloc((size_t)key_space, sizeof(int32_t));
    fcm m = {nmax, 0, hk, hv, pos};
    for (int i = 0; i < nput; i++) fcm_put(&m, pk[i], pv[i]);
    free(pos);
    return m.n;
}

/* where a path's contribution goes: the dense row (double[][] sim) or the vertex's FixedCacheMap */
typedef struct { double *row; fcm *cache; } sink_t;

/* SingleRandomWalk.walk + computePathSim for ONE source row (SingleRandomWalk.java:53-92).
 * row[V] is accumulated (caller zeroes); returns the number of walk steps executed.
 * seed_io: java.util.Random state carried across calls (the reference shares one static RNG). */
static int64_t single_random_walk(const graph *gp, int32_t v, int32_t sample, int32_t step, double C,
                                  uint64_t *seed_io, sink_t sk) {
    graph g = *gp;
    double *row = sk.row;
    jrand r; r.seed = *seed_io;
    int max_step = 2 * step;
    double *cache = (double *)malloc(sizeof(double) * (step + 1));
    for (int i = 1; i <= step; i++) cache[i] = pow(C, i);          /* :34-36 */
    int32_t *path = (int32_t *)malloc(sizeof(int32_t) * (max_step + 1));
    int64_t steps = 0;
    for (int s = 0; s < sample; s++) {
        int path_len = 0;
        for (int i = 0; i <= max_step; i++) path[i] = -1;
        path[0] = v;
        int cur = v;
        while (path_len < max_step) {                              /* :64-68 */
            cur = rand_neighbor(&g, &r, cur);
            if (cur == -1) break;
            path[++path_len] = cur;
            steps++;
        }
        if (path_len == 0) continue;                               /* :82 */
        for (int i = 1; i <= step && 2 * i <= path_len; i++) {     /* :84-91 */
            int inter = path[i], target = path[2 * i];
            if (target == v) continue;
            if (is_first_meet(path, 2 * i)) {
                double incre = cache[i] * deg(&g, inter) / deg(&g, target) / sample;
                if (sk.cache) fcm_put(sk.cache, target, (float)incre);     /* SingleRandomWalk_M.java:90-91 */
                else row[target] += incre;                                 /* SingleRandomWalk.java:89 */
            }
        }
    }
    if (row) row[v] = 0;                                           /* :43 (the _M class has no such line) */
    *seed_io = r.seed;
    free(cache); free(path);
    return steps;
}
int64_t or_single_random_walk_row(int64_t V, const int64_t *rp, const int32_t *col, int32_t v,
                                  int32_t sample, int32_t step, double C, uint64_t *seed_io,
                                  double *row) {
    graph g = {V, rp, col};
    sink_t sk = {row, NULL};
    return single_random_walk(&g, v, sample, step, C, seed_io, sk);
}
/* SingleRandomWalk_M.walk(v) (SingleRandomWalk_M.java:59-94; STEP is the class constant 5 there, a parameter
 * here): same walks, increments cast to float and put() into the vertex's FixedCacheMap of `capacity` slots.
 * hk/hv: 1-based heap arrays [capacity+1]; *size_out = live entries. */
int64_t or_single_random_walk_cache(int64_t V, const int64_t *rp, const int32_t *col, int32_t v,
                                    int32_t sample, int32_t step, double C, int32_t capacity,
                                    uint64_t *seed_io, int32_t *hk, float *hv, int32_t *size_out) {
    graph g = {V, rp, col};
    int32_t *pos = (int32_t *)calloc((size_t)V, sizeof(int32_t));
    fcm m = {capacity, 0, hk, hv, pos};
    sink_t sk = {NULL, &m};
    int64_t steps = single_random_walk(&g, v, sample, step, C, seed_io, sk);
    *size_out = m.n;
    free(pos);
    return steps;
}

/* ---------- TopSim_singleSample / TopSim_Enumerate (level-synchronous path tree) ---------- */
/* Only the weight of the LAST vertex is ever read: computePathSim at pathLen = 2i runs i = start = TopSim
 * alone (:80-83,:180), so path[2i].sample is the current weight.  2*STEP+1 <= 22. */
typedef struct { int32_t cur[22]; double w; } wpath;

static int wp_first_meet(const wpath *p, int dst) {
    int internal = dst / 2;
    for (int i = 0; i < internal; i++)
        if (p->cur[i] == p->cur[dst - i]) return 0;
    return 1;
}

/* mode 0 = TopSim_singleSample.walk (:62-158): enumerate while weight >= degree, else ceil(weight)
 *          random children;  mode 1 = TopSim_Enumerate.walk (:61-130): always enumerate.
 * Scores are UNNORMALISED (x SAMPLE), TopSim_singleSample.java:189.
 * Returns number of child paths created (work measure), or -1 on overflow of max_paths. */
static int64_t topsim(const graph *gp, int32_t v, int32_t sample, int32_t step, double C, int32_t mode,
                      int64_t max_paths, uint64_t *seed_io, sink_t sk) {
    graph g = *gp;
    double *row = sk.row;
    jrand r; r.seed = *seed_io;
    int max_step = 2 * step;
    if (max_step + 1 > 22) return -2;
    double cache[16];
    for (int i = 1; i <= step; i++) cache[i] = pow(C, i);
    wpath *q0 = (wpath *)malloc(sizeof(wpath) * (size_t)max_paths);
    wpath *q1 = (wpath *)malloc(sizeof(wpath) * (size_t)max_paths);
    int64_t n0 = 0, n1 = 0, made = 0;
    for (int i = 0; i <= max_step; i++) q0[0].cur[i] = -1;
    q0[0].cur[0] = v; q0[0].w = (double)sample;
    n0 = 1;
    int path_len = 0, topsim = 1, overflow = 0;
    for (;;) {
        /* computePathSim(queue, pathLen, TopSim) when pathLen/2 == TopSim (:80-83) and once
         * more after the loop with start = TopSim (:157). */
        int last = (path_len >= max_step);
        if (last || path_len / 2 == topsim) {
            int start = topsim;
            if (path_len != 0) {
                for (int64_t k = 0; k < n0; k++) {
                    const wpath *p = &q0[k];
                    for (int i = start; i <= step && 2 * i <= path_len; i++) {  /* :180-192 */
                        int inter = p->cur[i], target = p->cur[2 * i];
                        if (target == v) continue;
                        if (target == -1) continue;
                        if (wp_first_meet(p, 2 * i)) {
                            double x = p->w * cache[i] * (double)deg(&g, inter) / (double)deg(&g, target);
                            if (sk.cache) fcm_put(sk.cache, target, (float)(x / sample));   /* TopSim_singleSample_M.java:224-225 */
                            else row[target] += x;                                          /* TopSim_singleSample.java:189 */
                        }
                    }
                }
            }
            if (!last) topsim++;
        }
        if (last) break;
        n1 = 0;
        for (int64_t k = 0; k < n0 && !overflow; k++) {
            const wpath *p = &q0[k];
            int c = p->cur[path_len];
            double wt = p->w;
            int d = deg(&g, c);
            if (d != 0 && (mode == 1 || wt >= d)) {                /* :99-125 */
                double ns = wt / (double)d;
                if (n1 + d > max_paths) { overflow = 1; break; }
                for (int j = 0; j < d; j++) {
                    wpath *c1 = &q1[n1++];
                    *c1 = *p;
                    c1->cur[path_len + 1] = g.col[g.rp[c] + j];
                    c1->w = ns;
                    made++;
                }
            } else if (mode == 0) {                                /* :126-149 */
                int number = ((double)(int)wt == wt) ? (int)wt : (int)wt + 1;
                for (int j = 0; j < number; j++) {
                    int nb = rand_neighbor(&g, &r, c);
                    if (nb == -1) break;
                    if (n1 + 1 > max_paths) { overflow = 1; break; }
                    wpath *c1 = &q1[n1++];
                    *c1 = *p;
                    c1->cur[path_len + 1] = nb;
                    c1->w = wt / (double)number;
                    made++;
                }
            }
        }
        if (overflow) break;
        wpath *t = q0; q0 = q1; q1 = t;
        n0 = n1;
        path_len++;
    }
    if (row) row[v] = 0;                                           /* compute(): sim[i][i] = 0 (not in the _M class) */
    *seed_io = r.seed;
    free(q0); free(q1);
    return overflow ? -1 : made;
}
int64_t or_topsim_row(int64_t V, const int64_t *rp, const int32_t *col, int32_t v, int32_t sample,
                      int32_t step, double C, int32_t mode, int64_t max_paths, uint64_t *seed_io,
                      double *row) {
    graph g = {V, rp, col};
    sink_t sk = {row, NULL};
    return topsim(&g, v, sample, step, C, mode, max_paths, seed_io, sk);
}
/* TopSim_singleSample_M.walk(v) (TopSim_singleSample_M.java:61-176, computePathSim :202-239): the same path
 * tree, increments / SAMPLE cast to float and put() into the vertex's FixedCacheMap. */
int64_t or_topsim_cache(int64_t V, const int64_t *rp, const int32_t *col, int32_t v, int32_t sample,
                        int32_t step, double C, int32_t capacity, int64_t max_paths, uint64_t *seed_io,
                        int32_t *hk, float *hv, int32_t *size_out) {
    graph g = {V, rp, col};
    int32_t *pos = (int32_t *)calloc((size_t)V, sizeof(int32_t));
    fcm m = {capacity, 0, hk, hv, pos};
    sink_t sk = {NULL, &m};
    int64_t made = topsim(&g, v, sample, step, C, 0, max_paths, seed_io, sk);
    *size_out = m.n;
    free(pos);
    return made;
}

/* ---------------- simrank/DoubleRandomWalk.java ----------------
 * samplePaths (:50-65): for every vertex in order, SAMPLE walks of STEP steps, paths[v][i][step] (a dead end
 * stores -1 and stops; later slots stay 0 as Java's int[] initialises them -- getSim stops at the -1).
 * paths: int32 [V][sample][step], zero-initialised by the caller. */
void or_double_walk_paths(int64_t V, const int64_t *rp, const int32_t *col, int32_t sample, int32_t step,
                          uint64_t *seed_io, int32_t *paths) {
    graph g = {V, rp, col};
    jrand r; r.seed = *seed_io;
    for (int64_t src = 0; src < V; src++)
        for (int i = 0; i < sample; i++) {
            int cur = (int)src;
            for (int s = 0; s < step; s++) {
                cur = rand_neighbor(&g, &r, cur);
                paths[(src * sample + i) * step + s] = cur;
                if (cur == -1) break;
            }
        }
    *seed_io = r.seed;
}
/* getSim(v, w) (:77-91): all SAMPLE^2 path pairs, first common position counts cache[step+1]; / SAMPLE^2
 * (int product, as the reference writes it). */
double or_double_walk_sim(const int32_t *paths, int32_t sample, int32_t step, double C, int64_t v, int64_t w) {
    double cache[16];
    for (int i = 0; i <= step; i++) cache[i] = pow(C, i);          /* :33-35 */
    const int32_t *pv = paths + v * sample * step, *pw = paths + w * sample * step;
    double result = 0;
    for (int i = 0; i < sample; i++)
        for (int j = 0; j < sample; j++)
            for (int s = 0; s < step && pv[i * step + s] != -1 && pw[j * step + s] != -1; s++)
                if (pv[i * step + s] == pw[j * step + s]) { result += cache[s + 1]; break; }
    return result / (sample * sample);
}
/* computeSims (:67-75): upper triangle, mirrored; diagonal stays 0.  sim: V*V out. */
void or_double_walk_matrix(int64_t V, const int32_t *paths, int32_t sample, int32_t step, double C, double *sim) {
    memset(sim, 0, sizeof(double) * (size_t)V * V);
    for (int64_t i = 0; i < V; i++)
        for (int64_t j = i + 1; j < V; j++) {
            double x = or_double_walk_sim(paths, sample, step, C, i, j);
            sim[i * V + j] = x; sim[j * V + i] = x;
        }
}

/* ---------------- TopSim_doubleSample.sample / TopSim_Dev.sample (same code: TopSim_doubleSample.java:66-151,
 * TopSim_Dev.java:104-199) + computePath (:153-178 / :200-226) ----------------
 * The enumerate-or-sample tree of TopSim_singleSample, STEP levels deep, started with weight `weight0`; after each
 * level l = 1..STEP computePath stores mass[target][l] = weight of the path standing on `target` -- an OVERWRITE in
 * queue order (the last path wins), skipped when target == source.  Only the last vertex and the weight of a path
 * are ever read, so the queue holds (cur, w) pairs.  mass: V*(step+1) doubles, -1 = unset (caller fills).
 * Returns paths created, -1 when a level exceeds max_paths. */
int64_t or_mass_tree(int64_t V, const int64_t *rp, const int32_t *col, int32_t src, double weight0, int32_t step,
                     int64_t max_paths, uint64_t *seed_io, double *mass) {
    graph g = {V, rp, col};
    jrand r; r.seed = *seed_io;
    int32_t *c0 = (int32_t *)malloc(sizeof(int32_t) * (size_t)max_paths), *c1 = (int32_t *)malloc(sizeof(int32_t) * (size_t)max_paths);
    double *w0 = (double *)malloc(sizeof(double) * (size_t)max_paths), *w1 = (double *)malloc(sizeof(double) * (size_t)max_paths);
    int64_t n0 = 1, made = 0;
    int overflow = 0;
    c0[0] = src; w0[0] = weight0;
    for (int path_len = 0; path_len < step && !overflow; path_len++) {
        int64_t n1 = 0;
        for (int64_t k = 0; k < n0 && !overflow; k++) {
            int c = c0[k];
            double wt = w0[k];
            int d = deg(&g, c);
            if (d != 0 && wt >= d) {
                double ns = wt / (double)d;
                if (n1 + d > max_paths) { overflow = 1; break; }
                for (int j = 0; j < d; j++) { c1[n1] = g.col[g.rp[c] + j]; w1[n1] = ns; n1++; made++; }
            } else {
                int number = ((double)(int)wt == wt) ? (int)wt : (int)wt + 1;
                for (int j = 0; j < number; j++) {
                    int nb = rand_neighbor(&g, &r, c);
                    if (nb == -1) break;
                    if (n1 + 1 > max_paths) { overflow = 1; break; }
                    c1[n1] = nb; w1[n1] = wt / (double)number; n1++; made++;
                }
            }
        }
        if (overflow) break;
        int32_t *tc = c0; c0 = c1; c1 = tc;
        double *tw = w0; w0 = w1; w1 = tw;
        n0 = n1;
        int level = path_len + 1;                                  /* computePath(queue, level, level) */
        for (int64_t k = 0; k < n0; k++) {
            if (c0[k] == src) continue;
            mass[(int64_t)c0[k] * (step + 1) + level] = w0[k];
        }
    }
    *seed_io = r.seed;
    free(c0); free(c1); free(w0); free(w1);
    return overflow ? -1 : made;
}
/* getSim (TopSim_doubleSample.java:189-199, TopSim_Dev.java:233-244): sum over targets i and levels of
 * cache[level] * a * b where both masses are set (>= 0), in that loop order. */
double or_mass_sim(const double *ma, const double *mb, int64_t V, int32_t step, double C) {
    double cache[16];
    for (int i = 0; i <= step; i++) cache[i] = pow(C, i);
    double result = 0;
    for (int64_t i = 0; i < V; i++)
        for (int s = 1; s <= step; s++) {
            double a = ma[i * (step + 1) + s], b = mb[i * (step + 1) + s];
            if (a >= 0 && b >= 0) result += cache[s] * a * b;
        }
    return result;
}

/* ---------------- SimRank.java:36-77 (naive exact, Jacobi sweeps) ---------------- */
void or_simrank_exact(int64_t V, const int64_t *rp, const int32_t *col, double C, int32_t iters,
                      double *sim /* V*V, out */) {
    graph g = {V, rp, col};
    double *tmp = (double *)calloc((size_t)V * V, sizeof(double));
    memset(sim, 0, sizeof(double) * (size_t)V * V);
    for (int64_t i = 0; i < V; i++) { sim[i * V + i] = 1.0; tmp[i * V + i] = 1.0; }
    for (int r = 0; r < iters; r++) {
        for (int64_t i = 0; i < V; i++)
            for (int64_t j = i + 1; j < V; j++) {
                double res = 0;                                    /* sim(v,w) :67-77 */
                int di = deg(&g, (int)i), dj = deg(&g, (int)j);
                if (di != 0 && dj != 0) {
                    for (int64_t a = rp[i]; a < rp[i + 1]; a++) {
                        const double *srow = sim + (int64_t)col[a] * V;
                        for (int64_t b = rp[j]; b < rp[j + 1]; b++) res += srow[col[b]];
                    }
                    res = C * res / (di * dj);
                }
                tmp[i * V + j] = res;
                tmp[j * V + i] = res;
            }
        memcpy(sim, tmp, sizeof(double) * (size_t)V * V);
    }
    for (int64_t i = 0; i < V; i++) sim[i * V + i] = 0;            /* postProcess :62-65 */
    free(tmp);
}

/* ------------- FixedMaxPQ over one dense row (Print.java:31-37, FixedMaxPQ.java) ------------- */
typedef struct { int32_t key; double val; } pr;
static inline int pr_cmp(const pr *a, const pr *b) {               /* Double.compareTo, no NaN */
    return (a->val < b->val) ? -1 : (a->val > b->val) ? 1 : 0;
}
static void sift_up(pr *q, int k, pr x) {
    while (k > 0) {
        int parent = (k - 1) >> 1;
        if (pr_cmp(&x, &q[parent]) >= 0) break;
        q[k] = q[parent];
        k = parent;
    }
    q[k] = x;
}
static void sift_down(pr *q, int size, int k, pr x) {
    int half = size >> 1;
    while (k < half) {
        int child = 2 * k + 1, right = child + 1;
        if (right < size && pr_cmp(&q[child], &q[right]) > 0) child = right;
        if (pr_cmp(&x, &q[child]) <= 0) break;
        q[k] = q[child];
        k = child;
    }
    q[k] = x;
}
/* returns number of elements written (min(k, n)); descending, ties in heap-array order
 * (Collections.sort is stable). */
int32_t or_fixedmaxpq_topk_min(const double *row, int64_t n, int32_t k, double min_value, int32_t *out_ids,
                               double *out_vals);
int32_t or_fixedmaxpq_topk(const double *row, int64_t n, int32_t k, int32_t *out_ids,
                           double *out_vals) {
    return or_fixedmaxpq_topk_min(row, n, k, -INFINITY, out_ids, out_vals);
}
/* min_value: only entries >= min_value are offered (TopSim_Dev.java:76-83 filters candidate >= MIN) */
int32_t or_fixedmaxpq_topk_min(const double *row, int64_t n, int32_t k, double min_value, int32_t *out_ids,
                               double *out_vals) {
    pr *q = (pr *)malloc(sizeof(pr) * (size_t)(k > 0 ? k : 1));
    int size = 0;
    for (int64_t i = 0; i < n; i++) {
        if (!(row[i] >= min_value)) continue;
        pr e = {(int32_t)i, row[i]};
        if (size < k) { sift_up(q, size, e); size++; }            /* pq.offer */
        else if (k > 0 && pr_cmp(&q[0], &e) < 0) {                 /* peek().compareTo(e) < 0 */
            size--;                                                /* poll */
            pr x = q[size];
            if (size != 0) sift_down(q, size, 0, x);
            sift_up(q, size, e); size++;                           /* offer */
        }
    }
    /* stable insertion sort, descending */
    for (int i = 1; i < size; i++) {
        pr x = q[i]; int j = i - 1;
        while (j >= 0 && pr_cmp(&q[j], &x) < 0) { q[j + 1] = q[j]; j--; }
        q[j + 1] = x;
    }
    for (int i = 0; i < size; i++) { out_ids[i] = q[i].key; out_vals[i] = q[i].val; }
    free(q);
    return size;
}

/* ------------- free-running first-order/second-order node2vec CPU port (baseline) -------------
 * Same on-the-fly law as node2vec.py:61-81 for unweighted graphs: per step build nothing,
 * draw by inversion over the unnormalised weights of sorted N(cur).  Used only as a C-speed
 * CPU baseline beside the python port; xorshift RNG (statistical use only). */
static inline uint64_t xs64(uint64_t *s) { uint64_t x = *s; x ^= x << 13; x ^= x >> 7; x ^= x << 17; return *s = x; }
static int has_edge(const graph *g, int a, int b) {
    int64_t lo = g->rp[a], hi = g->rp[a + 1];
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (g->col[mid] < b) lo = mid + 1; else hi = mid; }
    return lo < g->rp[a + 1] && g->col[lo] == b;
}
int64_t or_node2vec_walks(int64_t V, const int64_t *rp, const int32_t *col, double p, double q,
                          int32_t walk_length, const int64_t *starts, int64_t n_starts,
                          uint64_t seed, int32_t *out /* n_starts*walk_length */) {
    graph g = {V, rp, col};
    uint64_t s = seed * 0x9E3779B97F4A7C15ULL + 1;
    int64_t steps = 0;
    for (int64_t w = 0; w < n_starts; w++) {
        int32_t *walk = out + w * walk_length;
        for (int i = 0; i < walk_length; i++) walk[i] = -1;
        int cur = (int)starts[w], prev = -1;
        walk[0] = cur;
        for (int l = 1; l < walk_length; l++) {
            int d = deg(&g, cur);
            if (d == 0) break;
            int nxt;
            if (prev < 0 || (p == 1.0 && q == 1.0)) {
                nxt = col[rp[cur] + (int64_t)((xs64(&s) >> 11) * (1.0 / 9007199254740992.0) * d)];
            } else {
                double tot = 0;
                for (int64_t e = rp[cur]; e < rp[cur + 1]; e++) {
                    int x = col[e];
                    tot += (x == prev) ? 1.0 / p : has_edge(&g, x, prev) ? 1.0 : 1.0 / q;
                }
                double u = (xs64(&s) >> 11) * (1.0 / 9007199254740992.0) * tot, acc = 0;
                nxt = col[rp[cur + 1] - 1];
                for (int64_t e = rp[cur]; e < rp[cur + 1]; e++) {
                    int x = col[e];
                    acc += (x == prev) ? 1.0 / p : has_edge(&g, x, prev) ? 1.0 : 1.0 / q;
                    if (u < acc) { nxt = x; break; }
                }
            }
            walk[l] = nxt; prev = cur; cur = nxt; steps++;
        }
    }
    return steps;
}

/* test hook: raw next(32) values (java.util.Random.nextInt()) */
void jr_fill32(int64_t seed, int32_t n, int32_t *out) {
    jrand r; jr_seed(&r, seed);
    for (int i = 0; i < n; i++) out[i] = jr_next(&r, 32);
}
