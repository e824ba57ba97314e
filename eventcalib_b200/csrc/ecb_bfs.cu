// k_bfs_order — member order of a DBSCAN cluster exactly as the reference produces it, and the std::nth_element median.
//
// DBSCAN::Run / expandCluster (dbscan/include/dbscan.h:143-162,229-259) append a cluster's members in the order a FIFO
// pops them; every popped core point pushes its not-yet-pushed neighbours in the order kd_nearest_range lists them =
// the REVERSE of find_nearest's visit order (node, near subtree, far subtree if |dx| < eps; head insertion,
// dbscan/src/kdtree.cpp:148-179,469-486).  Pop order = order of first push, so
//      level 0 = [seed];   level k+1 = members first reached from level k, ordered by
//                          (position of the first level-k parent u in the member list, position in u's result list).
// One warp per cluster: every lane walks the emulated kd-tree (exported by k_cluster) for one parent u with the
// reference's own pruning rule — stackless, through parent links — and claims the members it meets with a 64-bit
// atomicMin of (level, position of u, reversed visit counter); the claimed frontier is then ranked by key.
// Non-members a query meets (noise / border points, members of earlier clusters) never change the relative order.
//
// Consumers: (a) kept clusters whose median norm is tied — std::nth_element (restated in ecb_nth_element.h) then picks
// the same pixel as CirclesEventFrame.cpp:140-147; (b) ecb_dbscan_run_ordered: `Clusters` as ordered lists.
#include "ecb_cluster.cuh"
#include "ecb_nth_element.h"

namespace {

constexpr int BFS_THREADS = 128;
constexpr int BFS_WARPS = BFS_THREADS / 32;

// ---- tree policies -------------------------------------------------------------------------------------------------------
// PixTree: integer pixels, one 16-byte node {pixel, left, right, parent} (exported by k_cluster), 2-D.
// GenTree: arbitrary double coordinates of dimension <= ECB_GH_MAXD (the general grid-hash path, ecb_gridhash.cu), node
//          {left, right, parent, depth << 1 | side}; the split axis of a node is depth % dim like insert_rec's
//          new_dir = (dir + 1) % dim (kdtree.cpp:127).
struct PixTree {
    static constexpr int MAXD = 2;
    typedef int CT;  // pixel coordinates: the reference's double arithmetic is exact on them, so integers decide identically
    const uint32_t *pix;
    const uint4 *nodes;
    int eps2i, epsc;  // d^2 <= eps^2  <=>  d^2 <= floor(eps^2);  |dx| < eps  <=>  |dx| < ceil(eps)   (integer d^2, dx)
    struct Node {
        uint32_t l, r, par;
        int c[2];
    };
    __device__ __forceinline__ void set_eps(double eps) {
        eps2i = (int) floor(eps * eps);
        epsc = (int) ceil(eps);
    }
    __device__ __forceinline__ void point(uint32_t u, int *q) const {
        const uint32_t p = pix[u];
        q[0] = (int) ECB_PIX_X(p);
        q[1] = (int) ECB_PIX_Y(p);
    }
    __device__ __forceinline__ Node fetch(uint32_t n) const {
        const uint4 nd = __ldg(nodes + n);
        Node o;
        o.l = nd.y;
        o.r = nd.z;
        o.par = nd.w;
        o.c[0] = (int) ECB_PIX_X(nd.x);
        o.c[1] = (int) ECB_PIX_Y(nd.x);
        return o;
    }
    __device__ __forceinline__ bool within(const Node &nd, const int *q) const {
        const int ex = nd.c[0] - q[0], ey = nd.c[1] - q[1];
        return ex * ex + ey * ey <= eps2i;
    }
    __device__ __forceinline__ bool far_ok(int dx) const { return abs(dx) < epsc; }
    __device__ __forceinline__ int next_dir(int dir) const { return dir ^ 1; }
    __device__ __forceinline__ int prev_dir(int dir) const { return dir ^ 1; }
};

struct GenTree {
    static constexpr int MAXD = ECB_GH_MAXD;
    typedef double CT;
    const double *P;     // [n][dim], indexed from the problem's first point
    const uint4 *nodes;  // same indexing
    int d;
    double eps, eps2;
    struct Node {
        uint32_t l, r, par;
        double c[ECB_GH_MAXD];
    };
    __device__ __forceinline__ void set_eps(double e) {
        eps = e;
        eps2 = e * e;  // SQ(range), kdtree.cpp:159
    }
    __device__ __forceinline__ bool within(const Node &nd, const double *q) const {
        double d2 = 0.0;  // dist_sq += SQ(node->pos[i] - pos[i]) in axis order, no contraction (kdtree.cpp:155-158)
#pragma unroll
        for (int k = 0; k < MAXD; ++k)
            if (k < d) {
                const double e = __dsub_rn(nd.c[k], q[k]);
                d2 = __dadd_rn(d2, __dmul_rn(e, e));
            }
        return d2 <= eps2;
    }
    __device__ __forceinline__ bool far_ok(double dx) const { return fabs(dx) < eps; }
    __device__ __forceinline__ void point(uint32_t u, double *q) const {
#pragma unroll
        for (int k = 0; k < MAXD; ++k) q[k] = k < d ? P[(size_t) u * d + k] : 0.0;
    }
    __device__ __forceinline__ Node fetch(uint32_t n) const {
        const uint4 nd = __ldg(nodes + n);
        Node o;
        o.l = nd.x;
        o.r = nd.y;
        o.par = nd.z;
#pragma unroll
        for (int k = 0; k < MAXD; ++k) o.c[k] = k < d ? P[(size_t) n * d + k] : 0.0;
        return o;
    }
    __device__ __forceinline__ int next_dir(int dir) const { return dir + 1 == d ? 0 : dir + 1; }
    __device__ __forceinline__ int prev_dir(int dir) const { return dir == 0 ? d - 1 : dir - 1; }
};

template <int MAXD, typename CT>
__device__ __forceinline__ CT pick(const CT *v, int dir) {
    CT r = v[0];
#pragma unroll
    for (int k = 1; k < MAXD; ++k) r = dir == k ? v[k] : r;
    return r;
}

// claim keys: (level, position of the parent in the member list, reversed visit counter), 21 bits each
constexpr int KEY_B = 21;
constexpr unsigned long long KEY_M = (1ull << KEY_B) - 1ull;
constexpr int RANK_MAX = 192;  // frontiers up to this size are ranked by key; larger ones are placed by re-walking

// find_nearest(root, q, eps) of the reference (kdtree.cpp:148-179) without a stack: state = how the current node was
// entered (0: from the parent, 1: back from the near child, 2: back from the far child).  `visit(node, t)` is called for
// every node whose distance test passes (t = running visit counter).
template <class Tree, class F>
__device__ __forceinline__ void kd_walk(const Tree &tr, const typename Tree::CT *q, uint32_t root, F &&visit) {
    typedef typename Tree::CT CT;
    uint32_t node = root;
    int dir = 0, from = 0;
    uint32_t t = 0;
    typename Tree::Node nd = tr.fetch(node);
    for (;;) {
        const CT dx = pick<Tree::MAXD, CT>(q, dir) - pick<Tree::MAXD, CT>(nd.c, dir);
        if (from == 0) {
            if (tr.within(nd, q)) visit(node, t);
            ++t;
            const uint32_t nearc = dx <= (CT) 0 ? nd.l : nd.r;
            if (nearc != ECB_NONE) {
                node = nearc;
                nd = tr.fetch(node);
                dir = tr.next_dir(dir);
                continue;
            }
            from = 1;
        }
        if (from == 1) {
            const uint32_t farc = dx <= (CT) 0 ? nd.r : nd.l;
            if (tr.far_ok(dx) && farc != ECB_NONE) {
                node = farc;
                nd = tr.fetch(node);
                dir = tr.next_dir(dir);
                from = 0;
                continue;
            }
        }
        // this subtree is done: climb, and find out which child we are coming back from
        if (node == root) break;
        const uint32_t child = node;
        node = nd.par;
        nd = tr.fetch(node);
        dir = tr.prev_dir(dir);
        const CT dxp = pick<Tree::MAXD, CT>(q, dir) - pick<Tree::MAXD, CT>(nd.c, dir);
        from = child == (dxp <= (CT) 0 ? nd.l : nd.r) ? 1 : 2;
    }
}

// All point / node indices (O, F, lab, key, it.seed) live in the tree's index space; `root` is the tree's root there.
template <class Tree>
__device__ __forceinline__ int bfs_cluster(const Tree &tr, uint32_t root, const BfsItem &it, const int32_t *lab, uint32_t *O,
                                           uint32_t *F, unsigned long long *key, bool init_keys, unsigned *s_cnt, int lane) {
    typedef typename Tree::CT CT;
    const unsigned long long UNSEEN = ~0ull;
    const int sz = it.size, cid = it.cid;
    if (init_keys)
        for (int i = lane; i < sz; i += 32) key[O[i]] = UNSEEN;
    __syncwarp();
    if (lane == 0) {
        O[0] = (uint32_t) it.seed;
        key[it.seed] = 0;
    }
    __syncwarp();

    int lb = 0, le = 1;
    for (unsigned level = 1; lb < le; ++level) {
        if (lane == 0) *s_cnt = 0;
        __syncwarp();
        for (int base = lb; base < le; base += 32) {
            const int idx = base + lane;
            if (idx < le) {
                const uint32_t u = O[idx];
                CT q[Tree::MAXD];
                tr.point(u, q);
                kd_walk(tr, q, root, [&](uint32_t node, uint32_t t) {
                    if (node != u && lab[node] == cid) {
                        const unsigned long long nk = ((unsigned long long) level << (2 * KEY_B)) |
                                                      ((unsigned long long) idx << KEY_B) | (KEY_M - t);
                        const unsigned long long old = atomicMin(&key[node], nk);
                        if (old == UNSEEN) {
                            const unsigned slot = atomicAdd(s_cnt, 1u);
                            if (slot < (unsigned) RANK_MAX) F[le + slot] = node;
                        }
                    }
                });
            }
            __syncwarp();
        }
        __syncwarp();
        const int cnt = (int) *s_cnt;
        if (cnt <= RANK_MAX) {
            // rank the frontier by key (keys are unique: one (parent, visit counter) pair per claim)
            for (int i = lane; i < cnt; i += 32) {
                const uint32_t v = F[le + i];
                const unsigned long long kv = key[v];
                int r = 0;
                for (int j = 0; j < cnt; ++j) r += key[F[le + j]] < kv;
                O[le + r] = v;
            }
        } else {
            // large frontier: every parent walks again, counts the members whose final claim is its own, a running prefix
            // over the parents gives its block, and a third walk places them in reverse visit order
            int run = le;
            for (int base = lb; base < le; base += 32) {
                const int idx = base + lane;
                uint32_t own = 0;
                CT q[Tree::MAXD];
                uint32_t u = 0;
                const unsigned long long kbase = ((unsigned long long) level << (2 * KEY_B)) | ((unsigned long long) idx << KEY_B);
                if (idx < le) {
                    u = O[idx];
                    tr.point(u, q);
                    kd_walk(tr, q, root, [&](uint32_t node, uint32_t t) {
                        if (node != u && lab[node] == cid && key[node] == (kbase | (KEY_M - t))) ++own;
                    });
                }
                const uint32_t inc = warp_incl_scan(own);
                const int first = run + (int) (inc - own);
                if (idx < le && own) {
                    uint32_t ord = 0;
                    kd_walk(tr, q, root, [&](uint32_t node, uint32_t t) {
                        if (node != u && lab[node] == cid && key[node] == (kbase | (KEY_M - t))) {
                            O[first + (int) (own - 1u - ord)] = node;
                            ++ord;
                        }
                    });
                }
                run += (int) __shfl_sync(0xffffffffu, inc, 31);
            }
        }
        __syncwarp();
        lb = le;
        le += cnt;
    }
    return le;
}

__global__ void __launch_bounds__(BFS_THREADS) k_bfs_order(const BfsArgs a) {
    __shared__ unsigned s_cnt[BFS_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned n_items = min(*a.count, (unsigned) a.max_items);

    for (unsigned item = blockIdx.x * BFS_WARPS + wib; item < n_items; item += gridDim.x * BFS_WARPS) {
        const BfsItem it = a.items[item];
        const ProbDesc d = a.prob[it.pb];
        const uint32_t *pix = a.pix[d.pol] + d.off;
        const int32_t *lab = a.labels[d.pol] + d.off;
        uint32_t *O = a.members[d.pol] + d.off + it.mem_off;
        uint32_t *F = a.scratch[d.pol] + d.off + it.mem_off;
        unsigned long long *key = a.key[d.pol] + d.off;
        const int sz = it.size;
        PixTree tr;
        tr.pix = pix;
        tr.nodes = a.kd_nodes[d.pol] + d.off;  // one 16-byte load per node visit
        tr.set_eps(a.eps);
        const int le = bfs_cluster(tr, 0u, it, lab, O, F, key, a.init_keys != 0, &s_cnt[wib], lane);
        // the reference's median: nth_element over the member list by norm (CirclesEventFrame.cpp:137-147)
        if (it.kept >= 0 && a.ktab && lane == 0 && le == sz) {
            auto less = [&](uint32_t l, uint32_t r) {
                const uint32_t pl = pix[l], pr = pix[r];
                const uint32_t nl = ECB_PIX_X(pl) * ECB_PIX_X(pl) + ECB_PIX_Y(pl) * ECB_PIX_Y(pl);
                const uint32_t nr = ECB_PIX_X(pr) * ECB_PIX_X(pr) + ECB_PIX_Y(pr) * ECB_PIX_Y(pr);
                return nl < nr;
            };
            ecb_nth::nth_element(O, (long) sz, (long) (sz / 2), less);
            const uint32_t med = O[sz / 2];
            KeptCluster *kc = a.ktab + (size_t) it.pb * a.max_k + it.kept;
            kc->med_pid = (int32_t) med;
            kc->med_x = (int32_t) ECB_PIX_X(pix[med]) + 0;
            kc->med_y = (int32_t) ECB_PIX_Y(pix[med]) + 0;
        }
        __syncwarp();
    }
}

// general points (ecb_gridhash.cu): same BFS over the emulated tree of double coordinates
__global__ void __launch_bounds__(BFS_THREADS) k_bfs_order_general(const BfsArgs a, const double *__restrict__ P, int dim) {
    __shared__ unsigned s_cnt[BFS_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned n_items = min(*a.count, (unsigned) a.max_items);
    for (unsigned item = blockIdx.x * BFS_WARPS + wib; item < n_items; item += gridDim.x * BFS_WARPS) {
        BfsItem it = a.items[item];
        const ProbDesc d = a.prob[it.pb];
        // the emulated tree of the general path links points by their index in the whole batch: the walk runs in that index
        // space (root = first point of the problem) and the member list is rebased to pids afterwards
        GenTree tr;
        tr.P = P;
        tr.nodes = a.kd_nodes[0];
        tr.d = dim;
        tr.set_eps(a.eps);
        it.seed += (int32_t) d.off;
        uint32_t *O = a.members[0] + d.off + it.mem_off;
        bfs_cluster(tr, (uint32_t) d.off, it, a.labels[0], O, a.scratch[0] + d.off + it.mem_off, a.key[0], a.init_keys != 0,
                    &s_cnt[wib], lane);
        __syncwarp();
        for (int i = lane; i < it.size; i += 32) O[i] -= (uint32_t) d.off;
        __syncwarp();
    }
}

// Work items for ALL clusters of every problem (the DBSCAN::Run boundary with ordered `Clusters`): per cluster its size,
// its seed (lowest pid) and the offset of its list = exclusive scan of the sizes in discovery order.  One CTA per problem.
__global__ void __launch_bounds__(256) k_bfs_all_items(const ProbDesc *__restrict__ prob, const ProbHdr *__restrict__ hdr,
                                                        int n_prob, const int32_t *__restrict__ labels, uint32_t *csize,
                                                        uint32_t *cseed, uint32_t *coff, BfsItem *items, unsigned *count,
                                                        int cap) {
    __shared__ uint32_t ws[33];
    for (int pb = blockIdx.x; pb < n_prob; pb += gridDim.x) {
        const ProbDesc d = prob[pb];
        const int n = d.n, nc = hdr[pb].n_clusters;
        const int32_t *lab = labels + d.off;
        uint32_t *sz = csize + d.off, *sd = cseed + d.off, *co = coff + d.off;
        __syncthreads();
        for (int c = threadIdx.x; c < nc; c += blockDim.x) {
            sz[c] = 0;
            sd[c] = ECB_NONE;
        }
        __syncthreads();
        for (int pid = threadIdx.x; pid < n; pid += blockDim.x) {
            const int32_t l = lab[pid];
            if (l >= 0) {
                atomicAdd(&sz[l], 1u);
                atomicMin(&sd[l], (uint32_t) pid);
            }
        }
        __syncthreads();
        uint32_t run = 0;
        for (int c0 = 0; c0 < nc; c0 += blockDim.x) {
            const int c = c0 + threadIdx.x;
            const uint32_t v = c < nc ? sz[c] : 0;
            uint32_t tot;
            const uint32_t ex = block_excl_scan(v, ws, &tot);
            if (c < nc) {
                co[c] = run + ex;
                const unsigned slot = atomicAdd(count, 1u);
                if (slot < (unsigned) cap) {
                    BfsItem it;
                    it.pb = pb;
                    it.cid = c;
                    it.seed = (int32_t) sd[c];
                    it.size = (int32_t) v;
                    it.mem_off = (int32_t) (run + ex);
                    it.kept = -1;
                    items[slot] = it;
                }
            }
            run += tot;
        }
    }
}

}  // namespace

int ecb_launch_bfs_all_items(ecb_ctx *ctx, const ProbDesc *prob, const ProbHdr *hdr, int n_prob, const int32_t *labels,
                             uint32_t *csize, uint32_t *cseed, uint32_t *coff, BfsItem *items, unsigned *count, int cap) {
    if (n_prob <= 0) return ECB_OK;
    const int grid = n_prob < ctx->sm_count * 8 ? n_prob : ctx->sm_count * 8;
    k_bfs_all_items<<<grid, 256, 0, ctx->stream>>>(prob, hdr, n_prob, labels, csize, cseed, coff, items, count, cap);
    ECB_LAUNCHED(ctx);
    return ecb_check(ctx, cudaGetLastError(), "k_bfs_all_items launch");
}

int ecb_launch_bfs_general(ecb_ctx *ctx, BfsArgs &a, const double *P, int dim) {
    if (a.max_items <= 0) return ECB_OK;
    int grid = ctx->sm_count * 8;
    const int need = (a.max_items + BFS_WARPS - 1) / BFS_WARPS;
    if (grid > need) grid = need;
    k_bfs_order_general<<<grid, BFS_THREADS, 0, ctx->stream>>>(a, P, dim);
    ECB_LAUNCHED(ctx);
    return ecb_check(ctx, cudaGetLastError(), "k_bfs_order_general launch");
}

int ecb_launch_bfs(ecb_ctx *ctx, BfsArgs &a) {
    if (a.max_items <= 0) return ECB_OK;
    int grid = ctx->sm_count * 8;
    const int need = (a.max_items + BFS_WARPS - 1) / BFS_WARPS;
    if (grid > need) grid = need;
    ECB_PROF_BEGIN(ctx, ECB_STAGE_BFS);
    k_bfs_order<<<grid, BFS_THREADS, 0, ctx->stream>>>(a);
    ECB_PROF_END(ctx, ECB_STAGE_BFS);
    ECB_LAUNCHED(ctx);
    return ecb_check(ctx, cudaGetLastError(), "k_bfs_order launch");
}
