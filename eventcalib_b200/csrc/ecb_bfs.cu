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
#include <algorithm>

#include "ecb_cluster.cuh"
#include "ecb_nth_element.h"

namespace {

constexpr int BFS_THREADS = 64;  // two clusters per CTA: the small-cluster scratch is 20 KB of shared memory per warp
constexpr int BFS_WARPS = BFS_THREADS / 32;
constexpr int BFSG_THREADS = 128;  // k_bfs_order_general
constexpr int BFSG_WARPS = BFSG_THREADS / 32;

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

// ---- small clusters: neighbour lists from root paths, all (member, neighbour) pairs at once -----------------------------------
// The level-synchronous walk above keeps ~4 lanes busy (a ring cluster's frontier) and repeats a ~100-node tree walk per level:
// ncu counted 29 k warp instructions per 30-member cluster, all on one warp's critical path (0.27 ms per launch however small
// the batch).  For the usual kept cluster (<= 48 members, tree depth <= 48) the walk is replaced by what it computes:
//   * v is met by u's query iff d(u, v) <= eps and every ancestor a of v that has v on its FAR side as seen from u passes
//     find_nearest's pruning test |u[axis(a)] - a[axis(a)]| < eps (kdtree.cpp:166-171);
//   * the query meets nodes in pre-order, near subtree first.  Two met nodes are therefore visited in the order of their root
//     paths written as bits (0 near / 1 far per ancestor), compared left-aligned; when one path is a prefix of the other and
//     the rest is all "near", the shorter path (the ancestor) comes first.
// A0: every member climbs to the root once and stages its ancestors (coordinate on the ancestor's axis + its own side).
// A1: the in-range (u, v) pairs of the cluster are listed.  A2: one lane per PAIR evaluates visibility and the path key
// against v's staged chain.  A3: one lane per member ranks its visible neighbours, later visit first (the reference's list
// order: head insertion, kdtree.cpp:469-486).  B: the FIFO of expandCluster is replayed from the lists, one warp step per
// popped member.  Anything that does not fit falls back to the walk.
#ifndef ECB_BFS_SMALL
#define ECB_BFS_SMALL 1
#endif
constexpr int SM_MAXM = 48;      // members
constexpr int SM_MAXDEG = 32;    // visible neighbours of one member inside the cluster (one lane each when the member is popped)
constexpr int SM_MAXDEPTH = 48;  // tree depth of a member: 48 path bits + 6 depth bits in one 64-bit key
constexpr int SM_MAXPAIR = SM_MAXM * SM_MAXDEG;  // in-range ordered pairs of the cluster
constexpr unsigned long long SM_UNSEEN = ~0ull;

struct SmallScratch {
    uint32_t pix[SM_MAXM], node[SM_MAXM];
    uint8_t depth[SM_MAXM], deg[SM_MAXM], order[SM_MAXM];
    uint16_t pbase[SM_MAXM + 1];
    uint8_t nbr[SM_MAXM][SM_MAXDEG];
    uint8_t pair_v[SM_MAXPAIR];
    unsigned long long pair_key[SM_MAXPAIR];
    union {
        uint16_t chain[SM_MAXM][SM_MAXDEPTH];  // ancestors of member i, parent first: coordinate on the ancestor's axis | (i is on its right) << 15
        unsigned long long sel[SM_MAXM];       // afterwards: std::nth_element staging, norm^2 << 32 | node
    };
};

// returns the number of members placed in O (== sz on success), or -1 when the cluster does not fit the fast path
__device__ __forceinline__ int bfs_small(const PixTree &tr, const BfsItem &it, uint32_t *O, SmallScratch &w, int lane) {
    const int sz = it.size;
    if (sz > SM_MAXM) return -1;
    bool bad = false;
    // A0
    for (int i = lane; i < sz; i += 32) {
        const uint32_t nd = O[i];
        w.node[i] = nd;
        w.pix[i] = tr.pix[nd];
        int dep = 0;  // depth first (an ancestor's axis is the parity of ITS depth), then the chain; the second climb hits L1
        for (uint32_t c = __ldg(tr.nodes + nd).w; c != ECB_NONE && dep <= SM_MAXDEPTH; c = __ldg(tr.nodes + c).w) ++dep;
        bad |= dep > SM_MAXDEPTH;
        w.depth[i] = (uint8_t) dep;
        if (dep <= SM_MAXDEPTH) {
            uint32_t child = nd, anc = __ldg(tr.nodes + nd).w;
            for (int st = 0; st < dep; ++st) {
                const uint4 an = __ldg(tr.nodes + anc);
                const uint32_t c = ((dep - 1 - st) & 1) ? ECB_PIX_Y(an.x) : ECB_PIX_X(an.x);
                w.chain[i][st] = (uint16_t) (c | (child == an.z ? 0x8000u : 0u));
                child = anc;
                anc = an.w;
            }
        }
    }
    if (__any_sync(0xffffffffu, bad)) return -2;
    __syncwarp();
    // A1: in-range pairs, grouped by u
    int cnt[2] = {0, 0};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int i = lane + 32 * r;
        if (i < sz) {
            const uint32_t pu = w.pix[i];
            const int ux = (int) ECB_PIX_X(pu), uy = (int) ECB_PIX_Y(pu);
            for (int j = 0; j < sz; ++j) {
                const uint32_t pv = w.pix[j];
                const int ex = (int) ECB_PIX_X(pv) - ux, ey = (int) ECB_PIX_Y(pv) - uy;
                cnt[r] += (j != i) && (ex * ex + ey * ey <= tr.eps2i);
            }
        }
    }
    const uint32_t inc0 = warp_incl_scan((uint32_t) cnt[0]);
    const uint32_t tot0 = __shfl_sync(0xffffffffu, inc0, 31);
    const uint32_t inc1 = warp_incl_scan((uint32_t) cnt[1]);
    const int n_pair = (int) (tot0 + __shfl_sync(0xffffffffu, inc1, 31));
    if (n_pair > SM_MAXPAIR) return -3;
    {
        const int base[2] = {(int) (inc0 - cnt[0]), (int) (tot0 + inc1 - cnt[1])};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int i = lane + 32 * r;
            if (i < sz) {
                w.pbase[i] = (uint16_t) base[r];
                const uint32_t pu = w.pix[i];
                const int ux = (int) ECB_PIX_X(pu), uy = (int) ECB_PIX_Y(pu);
                int o = base[r];
                for (int j = 0; j < sz; ++j) {
                    const uint32_t pv = w.pix[j];
                    const int ex = (int) ECB_PIX_X(pv) - ux, ey = (int) ECB_PIX_Y(pv) - uy;
                    if ((j != i) && (ex * ex + ey * ey <= tr.eps2i)) {
                        w.pair_v[o] = (uint8_t) j;
                        w.pair_key[o] = (unsigned long long) i;  // u, replaced by the key in A2
                        ++o;
                    }
                }
            }
        }
        if (lane == 0) w.pbase[sz] = (uint16_t) n_pair;
    }
    __syncwarp();
    // A2: one lane per pair
    for (int p = lane; p < n_pair; p += 32) {
        const int u = (int) w.pair_key[p], v = (int) w.pair_v[p];
        const uint32_t pu = w.pix[u];
        const int ux = (int) ECB_PIX_X(pu), uy = (int) ECB_PIX_Y(pu);
        const int dv = (int) w.depth[v];
        unsigned long long bits = 0;
        bool seen = true;
        for (int st = dv - 1; st >= 0; --st) {  // root first
            const uint32_t e = w.chain[v][st];
            const int dx = (((dv - 1 - st) & 1) ? uy : ux) - (int) (e & 0x7FFFu);
            const bool far = (dx <= 0) == (bool) (e >> 15);  // the near child is the left one iff dx <= 0
            if (far && !tr.far_ok(dx)) {
                seen = false;
                break;
            }
            bits = (bits << 1) | (far ? 1ull : 0ull);
        }
        w.pair_key[p] = seen ? (((bits << (SM_MAXDEPTH - dv)) << 6) | (unsigned long long) dv) : SM_UNSEEN;
    }
    __syncwarp();
    // A3: rank the visible neighbours of every member by descending key (keys of one member are distinct)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int i = lane + 32 * r;
        if (i < sz) {
            const int b = w.pbase[i], e = w.pbase[i + 1];
            int deg = 0;
            for (int k = b; k < e; ++k) {
                const unsigned long long kk = w.pair_key[k];
                if (kk == SM_UNSEEN) continue;
                ++deg;
                int rk = 0;
                for (int q = b; q < e; ++q) {
                    const unsigned long long kq = w.pair_key[q];
                    rk += kq != SM_UNSEEN && kq > kk;
                }
                if (rk < SM_MAXDEG) w.nbr[i][rk] = w.pair_v[k];
            }
            bad |= deg > SM_MAXDEG;
            w.deg[i] = (uint8_t) deg;
        }
    }
    if (__any_sync(0xffffffffu, bad)) return -4;
    __syncwarp();
    // B: the FIFO of expandCluster, the unseen neighbours of the popped member appended in list order
    uint32_t seen_lo = 1u, seen_hi = 0u;  // member 0 = the seed (lowest pid)
    if (lane == 0) w.order[0] = 0;
    int tail = 1;
    for (int h = 0; h < tail; ++h) {
        __syncwarp();
        const int u = w.order[h];
        const int d = w.deg[u];
        int v = 0;
        bool fresh = false;
        if (lane < d) {
            v = w.nbr[u][lane];
            fresh = !(((v < 32 ? seen_lo : seen_hi) >> (v & 31)) & 1u);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, fresh);
        if (!bal) continue;
        if (fresh) w.order[tail + __popc(bal & ((1u << lane) - 1u))] = (uint8_t) v;
        seen_lo |= __reduce_or_sync(0xffffffffu, (fresh && v < 32) ? 1u << v : 0u);
        if (sz > 32) seen_hi |= __reduce_or_sync(0xffffffffu, (fresh && v >= 32) ? 1u << (v - 32) : 0u);
        tail += __popc(bal);
    }
    __syncwarp();
    for (int i = lane; i < tail; i += 32) O[i] = w.node[w.order[i]];
    __syncwarp();
    return tail;
}

__global__ void __launch_bounds__(BFS_THREADS) k_bfs_order(const BfsArgs a) {
    __shared__ unsigned s_cnt[BFS_WARPS];
#if ECB_BFS_SMALL
    __shared__ SmallScratch s_small[BFS_WARPS];
#endif
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
        int le = -1;
#if ECB_BFS_SMALL
        if (a.init_keys) le = bfs_small(tr, it, O, s_small[wib], lane);  // O holds the member list (front-end items) only then
#endif
        const bool small = le >= 0;
        if (!small) le = bfs_cluster(tr, 0u, it, lab, O, F, key, a.init_keys != 0, &s_cnt[wib], lane);
        // the reference's median: nth_element over the member list by norm (CirclesEventFrame.cpp:137-147)
        if (it.kept >= 0 && a.ktab && le == sz) {
            uint32_t med = 0;
            auto norm2 = [&](uint32_t m) {
                const uint32_t p = pix[m];
                return ECB_PIX_X(p) * ECB_PIX_X(p) + ECB_PIX_Y(p) * ECB_PIX_Y(p);
            };
#if ECB_BFS_SMALL
            if (small) {  // the selection is sequential: run it on a shared-memory copy, put the permuted list back in parallel
                unsigned long long *x = s_small[wib].sel;
                for (int i = lane; i < sz; i += 32) {
                    const uint32_t m = O[i];
                    x[i] = ((unsigned long long) norm2(m) << 32) | m;
                }
                __syncwarp();
                if (lane == 0) {
                    auto less = [](unsigned long long l, unsigned long long r) { return (uint32_t) (l >> 32) < (uint32_t) (r >> 32); };
                    ecb_nth::nth_element(x, (long) sz, (long) (sz / 2), less);
                }
                __syncwarp();
                for (int i = lane; i < sz; i += 32) O[i] = (uint32_t) x[i];
                med = (uint32_t) x[sz / 2];
            } else
#endif
            if (lane == 0) {
                auto less = [&](uint32_t l, uint32_t r) { return norm2(l) < norm2(r); };
                ecb_nth::nth_element(O, (long) sz, (long) (sz / 2), less);
                med = O[sz / 2];
            }
            if (lane == 0) {
                KeptCluster *kc = a.ktab + (size_t) it.pb * a.max_k + it.kept;
                kc->med_pid = (int32_t) med;
                kc->med_x = (int32_t) ECB_PIX_X(pix[med]) + 0;
                kc->med_y = (int32_t) ECB_PIX_Y(pix[med]) + 0;
            }
        }
        __syncwarp();
    }
}

// general points (ecb_gridhash.cu): same BFS over the emulated tree of double coordinates
__global__ void __launch_bounds__(BFSG_THREADS) k_bfs_order_general(const BfsArgs a, const double *__restrict__ P, int dim) {
    __shared__ unsigned s_cnt[BFSG_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned n_items = min(*a.count, (unsigned) a.max_items);
    for (unsigned item = blockIdx.x * BFSG_WARPS + wib; item < n_items; item += gridDim.x * BFSG_WARPS) {
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
    const int need = (a.max_items + BFSG_WARPS - 1) / BFSG_WARPS;
    if (grid > need) grid = need;
    k_bfs_order_general<<<grid, BFSG_THREADS, 0, ctx->stream>>>(a, P, dim);
    ECB_LAUNCHED(ctx);
    return ecb_check(ctx, cudaGetLastError(), "k_bfs_order_general launch");
}

int ecb_launch_bfs(ecb_ctx *ctx, BfsArgs &a) {
    if (a.max_items <= 0) return ECB_OK;
    int per_sm = 8;  // resident CTAs (the small-cluster scratch in shared memory decides); items are strided over the grid
    ECB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bfs_order, BFS_THREADS, 0));
    int grid = ctx->sm_count * std::max(per_sm, 1);
    const int need = (a.max_items + BFS_WARPS - 1) / BFS_WARPS;
    if (grid > need) grid = need;
    ECB_PROF_BEGIN(ctx, ECB_STAGE_BFS);
    k_bfs_order<<<grid, BFS_THREADS, 0, ctx->stream>>>(a);
    ECB_PROF_END(ctx, ECB_STAGE_BFS);
    ECB_LAUNCHED(ctx);
    return ecb_check(ctx, cudaGetLastError(), "k_bfs_order launch");
}
