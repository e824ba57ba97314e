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

__device__ __forceinline__ double pix_coord(uint32_t p, int dir) { return (double) (dir ? ECB_PIX_Y(p) : ECB_PIX_X(p)); }

__global__ void __launch_bounds__(BFS_THREADS) k_bfs_order(const BfsArgs a) {
    __shared__ unsigned s_cnt[BFS_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned n_items = min(*a.count, (unsigned) a.max_items);
    const double eps = a.eps, eps2 = eps * eps;
    const unsigned long long UNSEEN = ~0ull;

    for (unsigned item = blockIdx.x * BFS_WARPS + wib; item < n_items; item += gridDim.x * BFS_WARPS) {
        const BfsItem it = a.items[item];
        const ProbDesc d = a.prob[it.pb];
        const uint32_t *pix = a.pix[d.pol] + d.off;
        const int32_t *lab = a.labels[d.pol] + d.off;
        const uint4 *nodes = a.kd_nodes[d.pol] + d.off;  // one 16-byte load per node visit
        uint32_t *O = a.members[d.pol] + d.off + it.mem_off;
        uint32_t *F = a.scratch[d.pol] + d.off + it.mem_off;
        unsigned long long *key = a.key[d.pol] + d.off;
        const int sz = it.size, cid = it.cid;

        if (a.init_keys)
            for (int i = lane; i < sz; i += 32) key[O[i]] = UNSEEN;
        __syncwarp();
        if (lane == 0) {
            O[0] = (uint32_t) it.seed;
            key[it.seed] = 0;
        }
        __syncwarp();

        int lb = 0, le = 1;
        for (unsigned level = 1; lb < le; ++level) {
            if (lane == 0) s_cnt[wib] = 0;
            __syncwarp();
            for (int base = lb; base < le; base += 32) {
                const int idx = base + lane;
                if (idx < le) {
                    const uint32_t u = O[idx];
                    const uint32_t pu = pix[u];
                    const double qx = (double) ECB_PIX_X(pu), qy = (double) ECB_PIX_Y(pu);
                    // find_nearest(root, q, eps) without a stack: state = how the current node was entered
                    uint32_t node = 0;
                    int dir = 0, from = 0;  // 0: from the parent, 1: back from the near child, 2: back from the far child
                    uint32_t t = 0;
                    uint4 nd = __ldg(nodes + node);
                    for (;;) {
                        const uint32_t pn = nd.x;
                        const double dx = (dir ? qy : qx) - pix_coord(pn, dir);
                        if (from == 0) {
                            const double ex = (double) ECB_PIX_X(pn) - qx, ey = (double) ECB_PIX_Y(pn) - qy;
                            if (ex * ex + ey * ey <= eps2 && node != u && lab[node] == cid) {
                                const unsigned long long nk = ((unsigned long long) level << 44) |
                                                              ((unsigned long long) idx << 22) | (0x3FFFFFu - t);
                                const unsigned long long old = atomicMin(&key[node], nk);
                                if (old == UNSEEN) F[le + atomicAdd(&s_cnt[wib], 1u)] = node;
                            }
                            ++t;
                            const uint32_t nearc = dx <= 0.0 ? nd.y : nd.z;
                            if (nearc != ECB_NONE) {
                                node = nearc;
                                nd = __ldg(nodes + node);
                                dir ^= 1;
                                continue;
                            }
                            from = 1;
                        }
                        if (from == 1) {
                            const uint32_t farc = dx <= 0.0 ? nd.z : nd.y;
                            if (fabs(dx) < eps && farc != ECB_NONE) {
                                node = farc;
                                nd = __ldg(nodes + node);
                                dir ^= 1;
                                from = 0;
                                continue;
                            }
                        }
                        // this subtree is done: climb, and find out which child we are coming back from
                        if (node == 0) break;
                        const uint32_t child = node;
                        node = nd.w;
                        nd = __ldg(nodes + node);
                        dir ^= 1;
                        const double dxp = (dir ? qy : qx) - pix_coord(nd.x, dir);
                        from = child == (dxp <= 0.0 ? nd.y : nd.z) ? 1 : 2;
                    }
                }
                __syncwarp();
            }
            __syncwarp();
            const int cnt = (int) s_cnt[wib];
            // rank the frontier by key (keys are unique: one (parent, visit counter) pair per claim)
            for (int i = lane; i < cnt; i += 32) {
                const uint32_t v = F[le + i];
                const unsigned long long kv = key[v];
                int r = 0;
                for (int j = 0; j < cnt; ++j) r += key[F[le + j]] < kv;
                O[le + r] = v;
            }
            __syncwarp();
            lb = le;
            le += cnt;
        }
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
