// Ingest + window kernels.
//
//   k_ingest   25-byte packed reference records (Event.hpp:41-47) -> SoA  t f64[n], xyp u32[n]
//              (HBM streaming: TMA bulk copies of the byte stream into a 4-stage shared-memory ring, funnel-shift
//              unpack, coalesced 8-byte / 4-byte stores).  Algorithmic traffic 25 B read + 12 B written.
//   k_bounds   window [t0,t1] CLOSED -> event index range [lower_bound(t0), upper_bound(t1))
//              (EventFrame.cpp:14-15 on the time-ordered multimap)
//   k_window   one CTA per window: per-polarity set of distinct pixels in FIRST-ARRIVAL order (the insertion
//              sequence of the reference's two unordered_sets, EventFrame.cpp:12-21), +/- cancellation
//              (EventFrame.cpp:23-32), and the pid order handed to DBSCAN (order_mode 0: arrival order of the
//              survivors; order_mode 1: libstdc++ unordered_set iteration order, see k_uset_order).
#include "ecb_cluster.cuh"
#include "ecb_window.cuh"

namespace {

// ------------------------------------------------------------------------------------------ ingest ----
constexpr int ING_REC = 256;                 // records per tile
constexpr int ING_BYTES = ING_REC * 25;      // 6400 = 400 x 16 B
constexpr int ING_THREADS = 256;

__device__ __forceinline__ uint32_t ld_unaligned32(const uint32_t *w, int byte_off) {
    const int wi = byte_off >> 2, sh = (byte_off & 3) * 8;
    return __funnelshift_r(w[wi], w[wi + 1], sh);
}

// TMA (1-D bulk async copy) staging: one elected thread streams 6400-byte tiles of the packed record stream into a ring of
// shared-memory buffers (cp.async.bulk ... mbarrier::complete_tx), ING_STAGES tiles ahead of the threads that unpack
// them, so every SM keeps ~200 KB of HBM reads in flight instead of one synchronous tile per CTA.
constexpr int ING_STAGES = 4;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(ING_THREADS) k_ingest(const uint8_t *__restrict__ raw, int64_t n, int W, int H,
                                                         double *__restrict__ out_t, uint32_t *__restrict__ out_xyp,
                                                         uint32_t *__restrict__ flags) {
    __shared__ __align__(128) uint32_t sm[ING_STAGES][ING_BYTES / 4 + 4];
    __shared__ __align__(8) uint64_t full[ING_STAGES];
    const int64_t n_tiles = (n + ING_REC - 1) / ING_REC;
    const int64_t total_bytes = n * 25;
    const int tid = threadIdx.x;
    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int64_t my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto issue = [&](int64_t it) {  // thread 0: arm the stage's barrier and start the bulk copy of local tile `it`
        const int st = (int) (it % ING_STAGES);
        const int64_t byte0 = (blockIdx.x + it * gridDim.x) * (int64_t) ING_BYTES;
        const uint32_t bytes = (uint32_t) (min((int64_t) ING_BYTES, total_bytes - byte0) & ~(int64_t) 15);
        mbar_expect_tx(&full[st], bytes);
        if (bytes) bulk_g2s(sm[st], raw + byte0, bytes, &full[st]);
    };
    if (tid == 0) {
        for (int st = 0; st < ING_STAGES; ++st) mbar_init(&full[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int64_t it = 0; it < my_tiles && it < ING_STAGES; ++it) issue(it);
    uint32_t bad = 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
        const int st = (int) (it % ING_STAGES);
        const int64_t tile = blockIdx.x + it * gridDim.x;
        const int64_t rec0 = tile * ING_REC;
        const int64_t byte0 = rec0 * 25;
        const int64_t bytes = min((int64_t) ING_BYTES, total_bytes - byte0);
        uint32_t *buf = sm[st];
        // the < 16 trailing bytes of the stream are not a whole bulk-copy unit
        if ((bytes & 15) && tid < (int) (bytes & 15)) {
            const int o = (int) (bytes & ~(int64_t) 15) + tid;
            reinterpret_cast<uint8_t *>(buf)[o] = raw[byte0 + o];
        }
        mbar_wait(&full[st], (uint32_t) ((it / ING_STAGES) & 1));
        if (bytes & 15) __syncthreads();
        const int r = tid;
        if (rec0 + r < n) {
            const int o = r * 25;
            const uint32_t t_lo = ld_unaligned32(buf, o), t_hi = ld_unaligned32(buf, o + 4);
            const uint32_t x_lo = ld_unaligned32(buf, o + 8), x_hi = ld_unaligned32(buf, o + 12);
            const uint32_t y_lo = ld_unaligned32(buf, o + 16), y_hi = ld_unaligned32(buf, o + 20);
            const uint32_t pol = reinterpret_cast<const uint8_t *>(buf)[o + 24];
            const double t = __hiloint2double((int) t_hi, (int) t_lo);
            const double x = __hiloint2double((int) x_hi, (int) x_lo);
            const double y = __hiloint2double((int) y_hi, (int) y_lo);
            uint32_t w = pol ? 0x80000000u : 0u;
            const int xi = (int) x, yi = (int) y;
            if (!((double) xi == x && (double) yi == y && xi >= 0 && xi < W && yi >= 0 && yi < H)) {
                w |= ECB_PIX_INVALID;
                bad |= 1u;
            } else {
                w |= (uint32_t) xi | ((uint32_t) yi << 15);
            }
            // time order: compare with the previous record (same tile: shared memory; else global)
            double tp = t;
            if (r > 0) {
                tp = __hiloint2double((int) ld_unaligned32(buf, o - 25 + 4), (int) ld_unaligned32(buf, o - 25));
            } else if (rec0 > 0) {
                const uint8_t *q = raw + byte0 - 25;
                unsigned long long v = 0;
                for (int b = 7; b >= 0; --b) v = (v << 8) | q[b];
                tp = __longlong_as_double((long long) v);
            }
            if (!(tp <= t)) bad |= 2u;
            out_t[rec0 + r] = t;
            out_xyp[rec0 + r] = w;
        }
        __syncthreads();  // every thread is done with this stage's buffer
        if (tid == 0 && it + ING_STAGES < my_tiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads above before the async write below
            issue(it + ING_STAGES);
        }
    }
    if (bad) atomicOr(flags, bad);
}

// ------------------------------------------------------------------------------------------ bounds ----
__global__ void k_bounds(const double *__restrict__ t, int64_t n, const double *__restrict__ win, int n_win,
                         int64_t *__restrict__ lohi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n_win) return;
    const double v = win[i];
    int64_t lo = 0, hi = n;
    if ((i & 1) == 0) {  // lower_bound(first): first index with t >= v
        while (lo < hi) {
            int64_t m = (lo + hi) >> 1;
            if (t[m] < v) lo = m + 1; else hi = m;
        }
    } else {             // upper_bound(second): first index with t > v
        while (lo < hi) {
            int64_t m = (lo + hi) >> 1;
            if (t[m] <= v) lo = m + 1; else hi = m;
        }
    }
    lohi[i] = lo;
}

// ------------------------------------------------------------------------------------------ window ----
constexpr int WIN_THREADS = 512;
constexpr int WIN_HASH = 2048;

template <bool SM>  // SM: bit planes in shared memory, else in per-CTA L2 scratch (large sensors)
__global__ void __launch_bounds__(WIN_THREADS) k_window(const WindowArgs a) {
    extern __shared__ __align__(16) uint32_t smw[];
    __shared__ uint32_t ws[33];
    __shared__ uint32_t hash[WIN_HASH];
    const int tid = threadIdx.x, nthr = WIN_THREADS;
    const int RW = a.RW, NWp = RW * a.H;
    uint32_t *pl0;
    if constexpr (SM) pl0 = smw;
    else pl0 = a.gplanes + (size_t) blockIdx.x * 2 * NWp;
    uint32_t *plane[2] = {pl0, pl0 + NWp};

    for (int w = blockIdx.x; w < a.n_win; w += gridDim.x) {
        const int64_t lo = a.lohi[2 * w], hi = max(a.lohi[2 * w + 1], a.lohi[2 * w]);
        const int64_t off = a.ptoff[w];
        __syncthreads();
        for (int i = tid; i < 2 * NWp; i += nthr) pl0[i] = 0;
        for (int i = tid; i < WIN_HASH; i += nthr) hash[i] = ECB_NONE;
        __syncthreads();
        uint32_t base[2] = {0, 0};  // arrival counts so far (uniform across the block)
        uint32_t *arr[2] = {a.arrive[0] + off, a.arrive[1] + off};
        for (int64_t c0 = lo; c0 < hi; c0 += nthr) {
            const int64_t i = c0 + tid;
            bool cand = false;
            uint32_t e = 0, pol = 0, key = 0;
            int slot = -1;
            if (i < hi) {
                e = a.xyp[i];
                if (!(e & ECB_PIX_INVALID)) {
                    pol = e >> 31;
                    const int x = ECB_PIX_X(e), y = ECB_PIX_Y(e);
                    cand = !((plane[pol][y * RW + (x >> 5)] >> (x & 31)) & 1u);
                    key = (uint32_t) (y * a.W + x) | (pol << 20);
                }
            }
            if (cand) {  // in-chunk first arrival: min tid per (pixel, polarity) through a small hash
                const uint32_t mine = (key << 10) | (uint32_t) tid;
                uint32_t h = (key * 2654435761u) >> 21;  // 11 bits
                for (;;) {
                    uint32_t cur = hash[h];
                    if (cur == ECB_NONE) {
                        cur = atomicCAS(&hash[h], ECB_NONE, mine);
                        if (cur == ECB_NONE) {
                            slot = (int) h;
                            break;
                        }
                    }
                    if ((cur >> 10) == key) {
                        atomicMin(&hash[h], mine);
                        slot = (int) h;
                        break;
                    }
                    h = (h + 1) & (WIN_HASH - 1);
                }
            }
            __syncthreads();
            const bool win = cand && hash[slot] == ((key << 10) | (uint32_t) tid);
            __syncthreads();
            if (cand) hash[slot] = ECB_NONE;
            uint32_t tot;
            const uint32_t v = win ? (pol ? 0x10000u : 1u) : 0u;
            const uint32_t ex = block_excl_scan(v, ws, &tot);
            if (win) {
                const int x = ECB_PIX_X(e), y = ECB_PIX_Y(e);
                atomicOr(&plane[pol][y * RW + (x >> 5)], 1u << (x & 31));
                const uint32_t r = pol ? (ex >> 16) : (ex & 0xFFFF);
                arr[pol][base[pol] + r] = ECB_PIX_XY(e);
            }
            base[0] += tot & 0xFFFF;
            base[1] += tot >> 16;
            __syncthreads();
        }
        // ---- +/- cancellation and pid order -------------------------------------------------------
        // order_mode 0: pid = arrival order of the surviving pixels.  order_mode 1: only mark the cancelled pixels;
        // k_uset_order (ecb_order.cu) emits the survivors in the libstdc++ iteration order.
        for (int pol = 0; pol < 2; ++pol) {
            const uint32_t m = base[pol];
            const uint32_t *other = plane[pol ^ 1];
            uint32_t *dst = a.pts[pol] + off;
            uint32_t run = 0;
            for (uint32_t c0 = 0; c0 < m; c0 += nthr) {
                const uint32_t i = c0 + tid;
                uint32_t p = 0;
                bool keep = false;
                if (i < m) {
                    p = arr[pol][i];
                    const int x = ECB_PIX_X(p), y = ECB_PIX_Y(p);
                    keep = !((other[y * RW + (x >> 5)] >> (x & 31)) & 1u);
                }
                uint32_t tot;
                const uint32_t ex = block_excl_scan(keep ? 1u : 0u, ws, &tot);
                if (a.order_mode == 1) {
                    if (i < m && !keep) arr[pol][i] = p | 0x80000000u;  // cancelled; k_uset_order compacts in set order
                } else if (keep) {
                    dst[run + ex] = p;
                }
                run += tot;
            }
            if (tid == 0) {
                ProbDesc d;
                d.off = off;
                d.n = (int32_t) run;
                d.pol = pol;
                d.x0 = 0;
                d.y0 = 0;
                d.pad0 = (int32_t) m;  // arrival count (distinct pixels before cancellation)
                d.pad1 = 0;
                a.prob[2 * w + pol] = d;
                atomicMax(a.max_n, run);
                atomicMax(a.max_n + 1, m);
            }
        }
    }
}

}  // namespace

int ecb_launch_ingest(ecb_ctx *ctx, const void *d_raw, int64_t n) {
    if (n <= 0) return ECB_OK;
    int64_t tiles = (n + ING_REC - 1) / ING_REC;
    int grid = (int) (tiles < (int64_t) ctx->sm_count * 8 ? tiles : (int64_t) ctx->sm_count * 8);  // 8 CTAs x 4 stages x 6.4 KB per SM
    ECB_PROF_BEGIN(ctx, ECB_STAGE_INGEST);
    k_ingest<<<grid, ING_THREADS, 0, ctx->stream>>>((const uint8_t *) d_raw, n, ctx->width, ctx->height,
                                                     (double *) ctx->ev_t.p, (uint32_t *) ctx->ev_xyp.p,
                                                     (uint32_t *) ctx->ev_flag.p);
    ECB_PROF_END(ctx, ECB_STAGE_INGEST);
    ECB_LAUNCHED(ctx);
    return ecb_check(ctx, cudaGetLastError(), "k_ingest launch");
}

int ecb_launch_bounds(ecb_ctx *ctx, const double *d_win, int n_win, int64_t *d_lohi) {
    if (n_win <= 0) return ECB_OK;
    const int thr = 128;
    ECB_PROF_BEGIN(ctx, ECB_STAGE_BOUNDS);
    k_bounds<<<(2 * n_win + thr - 1) / thr, thr, 0, ctx->stream>>>((const double *) ctx->ev_t.p, ctx->n_events, d_win,
                                                                    n_win, d_lohi);
    ECB_PROF_END(ctx, ECB_STAGE_BOUNDS);
    ECB_LAUNCHED(ctx);
    return ecb_check(ctx, cudaGetLastError(), "k_bounds launch");
}

int ecb_launch_window(ecb_ctx *ctx, WindowArgs &a) {
    if (a.n_win <= 0) return ECB_OK;
    a.RW = (a.W + 31) >> 5;
    size_t smem = (size_t) 2 * a.RW * a.H * 4;
    const size_t limit = (size_t) ctx->smem_optin - 12 * 1024;
    const bool gpl = smem > limit;  // large sensors: planes in per-CTA L2 scratch
    if (gpl) smem = 0;
    void (*kern)(const WindowArgs) = gpl ? k_window<false> : k_window<true>;
    ECB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) limit)  /* constant: race-free */);
    int per_sm = 1;
    ECB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WIN_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    int grid = ctx->sm_count * per_sm;
    if (grid > a.n_win) grid = a.n_win;
    a.gplanes = nullptr;
    if (gpl) {
        int rc = ecb_reserve(ctx, ctx->scratch, (size_t) grid * 2 * a.RW * a.H * 4);
        if (rc) return rc;
        a.gplanes = (uint32_t *) ctx->scratch.p;
    }
    ECB_PROF_BEGIN(ctx, ECB_STAGE_WINDOW);
    kern<<<grid, WIN_THREADS, smem, ctx->stream>>>(a);
    ECB_PROF_END(ctx, ECB_STAGE_WINDOW);
    ECB_LAUNCHED(ctx);
    return ecb_check(ctx, cudaGetLastError(), "k_window launch");
}
