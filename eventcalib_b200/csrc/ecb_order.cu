// k_uset_order — the pid order the reference hands to DBSCAN::Run.
//
// EventFrame::EventFrame (event/src/EventFrame.cpp:12-35) inserts the window's pixels, in time order, into one
// std::unordered_set<Vector2d, EigenMatrixHash> per polarity (hash: core/utility/include/opengv2/utility/utility.hpp:38-51
// = boost hash_combine over std::hash<double>), erases the pixels present in both sets and copies the sets to
// positiveEvents_/negativeEvents_ in ITERATION order.  That order is a pure function of the insertion sequence and of
// libstdc++'s hashtable policy (external: libstdc++ hashtable.h _M_insert_bucket_begin / _M_rehash_aux, hashtable_c++0x.cc
// prime policy; g++ 13):
//   * a node goes to the FRONT of its bucket's run if the bucket is occupied, else to the front of the whole list;
//   * inserting the (B+1)-th element into B buckets first rehashes to the next policy prime (13, 29, 59, 127, ...),
//     re-inserting the nodes in current list order by the same rule;
//   * erase keeps the relative order of the remaining nodes.
// So one "stage" (a rehash followed by the insertions up to the next rehash) maps the sequence S = old list order ++ new
// arrivals to:  buckets in DESCENDING order of their first position in S, inside a bucket DESCENDING position in S.
// That is a grouping, not a sequential process: atomicMin gives the first position, an atomicExch chain the bucket's
// members, a suffix scan over the bucket heads the run starts.  Stage sizes grow geometrically, so the whole emulation
// costs about two passes over the final set.  Checked against the real std::unordered_set in tests/test_gpu_frontend.py.
//
// One CTA per (window, polarity); all arrays in shared memory when they fit, else in per-CTA L2 scratch.
#include <string.h>

#include <algorithm>
#include <vector>

#include "ecb_window.cuh"

namespace {

constexpr int ORD_THREADS = 256;       // problems that fit shared memory: several CTAs per SM
constexpr int ORD_THREADS_BIG = 1024;  // large problems in L2 scratch: few CTAs, as many threads each as possible
constexpr int ORD_NSTAGE = 17;
__constant__ uint32_t c_prime[ORD_NSTAGE] = {13,    29,     59,     127,    257,    541,     1109,   2357,   5087,
                                             10273, 20753,  42043,  85229,  172933, 351061,  712697, 1447153};
static const uint32_t h_prime[ORD_NSTAGE] = {13,    29,     59,     127,    257,    541,     1109,   2357,   5087,
                                             10273, 20753,  42043,  85229,  172933, 351061,  712697, 1447153};

#ifndef ECB_ORD_POS16
#define ECB_ORD_POS16 1  // 16-bit per-position arrays when the problems are small enough (more CTAs per SM)
#endif
#ifndef ECB_ORD_HTAB
#define ECB_ORD_HTAB 1  // coordinate hashes from a table (one entry per sensor column / row) instead of two murmur rounds per pixel
#endif

// std::hash<double> of libstdc++: 0 for +-0.0, else _Hash_bytes(&v, 8, 0xc70f6907) (64-bit murmur variant)
__host__ __device__ __forceinline__ uint64_t hash_double(double v) {
    if (v == 0.0) return 0;
    const uint64_t mul = (0xc6a4a793ull << 32) + 0x5bd1e995ull;
    uint64_t h = 0xc70f6907ull ^ (8ull * mul);
#ifdef __CUDA_ARCH__
    uint64_t d = (uint64_t) __double_as_longlong(v) * mul;
#else
    uint64_t d;
    memcpy(&d, &v, 8);
    d *= mul;
#endif
    d ^= d >> 47;
    d *= mul;
    h ^= d;
    h *= mul;
    h ^= h >> 47;
    h *= mul;
    h ^= h >> 47;
    return h;
}

// EigenMatrixHash<Vector2d> (utility.hpp:38-51)
__device__ __forceinline__ uint64_t hash_pixel(uint32_t pix) {
    uint64_t seed = 0;
    seed ^= hash_double((double) ECB_PIX_X(pix)) + 0x9e3779b9ull + (seed << 6) + (seed >> 2);
    seed ^= hash_double((double) ECB_PIX_Y(pix)) + 0x9e3779b9ull + (seed << 6) + (seed >> 2);
    return seed;
}

// the same with tab[v] = hash_double((double) v) + 0x9e3779b9 (the seed starts at 0, so the first combine is the table entry)
__device__ __forceinline__ uint64_t hash_pixel_tab(uint32_t pix, const unsigned long long *__restrict__ tab) {
    const uint64_t seed = __ldg(tab + ECB_PIX_X(pix));
    return seed ^ (__ldg(tab + ECB_PIX_Y(pix)) + (seed << 6) + (seed >> 2));
}

// h % c_prime[stage] with compile-time divisors (the stage is uniform across the CTA)
__device__ __forceinline__ uint32_t bucket_of(uint64_t h, int stage) {
    switch (stage) {
        case 0: return (uint32_t) (h % 13ull);
        case 1: return (uint32_t) (h % 29ull);
        case 2: return (uint32_t) (h % 59ull);
        case 3: return (uint32_t) (h % 127ull);
        case 4: return (uint32_t) (h % 257ull);
        case 5: return (uint32_t) (h % 541ull);
        case 6: return (uint32_t) (h % 1109ull);
        case 7: return (uint32_t) (h % 2357ull);
        case 8: return (uint32_t) (h % 5087ull);
        case 9: return (uint32_t) (h % 10273ull);
        case 10: return (uint32_t) (h % 20753ull);
        case 11: return (uint32_t) (h % 42043ull);
        case 12: return (uint32_t) (h % 85229ull);
        case 13: return (uint32_t) (h % 172933ull);
        case 14: return (uint32_t) (h % 351061ull);
        case 15: return (uint32_t) (h % 712697ull);
        default: return (uint32_t) (h % 1447153ull);
    }
}

// SM: arrays in shared memory (pointers derived from the shared array only -> LDS/STS/ATOMS), else per-CTA L2 scratch
// PosT: type of the per-position arrays (positions, arrival indices, bucket numbers): 16 bits when the largest problem and its
// bucket count stay below 2^15 — the shared-memory footprint decides how many CTAs an SM holds, and the kernel is latency-bound
template <bool SM, typename PosT>
__global__ void __launch_bounds__(SM ? ORD_THREADS : ORD_THREADS_BIG) k_uset_order(const OrderArgs a) {
    extern __shared__ __align__(16) uint32_t smo[];
    __shared__ uint32_t ws[33];
    __shared__ uint32_t s_pb;
    const int tid = threadIdx.x, nthr = SM ? ORD_THREADS : ORD_THREADS_BIG, lane = tid & 31, wid = tid >> 5, nwarp = nthr >> 5;
    uint32_t *base;
    if constexpr (SM) base = smo;
    else base = a.gscratch + (size_t) blockIdx.x * a.gscratch_stride;
    const int MC = a.m_cap, BC = a.b_cap;
    constexpr PosT P_NONE = (PosT) ~(PosT) 0, P_FLAG = (PosT) ((PosT) 1 << (8 * sizeof(PosT) - 1)), P_MASK = (PosT) (P_FLAG - 1);
    uint32_t *first = base;                        // [BC] first position, then the run start of the bucket
    uint32_t *head = first + BC;                   // [BC] chain head
    uint32_t *cnt = head + BC;                     // [BC] bucket size
    PosT *cur = reinterpret_cast<PosT *>(cnt + BC), *nxtL = cur + MC;  // ping-pong order lists (arrival indices)
    PosT *chain = cur + 2 * MC;                    // next position in the bucket's chain
    PosT *bk = cur + 3 * MC;                       // bucket of position p (top bit: p is the bucket's first position)

    for (;;) {
        __syncthreads();
        if (tid == 0) s_pb = atomicAdd(a.work_counter, 1u);
        __syncthreads();
        const uint32_t pb = s_pb;
        if (pb >= (uint32_t) a.n_prob) break;
        const ProbDesc d = a.prob[pb];
        const int m = d.pad0;  // arrival count = size of the set before the +/- cancellation
        const uint32_t *arr = (d.pol ? a.arrive[1] : a.arrive[0]) + d.off;
        uint32_t *dst = (d.pol ? a.pts[1] : a.pts[0]) + d.off;
        if (m <= 0) continue;

        for (int stage = 0;; ++stage) {
            const int B = (int) c_prime[stage];
            const int prevN = stage ? (int) c_prime[stage - 1] : 0;
            const int newN = min(m, B);
            for (int b = tid; b < B; b += nthr) {
                first[b] = ECB_NONE;
                head[b] = ECB_NONE;
                cnt[b] = 0;
            }
            __syncthreads();
            // A: bucket of every position of S = (old list order) ++ (new arrivals), first position, chains, sizes
            for (int p = tid; p < newN; p += nthr) {
                const uint32_t e = p < prevN ? (uint32_t) cur[p] : (uint32_t) p;
                if (p >= prevN) cur[p] = (PosT) e;
                const uint32_t px = arr[e] & 0x3FFFFFFFu;
                const uint32_t b = bucket_of((ECB_ORD_HTAB && a.htab) ? hash_pixel_tab(px, a.htab) : hash_pixel(px), stage);
                bk[p] = (PosT) b;
                atomicMin(&first[b], (uint32_t) p);
                const uint32_t prev = atomicExch(&head[b], (uint32_t) p);
                chain[p] = prev == ECB_NONE ? P_NONE : (PosT) prev;
                atomicAdd(&cnt[b], 1u);
            }
            __syncthreads();
            // B: run starts = suffix sum of the bucket sizes over the head positions (descending first position).
            // Warp w owns positions [w*L, (w+1)*L), lanes interleaved.
            const int L = (((newN + nwarp - 1) / nwarp) + 31) & ~31;
            const int p0 = wid * L, p1 = min(newN, p0 + L);
            uint32_t wsum = 0;
            for (int c = p0; c < p1; c += 32) {
                const int p = c + lane;
                uint32_t hs = 0;
                if (p < p1) {
                    const PosT b = bk[p];
                    if (first[b] == (uint32_t) p) {
                        hs = cnt[b];
                        bk[p] = (PosT) (b | P_FLAG);
                    }
                }
                wsum += hs;
            }
            for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
            if (lane == 0) ws[wid] = wsum;
            __syncthreads();  // also orders all reads of first[] == p before the overwrite below
            uint32_t carry = 0;
            for (int w = wid + 1; w < nwarp; ++w) carry += ws[w];
            for (int c = ((p1 - p0 + 31) & ~31) - 32 + p0; c >= p0; c -= 32) {
                const int p = c + lane;
                uint32_t hs = 0;
                PosT b = 0;
                if (p < p1) {
                    b = bk[p];
                    if (b & P_FLAG) hs = cnt[b & P_MASK];
                }
                // exclusive suffix sum inside the warp (lanes above)
                uint32_t inc = hs;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_down_sync(0xffffffffu, inc, o);
                    if (lane + o < 32) inc += t;
                }
                if (hs) first[b & P_MASK] = carry + inc - hs;
                carry += __shfl_sync(0xffffffffu, inc, 0);
            }
            __syncthreads();
            // C: place — inside the run, descending position
            for (int p = tid; p < newN; p += nthr) {
                const uint32_t b = bk[p] & P_MASK;
                uint32_t g = 0;
                for (uint32_t q = head[b]; q != ECB_NONE;) {
                    g += q > (uint32_t) p;
                    const PosT nq = chain[q];
                    q = nq == P_NONE ? ECB_NONE : (uint32_t) nq;
                }
                nxtL[first[b] + g] = cur[p];
            }
            __syncthreads();
            PosT *t = cur;
            cur = nxtL;
            nxtL = t;
            if (m <= B) break;
        }
        // survivors of the +/- cancellation (bit 31 of the arrival word), in iteration order
        uint32_t run = 0;
        for (int c0 = 0; c0 < m; c0 += nthr) {
            const int i = c0 + tid;
            uint32_t p = 0;
            bool keep = false;
            if (i < m) {
                p = arr[cur[i]];
                keep = !(p & 0x80000000u);
            }
            uint32_t tot;
            const uint32_t ex = block_excl_scan(keep ? 1u : 0u, ws, &tot);
            if (keep) dst[run + ex] = p;
            run += tot;
        }
    }
}

}  // namespace

int ecb_launch_order(ecb_ctx *ctx, OrderArgs &a, int max_m) {
    if (a.n_prob <= 0 || max_m <= 0) return ECB_OK;
    if (max_m > (int) h_prime[ORD_NSTAGE - 1]) return ecb_fail(ctx, ECB_ERR_UNSUPPORTED, "window with %d distinct pixels", max_m);
    int st = 0;
    while ((int) h_prime[st] < max_m) ++st;
    a.htab = nullptr;
    if (ECB_ORD_HTAB) {  // the coordinate hashes of this sensor, uploaded once per context
        const int nt = std::max(ctx->width, ctx->height);
        if (nt > 0 && ctx->ord_htab_n != nt) {
            std::vector<unsigned long long> tab((size_t) nt);
            for (int v = 0; v < nt; ++v) tab[(size_t) v] = hash_double((double) v) + 0x9e3779b9ull;
            int rc = ecb_reserve(ctx, ctx->ord_htab, (size_t) nt * 8);
            if (rc) return rc;
            ECB_CUDA(ctx, cudaMemcpyAsync(ctx->ord_htab.p, tab.data(), (size_t) nt * 8, cudaMemcpyHostToDevice, ctx->stream));
            ECB_CUDA(ctx, ecb_stream_sync(ctx));
            ctx->ord_htab_n = nt;
        }
        if (ctx->ord_htab_n == nt && nt > 0) a.htab = (const unsigned long long *) ctx->ord_htab.p;
    }
    a.m_cap = (max_m + 3) & ~3;
    a.b_cap = ((int) h_prime[st] + 3) & ~3;
    const bool pos16 = ECB_ORD_POS16 && a.m_cap < 0x7FFF && a.b_cap < 0x7FFF;  // positions, arrival indices and bucket numbers in 15 bits
    const size_t words = (pos16 ? (size_t) 2 : (size_t) 4) * a.m_cap + (size_t) 3 * a.b_cap;
    const size_t limit = (size_t) ctx->smem_optin - 2 * 1024;
    a.arrays_in_smem = words * 4 <= limit;
    const size_t smem = a.arrays_in_smem ? words * 4 : 0;
    void (*kern)(const OrderArgs) = a.arrays_in_smem ? (pos16 ? k_uset_order<true, uint16_t> : k_uset_order<true, uint32_t>)
                                                     : (pos16 ? k_uset_order<false, uint16_t> : k_uset_order<false, uint32_t>);
    ECB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) limit)  /* constant: race-free */);
    int per_sm = 1;
    const int threads = a.arrays_in_smem ? ORD_THREADS : ORD_THREADS_BIG;
    ECB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) per_sm = 1;
    int grid = ctx->sm_count * per_sm;
    if (grid > a.n_prob) grid = a.n_prob;
    if (!a.arrays_in_smem) {
        a.gscratch_stride = words;
        int rc = ecb_reserve(ctx, ctx->scratch, (size_t) grid * words * 4);
        if (rc) return rc;
        a.gscratch = (uint32_t *) ctx->scratch.p;
    }
    ECB_CUDA(ctx, cudaMemsetAsync(a.work_counter, 0, 4, ctx->stream));
    ECB_PROF_BEGIN(ctx, ECB_STAGE_ORDER);
    kern<<<grid, threads, smem, ctx->stream>>>(a);
    ECB_PROF_END(ctx, ECB_STAGE_ORDER);
    ECB_LAUNCHED(ctx);
    return ecb_check(ctx, cudaGetLastError(), "k_uset_order launch");
}
