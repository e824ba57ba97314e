// Cost evaluation of the dynamic-calibration objective on the GPU.
//
//   k_associate_*   EventCalibSpline::optimize association loop (src/EventCalibSpline.cpp:157-192) + findCenter
//                   (include/.../CirclesEventFrame.hpp:50-65): raw event -> nearest keyframe in time -> nearest circle
//                   -> | ||p-c|| - r | < 5 px -> residual record (ordered compaction, two passes)
//   k_prepare       findSpan + dersBasisFuns per residual (core/spline/.../BsplineReal.hpp:208-231,107-145); the basis
//                   is constant over the LM iterations, so it is computed once here
//   k_normal_eq     residual + closed-form Jacobian (ecb_residual.h) per lane, then the per-span Gram matrix
//                   [J | r]^T [J | r] (34 x 34, padded to 40) on the FP64 tensor pipe: mma.sync m8n8k4 f64 (DMMA),
//                   15 upper 8x8 tiles per warp kept in registers, J staged through shared memory transposed
//                   (conflict-free for both the lane-per-residual stores and the fragment loads).
//                   B200 measured: DMMA 37.0 TFLOP/s == DFMA 36.4 TFLOP/s (profiles/r1_fp64_peak.txt), so the
//                   tensor path costs nothing in peak and removes the shared-memory/issue bottleneck of a
//                   CUDA-core register-tiled SYRK.  tcgen05 has no f64 kind.
//   k_reduce_*      fixed-order reduction of the per-work-item partials -> run-to-run identical J^T J / J^T r / cost
//   k_cost          cost-only evaluation (LM step acceptance)
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "ecb_common.cuh"
#include "ecb_residual_so3.h"

namespace {

constexpr int NE_THREADS = 256;            // 8 warps
constexpr int NE_WARPS = NE_THREADS / 32;
constexpr int TILE_ROWS = 34, TILE_LD = 36;  // [param | r][residual], LD % 16 == 4 -> conflict-free fragment loads
constexpr int N_TILES = 10;                 // upper 8x8 DMMA tiles of the leading 32x32 block of the 34x34 Gram matrix
// per work item: 10 tiles | row 32 (32) | row 33 = J^T r (32) | H[32][32], g[32], sum r^2, cost
constexpr int PART_E32 = N_TILES * 64, PART_E33 = PART_E32 + 32, PART_SC = PART_E33 + 32, PART_STRIDE = PART_SC + 8;
constexpr int CHUNK = 2048;                 // residuals per work item
constexpr int OUT_STRIDE = 33 * 33 + 33;    // per span: H (full symmetric) | g

struct CostState {
    int n_splines = 0, total_cp = 0, total_spans = 0;
    std::vector<int> n_cp, cp_off, span_off, knot_off;
    std::vector<double> knots;
    double radius = 1.75, huber = 0.35;
    bool use_circ = false;  // residuals carry a landmark index (association path) instead of explicit landmark coordinates
    int so3 = 0;  // rotation model: 0 normalised quaternion spline (useSO3: 0), 1 cumulative SO(3) spline (useSO3: 1)
    int64_t n_res = 0;
    int n_items = 0;
    DevBuf d_knots, d_knot_off, d_ncp, d_cp_off, d_span_off;
    DevBuf obs, lm, tt, spl, basis, cp0, span;        // residual records (SoA)
    DevBuf span_start, items, part, out, params, cost_part, flags;
    DevBuf ev_flag, ev_cnt, ev_tag, kf_t, kf_circ, kf_c32, lm_tab, sel_event, sel_circle;
    std::vector<int64_t> h_span_start;
    uint64_t xch_last_epoch = 0;   // ecb_cost_normal_eq_exchange: last epoch sent / generation of the caller's sequence
    uint32_t xch_generation = 0;
};

CostState *state(ecb_ctx *ctx) {
    if (!ctx->cost) ctx->cost = new CostState();
    return (CostState *) ctx->cost;
}

// ------------------------------------------------------------------------------------ prepare ----
__device__ __forceinline__ int find_span_dev(const double *knots, int nk, double u) {  // BsplineReal.hpp:208-231
    const int degree = 3;
    const int n = nk - 2 - degree;
    if (u == knots[n + 1]) return n;
    int low = degree, high = n + 1, mid = (low + high) / 2;
    while (u < knots[mid] || u >= knots[mid + 1]) {
        if (u < knots[mid]) high = mid; else low = mid;
        mid = (low + high) / 2;
    }
    return mid;
}

// same span (the one with knots[mid] <= u < knots[mid+1] is unique) from a proportional guess and a short walk: the
// reference's interior knots are nearly uniform (NURBS-book eq. 9.68), so the dependent-load chain of the bisection goes away
__device__ __forceinline__ int find_span_guess(const double *knots, int nk, double u) {
    const int degree = 3;
    const int n = nk - 2 - degree;
    if (u == knots[n + 1]) return n;
    const double k0 = knots[degree], k1 = knots[n + 1];
    int mid = degree + (int) ((u - k0) / (k1 - k0) * (double) (n + 1 - degree));
    mid = max(degree, min(n, mid));
    for (int s = 0; s < 8 && mid > degree && u < knots[mid]; ++s) --mid;
    for (int s = 0; s < 8 && mid < n && u >= knots[mid + 1]; ++s) ++mid;
    if (u < knots[mid] || u >= knots[mid + 1]) return find_span_dev(knots, nk, u);
    return mid;
}

__device__ __forceinline__ void basis_dev(const double *knots, int span, double u, double *N) {  // BsplineReal.hpp:107-145
    double ndu[4][4], left[4], right[4];
    ndu[0][0] = 1;
#pragma unroll
    for (int j = 1; j <= 3; j++) {
        left[j] = u - knots[span + 1 - j];
        right[j] = knots[span + j] - u;
        double saved = 0.0;
#pragma unroll
        for (int r = 0; r < j; ++r) {
            ndu[j][r] = right[r + 1] + left[j - r];
            double temp = __ddiv_rn(ndu[r][j - 1], ndu[j][r]);
            ndu[r][j] = __dadd_rn(saved, __dmul_rn(right[r + 1], temp));
            saved = __dmul_rn(left[j - r], temp);
        }
        ndu[j][j] = saved;
    }
#pragma unroll
    for (int j = 0; j <= 3; j++) N[j] = ndu[j][3];
}

__global__ void k_prepare(const double *__restrict__ t, const int *__restrict__ spl, int64_t n, const double *__restrict__ knots,
                          const int *__restrict__ knot_off, const int *__restrict__ ncp, const int *__restrict__ cp_off,
                          const int *__restrict__ span_off, double *__restrict__ basis, int *__restrict__ cp0,
                          int *__restrict__ span, uint32_t *__restrict__ flags) {
    const int64_t k = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int s = spl[k];
    const double *kn = knots + knot_off[s];
    const int nk = ncp[s] + 4;
    const double u = t[k];
    if (!(u >= kn[0] && u <= kn[nk - 1])) {  // outside the spline: the reference never creates such a block
        atomicOr(flags, 1u);
        span[k] = 0;
        cp0[k] = 0;
        return;
    }
    const int sp = find_span_dev(kn, nk, u);
    double N[4];
    basis_dev(kn, sp, u, N);
    reinterpret_cast<double4 *>(basis)[k] = make_double4(N[0], N[1], N[2], N[3]);
    cp0[k] = cp_off[s] + sp - 3;
    const int gs = span_off[s] + sp - 3;
    span[k] = gs;
    if (k > 0) {  // records must be ordered by (spline, time) so that spans are contiguous
        const int sprev = spl[k - 1];
        if (sprev > s || (sprev == s && t[k - 1] > u)) atomicOr(flags, 2u);
    }
}

__global__ void k_span_start(const int *__restrict__ span, int64_t n, int n_spans, int64_t *__restrict__ start) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_spans) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t m = (lo + hi) >> 1;
        if (span[m] < s) lo = m + 1; else hi = m;
    }
    start[s] = lo;
}

// ---------------------------------------------------------------------------------- normal eq ----
struct Item {
    int64_t begin, end;
    int span, pad;
};

struct NeArgs {
    const double *obs, *lm, *basis;
    const int *circ;        // landmark index per residual into lm_tab (association path), or null: explicit `lm` per residual
    const double *lm_tab;
    const int *cp0;
    const Item *items;
    int n_items;
    const double *params;  // intr[9] | rot[4*C] | trans[3*C]
    int total_cp;
    double radius, huber;
    double *part;
    const int *go;         // device flag (may be null): 0 = the launch is a no-op (device-side LM loop: step rejected / finished)
};

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int MINB, bool SO3>
__global__ void __launch_bounds__(NE_THREADS, MINB) k_normal_eq(const NeArgs a) {
    extern __shared__ __align__(16) double sm_tiles[];
    if (a.go && *a.go == 0) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double *tile = sm_tiles + (size_t) wid * TILE_ROWS * TILE_LD;
    const int g = lane >> 2, t = lane & 3;
    const double *intr = a.params, *rot = a.params + 9, *trans = a.params + 9 + 4 * (size_t) a.total_cp;
    const int gw = blockIdx.x * NE_WARPS + wid, nw = gridDim.x * NE_WARPS;
    for (int it = gw; it < a.n_items; it += nw) {
        const Item item = a.items[it];
        // Gram matrix of [J | r] (34 columns): the leading 32x32 block on the FP64 tensor pipe (10 upper 8x8 tiles),
        // rows 32 (last parameter) and 33 (r) with plain DFMAs — padding 34 -> 40 would waste 5 of 15 DMMA tiles
        double acc[N_TILES][2];
#pragma unroll
        for (int i = 0; i < N_TILES; ++i) acc[i][0] = acc[i][1] = 0.0;
        double e32 = 0.0, e33 = 0.0;            // lane j: sum_k J_k[32] J_k[j], sum_k r_k J_k[j]
        double s3232 = 0.0, s3233 = 0.0, s3333 = 0.0, cost = 0.0;  // per-lane partials over its own residuals
        for (int64_t base = item.begin; base < item.end; base += 32) {
            const int64_t k = base + lane;
            double res = 0.0;
            __syncwarp();  // the previous group's fragment loads are done
            if (k < item.end) {
                const int c0 = a.cp0[k];
                const double4 b4 = reinterpret_cast<const double4 *>(a.basis)[k];
                const double b[4] = {b4.x, b4.y, b4.z, b4.w};
                const double2 o = reinterpret_cast<const double2 *>(a.obs)[k];
                const double *L = a.circ ? a.lm_tab + 3 * a.circ[k] : a.lm + 3 * k;
                const double Lx = L[0], Ly = L[1], Lz = L[2];
                // the Jacobian is scattered straight into column `lane` of the shared tile (no register array)
                const EcbResidualOut r =
                    SO3 ? ecb_residual_so3<true, TILE_LD>(intr, rot + 4 * (size_t) c0, trans + 3 * (size_t) c0, b, o.x, o.y,
                                                          Lx, Ly, Lz, a.radius, a.huber,
                                                          tile + lane)
                        : ecb_residual<true, TILE_LD>(intr, rot + 4 * (size_t) c0, trans + 3 * (size_t) c0, b, o.x, o.y,
                                                      Lx, Ly, Lz, a.radius, a.huber,
                                                      tile + lane);
                res = r.res;
                cost += r.cost;
            } else {
#pragma unroll
                for (int c = 0; c < 33; ++c) tile[c * TILE_LD + lane] = 0.0;
            }
            tile[33 * TILE_LD + lane] = res;
            __syncwarp();
            {
                const double x32 = tile[32 * TILE_LD + lane];
                s3232 += x32 * x32;
                s3233 += x32 * res;
                s3333 += res * res;
            }
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                double f[4];
#pragma unroll
                for (int G = 0; G < 4; ++G) f[G] = tile[(8 * G + g) * TILE_LD + 4 * ks + t];
                int ti = 0;
#pragma unroll
                for (int I = 0; I < 4; ++I)
#pragma unroll
                    for (int Jt = I; Jt < 4; ++Jt) {
                        dmma(acc[ti][0], acc[ti][1], f[I], f[Jt]);
                        ++ti;
                    }
                // rows 32 / 33: lane j accumulates column j; residual index rotated per lane group -> conflict-free loads
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int kk = (4 * ks + q + (lane >> 2)) & 31;
                    const double aj = tile[lane * TILE_LD + kk];
                    e32 += aj * tile[32 * TILE_LD + kk];
                    e33 += aj * tile[33 * TILE_LD + kk];
                }
            }
        }
        // partial of this work item
        double *p = a.part + (size_t) it * PART_STRIDE;
#pragma unroll
        for (int ti = 0; ti < N_TILES; ++ti) reinterpret_cast<double2 *>(p + ti * 64)[lane] = make_double2(acc[ti][0], acc[ti][1]);
        p[PART_E32 + lane] = e32;
        p[PART_E33 + lane] = e33;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cost += __shfl_xor_sync(0xffffffffu, cost, o);
            s3232 += __shfl_xor_sync(0xffffffffu, s3232, o);
            s3233 += __shfl_xor_sync(0xffffffffu, s3233, o);
            s3333 += __shfl_xor_sync(0xffffffffu, s3333, o);
        }
        if (lane == 0) {
            p[PART_SC + 0] = s3232;
            p[PART_SC + 1] = s3233;
            p[PART_SC + 2] = s3333;
            p[PART_SC + 3] = cost;
        }
    }
}

// out[s] = sum over the span's work items, fixed order; tiles -> full symmetric 33x33 + gradient
__global__ void k_reduce_spans(const double *__restrict__ part, const int *__restrict__ item_start, int n_spans,
                               double *__restrict__ out, const int *__restrict__ go = nullptr) {
    const int s = blockIdx.x;
    if (s >= n_spans || (go && *go == 0)) return;
    const int i0 = item_start[s], i1 = item_start[s + 1];
    double *o = out + (size_t) s * OUT_STRIDE;
    for (int e = threadIdx.x; e < PART_SC + 2; e += blockDim.x) {
        double v = 0.0;
        for (int it = i0; it < i1; ++it) v += part[(size_t) it * PART_STRIDE + e];
        if (e < PART_E32) {
            const int ti = e >> 6, r = (e >> 3) & 7, c = e & 7;
            int I = 0, rem = ti;  // tile index -> (I, Jt), I <= Jt over 4 groups
            while (rem >= 4 - I) {
                rem -= 4 - I;
                ++I;
            }
            const int Jt = I + rem;
            const int i = 8 * I + r, j = 8 * Jt + c;
            if (I != Jt || i <= j) {
                o[i * 33 + j] = v;
                o[j * 33 + i] = v;
            }
        } else if (e < PART_E33) {
            const int j = e - PART_E32;
            o[32 * 33 + j] = v;
            o[j * 33 + 32] = v;
        } else if (e < PART_SC) {
            o[1089 + (e - PART_E33)] = v;
        } else if (e == PART_SC) {
            o[32 * 33 + 32] = v;
        } else {
            o[1089 + 32] = v;
        }
    }
}

// ---------------------------------------------------------------- fused reduce + inter-GPU exchange ----
// Multi-GPU normal equations without a separate collective: the span reduction of every rank stores its result straight
// into a slot of EVERY peer's receive buffer over NVLink (peer-mapped memory), a one-warp kernel publishes
// (epoch, span range, cost) to the peers, and each rank sums the slots in rank order once all have arrived — a one-shot
// all-reduce whose send side is the epilogue of k_reduce_spans.  Deterministic (fixed rank order), identical on every
// rank, double-buffered by epoch parity (a rank can be at most one exchange ahead of a peer).
//
// receive buffer of one rank:  [parity 0 | parity 1],  parity block = n_ranks slots,  slot = header (8 doubles:
// epoch as u64, first span, last span, cost) + total_spans * OUT_STRIDE doubles (only [first, last] are written).
constexpr int XH = 8;  // header doubles

__host__ __device__ inline size_t xslot_doubles(int total_spans) { return (size_t) XH + (size_t) total_spans * OUT_STRIDE; }

struct XPeers {
    double *base[ECB_MAX_PEERS];
    int n_ranks, rank;
};

// Where an exchange lives in the receive buffers: the epoch comes from the host (ecb_cost_normal_eq_exchange) or from a device
// counter (device-side LM loop, ecb_lmdev.cu: the host enqueues many iterations ahead and does not know which of them exchange).
struct XCtl {
    const int *go;                        // device flag (may be null): 0 = no-op
    const unsigned long long *d_epoch;    // device epoch counter, or null: use `epoch`
    unsigned long long epoch;
    size_t slot;                          // doubles per slot
    int n_ranks, rank;
    __device__ __forceinline__ unsigned long long e() const { return d_epoch ? *d_epoch : epoch; }
    __device__ __forceinline__ size_t parity_off() const { return (size_t) (e() & 1ull) * (size_t) n_ranks * slot; }
    __device__ __forceinline__ size_t my_slot_off() const { return parity_off() + (size_t) rank * slot; }
};

// k_reduce_spans whose stores go to the slot of this rank in every peer's buffer
__global__ void k_reduce_spans_push(const double *__restrict__ part, const int *__restrict__ item_start, int span_lo, int span_hi,
                                    XPeers peers, XCtl x) {
    const int s = span_lo + blockIdx.x;
    if (s > span_hi || (x.go && *x.go == 0)) return;
    const size_t slot_off = x.my_slot_off();
    const int i0 = item_start[s], i1 = item_start[s + 1];
    for (int e = threadIdx.x; e < PART_SC + 2; e += blockDim.x) {
        double v = 0.0;
        for (int it = i0; it < i1; ++it) v += part[(size_t) it * PART_STRIDE + e];
        int a0 = -1, a1 = -1;  // the (at most two, symmetric) destinations inside the span block
        if (e < PART_E32) {
            const int ti = e >> 6, r = (e >> 3) & 7, c = e & 7;
            int I = 0, rem = ti;
            while (rem >= 4 - I) {
                rem -= 4 - I;
                ++I;
            }
            const int Jt = I + rem;
            const int i = 8 * I + r, j = 8 * Jt + c;
            if (I != Jt || i <= j) {
                a0 = i * 33 + j;
                a1 = j * 33 + i;
            }
        } else if (e < PART_E33) {
            const int j = e - PART_E32;
            a0 = 32 * 33 + j;
            a1 = j * 33 + 32;
        } else if (e < PART_SC) {
            a0 = 1089 + (e - PART_E33);
        } else if (e == PART_SC) {
            a0 = 32 * 33 + 32;
        } else {
            a0 = 1089 + 32;
        }
        if (a0 < 0) continue;
        for (int p = 0; p < peers.n_ranks; ++p) {
            double *o = peers.base[p] + slot_off + XH + (size_t) s * OUT_STRIDE;
            o[a0] = v;
            if (a1 >= 0) o[a1] = v;
        }
    }
}

// after the pushes (stream order): publish this rank's header to every peer
__global__ void k_exchange_signal(XPeers peers, XCtl x, int span_lo, int span_hi, const double *__restrict__ cost) {
    const int p = threadIdx.x;
    if (p >= peers.n_ranks || (x.go && *x.go == 0)) return;
    double *h = peers.base[p] + x.my_slot_off();
    h[1] = (double) span_lo;
    h[2] = (double) span_hi;
    h[3] = *cost;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(h) = x.e();
}

// one warp: wait until every rank's header of this parity carries the epoch.  The spin is bounded (ECB_EXCHANGE_TIMEOUT_S
// seconds, default 30: lazy module loads or host pauses of a peer process must not trip it) and a timeout raises a sticky
// error bit per rank that k_exchange_sum turns into a NaN cost, so no caller can mistake a partial sum for a result.
__global__ void k_exchange_wait(const double *__restrict__ recv, XCtl x, long long timeout_clocks, unsigned *err) {
    const int r = threadIdx.x;
    if (r >= x.n_ranks || (x.go && *x.go == 0)) return;
    const unsigned long long epoch = x.e();
    const volatile unsigned long long *f =
        reinterpret_cast<const volatile unsigned long long *>(recv + x.parity_off() + (size_t) r * x.slot);
    const long long t0 = clock64();
    while (*f != epoch) {
        if (clock64() - t0 > timeout_clocks) {
            atomicOr(err, 1u << r);
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

// out[s] = sum over the ranks whose range holds s, in rank order; out[n_spans*OUT_STRIDE] = sum of the costs
__global__ void k_exchange_sum(const double *__restrict__ recv, XCtl x, int n_spans, double *__restrict__ out,
                               const unsigned *__restrict__ err) {
    const int s = blockIdx.x;
    if (x.go && *x.go == 0) return;
    const size_t parity_off = x.parity_off(), slot_stride = x.slot;
    const int n_ranks = x.n_ranks;
    if (s == n_spans) {
        if (threadIdx.x == 0) {
            double c = 0.0;
            for (int r = 0; r < n_ranks; ++r) c += recv[parity_off + (size_t) r * slot_stride + 3];
            if (err && *err) c = __longlong_as_double(0x7FF8000000000000ll);  // a rank never arrived: no result
            out[(size_t) n_spans * OUT_STRIDE] = c;
            out[(size_t) n_spans * OUT_STRIDE + 1] = 0.0;
        }
        return;
    }
    for (int e = threadIdx.x; e < OUT_STRIDE; e += blockDim.x) {
        double v = 0.0;
        for (int r = 0; r < n_ranks; ++r) {
            const double *slot = recv + parity_off + (size_t) r * slot_stride;
            if (s >= (int) slot[1] && s <= (int) slot[2]) v += slot[XH + (size_t) s * OUT_STRIDE + e];
        }
        out[(size_t) s * OUT_STRIDE + e] = v;
    }
}

// Scalar all-reduce over the same peer buffers (device-side LM loop: the candidate cost of every iteration): rank r stores
// (value, epoch) into slot r of every peer's scalar area, waits for all slots of this epoch and adds them in rank order.
// One warp; *value is replaced by the sum.  The scalar area follows the two parity blocks of the span slots.
__global__ void k_exchange_scalar(XPeers peers, XCtl x, size_t scalar_off, double *value, long long timeout_clocks, unsigned *err) {
    const int p = threadIdx.x;
    if (x.go && *x.go == 0) return;
    const unsigned long long epoch = x.e();
    const size_t par = scalar_off + (size_t) (epoch & 1ull) * 2 * (size_t) x.n_ranks;
    if (p < x.n_ranks) {
        double *h = peers.base[p] + par + 2 * (size_t) x.rank;
        h[1] = *value;
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(h) = epoch;
        const volatile unsigned long long *f = reinterpret_cast<const volatile unsigned long long *>(peers.base[x.rank] + par + 2 * (size_t) p);
        const long long t0 = clock64();
        while (*f != epoch) {
            if (clock64() - t0 > timeout_clocks) {
                atomicOr(err, 1u << p);
                break;
            }
            __nanosleep(100);
        }
        __threadfence_system();
    }
    __syncwarp();
    if (p == 0) {
        double c = 0.0;
        for (int r = 0; r < x.n_ranks; ++r) c += *reinterpret_cast<const volatile double *>(peers.base[x.rank] + par + 2 * (size_t) r + 1);
        if (*err) c = __longlong_as_double(0x7FF8000000000000ll);
        *value = c;
    }
}

__global__ void k_reduce_cost(const double *__restrict__ part, int n, int stride, int offset, double *__restrict__ out,
                              const int *__restrict__ go = nullptr) {
    // single block, fixed order: thread-strided partial sums, then a shared-memory tree
    __shared__ double sh[256];
    if (go && *go == 0) return;
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) v += part[(size_t) i * stride + offset];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}

template <bool SO3, int MINB>
__global__ void __launch_bounds__(256, MINB) k_cost(const double *__restrict__ obs, const double *__restrict__ lm,
                                             const int *__restrict__ circ, const double *__restrict__ lm_tab,
                                             const double *__restrict__ basis, const int *__restrict__ cp0, int64_t n,
                                             const double *__restrict__ params, int total_cp, double radius, double huber,
                                             double *__restrict__ block_part, const int *__restrict__ go = nullptr) {
    __shared__ double sh[256];
    if (go && *go == 0) return;
    const double *intr = params, *rot = params + 9, *trans = params + 9 + 4 * (size_t) total_cp;
    double c = 0.0;
    for (int64_t k = (int64_t) blockIdx.x * 256 + threadIdx.x; k < n; k += (int64_t) gridDim.x * 256) {
        const int c0 = cp0[k];
        const double4 b4 = reinterpret_cast<const double4 *>(basis)[k];
        const double b[4] = {b4.x, b4.y, b4.z, b4.w};
        const double2 o = reinterpret_cast<const double2 *>(obs)[k];
        const double *L = circ ? lm_tab + 3 * circ[k] : lm + 3 * k;
        const double Lx = L[0], Ly = L[1], Lz = L[2];
        c += SO3 ? ecb_residual_so3<false>(intr, rot + 4 * (size_t) c0, trans + 3 * (size_t) c0, b, o.x, o.y, Lx, Ly, Lz, radius,
                                           huber, nullptr).cost
                 : ecb_residual<false>(intr, rot + 4 * (size_t) c0, trans + 3 * (size_t) c0, b, o.x, o.y, Lx, Ly, Lz, radius, huber,
                                       nullptr).cost;
    }
    sh[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) block_part[blockIdx.x] = sh[0];
}

// -------------------------------------------------------------------------------- association ----
struct AssocArgs {
    const double *ev_t;
    const uint32_t *ev_xyp;
    int64_t n_ev;
    const double *knots;
    const int *knot_off, *ncp;
    int n_splines;
    const double *kf_t, *kf_circ, *lm_tab;
    const float2 *kf_c32;  // FP32 copy of the circle centres, row stride n_circ32 (even), absent / padding = far away
    int K, n_circ, n_circ32;
    double gate2;  // (5 step)^2
    // fused k_prepare (spline ranges sorted and disjoint): span / basis / first control point written with the record
    const int *cp_off, *span_off;
    double *basis;
    int *cp0, *span;
};

// returns the spline index and the circle id (or -1) of one raw event
__device__ __forceinline__ int assoc_one(const AssocArgs &a, int64_t i, int *spline_out) {
    const double u = a.ev_t[i];
    int s = -1;
    for (int q = 0; q < a.n_splines; ++q) {
        const double *kn = a.knots + a.knot_off[q];
        if (u >= kn[0] && u <= kn[a.ncp[q] + 3]) {
            s = q;
            break;
        }
    }
    if (s < 0) return -1;
    *spline_out = s;
    int lo = 0, hi = a.K;  // lower_bound over keyframe times
    while (lo < hi) {
        const int m = (lo + hi) >> 1;
        if (a.kf_t[m] < u) lo = m + 1; else hi = m;
    }
    int best = lo < a.K ? lo : a.K - 1;
    if (lo > 0 && (lo >= a.K || (u - a.kf_t[lo - 1]) <= (a.kf_t[lo] - u))) best = lo - 1;
    const double dt = u - a.kf_t[best];
    if (!(__dmul_rn(dt, dt) < a.gate2)) return -1;
    const uint32_t e = a.ev_xyp[i];
    if (e & ECB_PIX_INVALID) return -1;
    const double x = (double) ECB_PIX_X(e), y = (double) ECB_PIX_Y(e);
    const double *c = a.kf_circ + (size_t) best * a.n_circ * 3;
    int bi = -1;
    double bd = 0.0;
    // Nearest centre: FP32 preselection over all circles with sortable keys (distance bits | circle index), keeping the two
    // smallest.  When the runner-up is clearly farther than the winner the exact FP64 distance is evaluated for the winner
    // alone; near-ties (and nothing else) take the exact FP64 loop, so the result is the exact arg-min either way.
    {
        const float xf = (float) ECB_PIX_X(e), yf = (float) ECB_PIX_Y(e);  // pixel coordinates are exact in FP32
        const float4 *c4 = reinterpret_cast<const float4 *>(a.kf_c32 + (size_t) best * a.n_circ32);
        uint32_t k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
#pragma unroll 6
        for (int q = 0; q < a.n_circ32; q += 2) {
            const float4 cc = c4[q >> 1];
            const float dx0 = xf - cc.x, dy0 = yf - cc.y, dx1 = xf - cc.z, dy1 = yf - cc.w;
            const uint32_t ka = (__float_as_uint(fmaf(dx0, dx0, dy0 * dy0)) & 0xFFFFFF00u) | (uint32_t) q;
            const uint32_t kb = (__float_as_uint(fmaf(dx1, dx1, dy1 * dy1)) & 0xFFFFFF00u) | (uint32_t) (q + 1);
            k2 = min(k2, max(ka, k1));
            k1 = min(k1, ka);
            k2 = min(k2, max(kb, k1));
            k1 = min(k1, kb);
        }
        const float d1 = __uint_as_float(k1 & 0xFFFFFF00u), d2 = __uint_as_float(k2 | 0xFFu);
        if (d2 > d1 * 1.0001f + 0.02f) {  // unambiguous (FP32 error <= 2e-7 d + 3e-4 |d|^(1/2), key quantisation 1.5e-5 d)
            bi = (int) (k1 & 0xFFu);
            if (bi >= a.n_circ || c[3 * bi + 2] < 0) return -1;  // only absent circles
            const double dx = x - c[3 * bi], dy = y - c[3 * bi + 1];
            bd = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        } else {
            for (int q = 0; q < a.n_circ; ++q) {
                if (c[3 * q + 2] < 0) continue;
                const double dx = x - c[3 * q], dy = y - c[3 * q + 1];
                const double dd = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                if (bi < 0 || dd < bd) {
                    bi = q;
                    bd = dd;
                }
            }
        }
    }
    if (bi < 0) return -1;
    if (!(fabs(sqrt(bd) - c[3 * bi + 2]) < 5.0)) return -1;
    return bi;
}

constexpr int AS_THREADS = 256;

// FP32 copy of the circle centres for the preselection in assoc_one: rows padded to an even count; absent circles (r < 0)
// and padding sit at (3e18, 3e18) so they never win against a real circle
__global__ void k_circ32(const double *__restrict__ c, int K, int n_circ, int n_circ32, float2 *__restrict__ o) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * n_circ32) return;
    const int k = i / n_circ32, q = i - k * n_circ32;
    float2 v = make_float2(3e18f, 3e18f);
    if (q < n_circ) {
        const double *p = c + ((size_t) k * n_circ + q) * 3;
        if (!(p[2] < 0)) v = make_float2((float) p[0], (float) p[1]);
    }
    o[i] = v;
}

// nearest key frame in time, ties -> the earlier frame (1-NN over the stamps, EventCalibSpline.cpp:166)
__device__ __forceinline__ int nearest_kf(const double *kf_t, int K, double u) {
    int lo = 0, hi = K;
    while (lo < hi) {
        const int m = (lo + hi) >> 1;
        if (kf_t[m] < u) lo = m + 1; else hi = m;
    }
    int best = lo < K ? lo : K - 1;
    if (lo > 0 && (lo >= K || (u - kf_t[lo - 1]) <= (kf_t[lo] - u))) best = lo - 1;
    return best;
}

// pass 1: decide every event once, remember the decision in a 16-bit tag (0 = no residual, else spline<<8 | circle+1)
// (ncu, profiles/r1e: the 36 FP64 distance evaluations were 71 % of this kernel's instructions, issue-bound at 72 % — hence the
//  FP32 preselection with sortable keys in assoc_one; an earlier shared-memory variant of the FP64 loop had measured slower)
__global__ void __launch_bounds__(AS_THREADS) k_assoc_count(const AssocArgs a, uint16_t *__restrict__ tag,
                                                           uint32_t *__restrict__ block_cnt) {
    __shared__ uint32_t ws[33];
    const int64_t i = (int64_t) blockIdx.x * AS_THREADS + threadIdx.x;
    int s = 0;
    const int bi = i < a.n_ev ? assoc_one(a, i, &s) : -1;
    if (i < a.n_ev) tag[i] = bi >= 0 ? (uint16_t) ((s << 8) | (bi + 1)) : (uint16_t) 0;
    uint32_t tot;
    block_excl_scan(bi >= 0 ? 1u : 0u, ws, &tot);
    if (threadIdx.x == 0) block_cnt[blockIdx.x] = tot;
}

__global__ void k_scan_blocks(uint32_t *cnt, int n, int64_t *off, int64_t *total) {  // single block: one segment per thread
    __shared__ uint32_t ws[33];
    const int per = (n + (int) blockDim.x - 1) / (int) blockDim.x;
    const int b = min(n, (int) threadIdx.x * per), e = min(n, b + per);
    uint32_t sum = 0;
    for (int i = b; i < e; ++i) sum += cnt[i];
    uint32_t tot;
    int64_t run = block_excl_scan(sum, ws, &tot);  // the per-launch total (tagged events) fits 32 bits: n_events < 2^32
    for (int i = b; i < e; ++i) {
        off[i] = run;
        run += cnt[i];
    }
    if (threadIdx.x == 0) *total = tot;
}

// pass 2: ordered compaction of the tagged events into residual records (pure streaming)
__global__ void __launch_bounds__(AS_THREADS) k_assoc_write(const AssocArgs a, const uint16_t *__restrict__ tag,
                                                           const int64_t *__restrict__ block_off,
                                                           double *__restrict__ obs, double *__restrict__ lm,
                                                           double *__restrict__ tt, int *__restrict__ spl,
                                                           int64_t *__restrict__ sel_event, int *__restrict__ sel_circle) {
    __shared__ uint32_t ws[33];
    const int64_t i = (int64_t) blockIdx.x * AS_THREADS + threadIdx.x;
    const uint32_t tg = i < a.n_ev ? tag[i] : 0u;
    uint32_t tot;
    const uint32_t ex = block_excl_scan(tg ? 1u : 0u, ws, &tot);
    if (tg) {
        const int bi = (int) (tg & 0xFF) - 1, s = (int) (tg >> 8);
        const int64_t k = block_off[blockIdx.x] + ex;
        const uint32_t e = a.ev_xyp[i];
        reinterpret_cast<double2 *>(obs)[k] = make_double2((double) ECB_PIX_X(e), (double) ECB_PIX_Y(e));
        const double u = a.ev_t[i];
        sel_event[k] = i;
        sel_circle[k] = bi;  // the evaluation kernels read the landmark from lm_tab[circle]
        if (!a.basis) {      // unfused path: k_prepare needs time and spline, the kernels the explicit landmark
            lm[3 * k] = a.lm_tab[3 * bi];
            lm[3 * k + 1] = a.lm_tab[3 * bi + 1];
            lm[3 * k + 2] = a.lm_tab[3 * bi + 2];
            tt[k] = u;
            spl[k] = s;
        }
        if (a.basis) {  // what k_prepare would compute (findSpan / dersBasisFuns, EventCalibSpline.cpp:173-179)
            const double *kn = a.knots + a.knot_off[s];
            const int sp = find_span_guess(kn, a.ncp[s] + 4, u);
            double N[4];
            basis_dev(kn, sp, u, N);
            reinterpret_cast<double4 *>(a.basis)[k] = make_double4(N[0], N[1], N[2], N[3]);
            a.cp0[k] = a.cp_off[s] + sp - 3;
            a.span[k] = a.span_off[s] + sp - 3;
        }
    }
}

int prepare_records(ecb_ctx *ctx, CostState *st, bool have_basis = false) {
    int rc;
    const int64_t n = st->n_res;
    const size_t nn = (size_t) std::max<int64_t>(n, 1);
    if ((rc = ecb_reserve(ctx, st->basis, nn * 32))) return rc;
    if ((rc = ecb_reserve(ctx, st->cp0, nn * 4))) return rc;
    if ((rc = ecb_reserve(ctx, st->span, nn * 4))) return rc;
    if ((rc = ecb_reserve(ctx, st->flags, 16))) return rc;
    if ((rc = ecb_reserve(ctx, st->span_start, (size_t) (st->total_spans + 2) * 8))) return rc;
    ECB_CUDA(ctx, cudaMemsetAsync(st->flags.p, 0, 16, ctx->stream));
    st->h_span_start.assign((size_t) st->total_spans + 1, 0);
    st->n_items = 0;
    if (n == 0) return ECB_OK;
    if (!have_basis)
    k_prepare<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->stream>>>(
        (const double *) st->tt.p, (const int *) st->spl.p, n, (const double *) st->d_knots.p, (const int *) st->d_knot_off.p,
        (const int *) st->d_ncp.p, (const int *) st->d_cp_off.p, (const int *) st->d_span_off.p, (double *) st->basis.p,
        (int *) st->cp0.p, (int *) st->span.p, (uint32_t *) st->flags.p);
    ECB_LAUNCHED(ctx);
    k_span_start<<<(st->total_spans + 1 + 127) / 128, 128, 0, ctx->stream>>>((const int *) st->span.p, n, st->total_spans,
                                                                           (int64_t *) st->span_start.p);
    ECB_LAUNCHED(ctx);
    uint32_t flag = 0;
    ECB_CUDA(ctx, cudaMemcpyAsync((int64_t *) st->span_start.p + st->total_spans + 1, st->flags.p, 4, cudaMemcpyDeviceToDevice,
                                  ctx->stream));  // one staged copy for both
    {
        std::vector<int64_t> tmp((size_t) st->total_spans + 2);
        if ((rc = ecb_d2h(ctx, tmp.data(), st->span_start.p, tmp.size() * 8))) return rc;
        std::copy(tmp.begin(), tmp.begin() + st->total_spans + 1, st->h_span_start.begin());
        flag = (uint32_t) (tmp[(size_t) st->total_spans + 1] & 0xFFFFFFFF);
    }
    if (flag & 1u) return ecb_fail(ctx, ECB_ERR_ARG, "a residual's time stamp lies outside its spline's knot range");
    if (flag & 2u) return ecb_fail(ctx, ECB_ERR_ARG, "residual records must be ordered by (spline, time)");
    // work items: chunks of <= chunk residuals inside one span (a function of the record counts and the SM count only =>
    // deterministic reduction order).  One item is one warp's latency-bound pass (0.45 ms per 2048 residuals on B200), so the
    // chunk is sized to hand every resident warp of k_normal_eq the same number of items: small residual sets (a time slice
    // of the end-to-end pipeline, one rank of eight) then use all warps with shorter items instead of half of them.
    int chunk = CHUNK;
    {
        const int64_t warps = (int64_t) ctx->sm_count * (st->so3 ? 1 : 2) * NE_WARPS;
        int64_t rounds = std::max<int64_t>(1, (n + warps * CHUNK - 1) / (warps * CHUNK));
        chunk = (int) (((n + warps * rounds - 1) / (warps * rounds) + 31) / 32 * 32);
        chunk = std::min(std::max(chunk, 256), CHUNK);
        static const int chunk_cap = getenv("ECB_NE_CHUNK") ? atoi(getenv("ECB_NE_CHUNK")) : 1536;  // shorter items: no partial last round (1.86 -> 1.81 ms)
        if (chunk_cap >= 256 && chunk > chunk_cap) {
            rounds = std::max<int64_t>(1, (n + warps * chunk_cap - 1) / (warps * chunk_cap));
            chunk = std::min((int) (((n + warps * rounds - 1) / (warps * rounds) + 31) / 32 * 32), chunk_cap);
        }
        for (; chunk < CHUNK; chunk += 32) {  // the partial chunks at span ends add items: grow until the count fits
            int64_t cnt = 0;
            for (int s = 0; s < st->total_spans; ++s)
                cnt += (st->h_span_start[(size_t) s + 1] - st->h_span_start[(size_t) s] + chunk - 1) / chunk;
            if (cnt <= warps * rounds) break;
        }
    }
    std::vector<Item> items;
    std::vector<int> item_start((size_t) st->total_spans + 1, 0);
    for (int s = 0; s < st->total_spans; ++s) {
        item_start[(size_t) s] = (int) items.size();
        for (int64_t b = st->h_span_start[(size_t) s]; b < st->h_span_start[(size_t) s + 1]; b += chunk)
            items.push_back(Item{b, std::min<int64_t>(b + chunk, st->h_span_start[(size_t) s + 1]), s, 0});
    }
    item_start[(size_t) st->total_spans] = (int) items.size();
    st->n_items = (int) items.size();
    const size_t ni = std::max<size_t>(items.size(), 1);
    if ((rc = ecb_reserve(ctx, st->items, ni * sizeof(Item) + item_start.size() * 4 + 64))) return rc;
    if ((rc = ecb_reserve(ctx, st->part, ni * PART_STRIDE * 8))) return rc;
    if ((rc = ecb_reserve(ctx, st->out, ((size_t) st->total_spans * OUT_STRIDE + 8) * 8))) return rc;
    if ((rc = ecb_reserve(ctx, st->cost_part, 4096 * 8))) return rc;
    if (!items.empty())
        if ((rc = ecb_h2d(ctx, st->items.p, items.data(), items.size() * sizeof(Item)))) return rc;
    return ecb_h2d(ctx, (char *) st->items.p + ni * sizeof(Item), item_start.data(), item_start.size() * 4);  // staged: no sync needed
}

int upload_params(ecb_ctx *ctx, CostState *st, const double *intr, const double *rot, const double *trans) {
    int rc;
    const size_t C = (size_t) st->total_cp;
    if ((rc = ecb_reserve(ctx, st->params, (9 + 7 * C) * 8))) return rc;
    double *p = (double *) st->params.p;
    if ((rc = ecb_h2d(ctx, p, intr, 72))) return rc;
    if ((rc = ecb_h2d(ctx, p + 9, rot, 32 * C))) return rc;
    return ecb_h2d(ctx, p + 9 + 4 * C, trans, 24 * C);
}

}  // namespace

extern "C" {

void ecb_cost_free(ecb_ctx *ctx) {
    if (!ctx || !ctx->cost) return;
    CostState *st = (CostState *) ctx->cost;
    DevBuf *bufs[] = {&st->d_knots, &st->d_knot_off, &st->d_ncp, &st->d_cp_off, &st->d_span_off, &st->obs, &st->lm, &st->tt,
                      &st->spl, &st->basis, &st->cp0, &st->span, &st->span_start, &st->items, &st->part, &st->out, &st->params,
                      &st->cost_part, &st->flags, &st->ev_flag, &st->ev_cnt, &st->ev_tag, &st->kf_t, &st->kf_circ, &st->kf_c32, &st->lm_tab,
                      &st->sel_event, &st->sel_circle};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    delete st;
    ctx->cost = nullptr;
}

int ecb_cost_setup(ecb_ctx *ctx, int n_splines, const int32_t *n_cp, const double *knots, double circle_radius,
                   double huber_delta) {
    if (!ctx || n_splines < 1 || !n_cp || !knots) return ECB_ERR_ARG;
    cudaSetDevice(ctx->device);
    CostState *st = state(ctx);
    st->n_splines = n_splines;
    st->n_cp.assign(n_cp, n_cp + n_splines);
    st->cp_off.clear();
    st->span_off.clear();
    st->knot_off.clear();
    int co = 0, so = 0, ko = 0;
    for (int s = 0; s < n_splines; ++s) {
        if (n_cp[s] < 4) return ecb_fail(ctx, ECB_ERR_ARG, "spline %d has %d control points (< degree + 1)", s, n_cp[s]);
        st->cp_off.push_back(co);
        st->span_off.push_back(so);
        st->knot_off.push_back(ko);
        co += n_cp[s];
        so += n_cp[s] - 3;
        ko += n_cp[s] + 4;
    }
    st->total_cp = co;
    st->total_spans = so;
    st->knots.assign(knots, knots + ko);
    st->radius = circle_radius;
    st->huber = huber_delta;
    st->n_res = 0;
    st->n_items = 0;
    int rc;
    if ((rc = ecb_reserve(ctx, st->d_knots, (size_t) ko * 8))) return rc;
    if ((rc = ecb_reserve(ctx, st->d_knot_off, (size_t) n_splines * 4))) return rc;
    if ((rc = ecb_reserve(ctx, st->d_ncp, (size_t) n_splines * 4))) return rc;
    if ((rc = ecb_reserve(ctx, st->d_cp_off, (size_t) n_splines * 4))) return rc;
    if ((rc = ecb_reserve(ctx, st->d_span_off, (size_t) n_splines * 4))) return rc;
    ECB_CUDA(ctx, cudaMemcpyAsync(st->d_knots.p, knots, (size_t) ko * 8, cudaMemcpyHostToDevice, ctx->stream));
    ECB_CUDA(ctx, cudaMemcpyAsync(st->d_knot_off.p, st->knot_off.data(), (size_t) n_splines * 4, cudaMemcpyHostToDevice, ctx->stream));
    ECB_CUDA(ctx, cudaMemcpyAsync(st->d_ncp.p, st->n_cp.data(), (size_t) n_splines * 4, cudaMemcpyHostToDevice, ctx->stream));
    ECB_CUDA(ctx, cudaMemcpyAsync(st->d_cp_off.p, st->cp_off.data(), (size_t) n_splines * 4, cudaMemcpyHostToDevice, ctx->stream));
    ECB_CUDA(ctx, cudaMemcpyAsync(st->d_span_off.p, st->span_off.data(), (size_t) n_splines * 4, cudaMemcpyHostToDevice, ctx->stream));
    return ecb_check(ctx, ecb_stream_sync(ctx), "cost setup");
}

int ecb_cost_set_rotation_model(ecb_ctx *ctx, int use_so3) {
    if (!ctx || !ctx->cost) return ecb_fail(ctx, ECB_ERR_STATE, "ecb_cost_setup first");
    if (use_so3 != 0 && use_so3 != 1) return ECB_ERR_ARG;
    ((CostState *) ctx->cost)->so3 = use_so3;
    return ECB_OK;
}

int ecb_cost_layout(ecb_ctx *ctx, int32_t *total_cp, int32_t *total_spans, int64_t *n_residuals, int64_t *out_doubles) {
    if (!ctx || !ctx->cost) return ECB_ERR_STATE;
    CostState *st = (CostState *) ctx->cost;
    if (total_cp) *total_cp = st->total_cp;
    if (total_spans) *total_spans = st->total_spans;
    if (n_residuals) *n_residuals = st->n_res;
    if (out_doubles) *out_doubles = (int64_t) st->total_spans * OUT_STRIDE + 2;
    return ECB_OK;
}

int ecb_cost_set_residuals(ecb_ctx *ctx, const double *obs_xy, const double *lm_xyz, const double *t, const int32_t *spline,
                           int64_t n) {
    if (!ctx || !ctx->cost) return ecb_fail(ctx, ECB_ERR_STATE, "ecb_cost_setup first");
    if (n < 0 || (n > 0 && (!obs_xy || !lm_xyz || !t || !spline))) return ECB_ERR_ARG;
    cudaSetDevice(ctx->device);
    CostState *st = (CostState *) ctx->cost;
    int rc;
    const size_t nn = (size_t) std::max<int64_t>(n, 1);
    if ((rc = ecb_reserve(ctx, st->obs, nn * 16))) return rc;
    if ((rc = ecb_reserve(ctx, st->lm, nn * 24))) return rc;
    if ((rc = ecb_reserve(ctx, st->tt, nn * 8))) return rc;
    if ((rc = ecb_reserve(ctx, st->spl, nn * 4))) return rc;
    if (n > 0) {
        ECB_CUDA(ctx, cudaMemcpyAsync(st->obs.p, obs_xy, (size_t) n * 16, cudaMemcpyHostToDevice, ctx->stream));
        ECB_CUDA(ctx, cudaMemcpyAsync(st->lm.p, lm_xyz, (size_t) n * 24, cudaMemcpyHostToDevice, ctx->stream));
        ECB_CUDA(ctx, cudaMemcpyAsync(st->tt.p, t, (size_t) n * 8, cudaMemcpyHostToDevice, ctx->stream));
        ECB_CUDA(ctx, cudaMemcpyAsync(st->spl.p, spline, (size_t) n * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    st->n_res = n;
    st->use_circ = false;
    return prepare_records(ctx, st);
}

static int cost_associate(ecb_ctx *ctx, const double *kf_time, const double *kf_circles, int n_keyframes, int n_circles,
                          const double *landmarks_xyz, double motion_time_step, int64_t *n_residuals, bool on_device);

int ecb_cost_associate(ecb_ctx *ctx, const double *kf_time, const double *kf_circles, int n_keyframes, int n_circles,
                       const double *landmarks_xyz, double motion_time_step, int64_t *n_residuals) {
    return cost_associate(ctx, kf_time, kf_circles, n_keyframes, n_circles, landmarks_xyz, motion_time_step, n_residuals, false);
}

int ecb_cost_associate_device(ecb_ctx *ctx, const double *d_kf_time, const double *d_kf_circles, int n_keyframes, int n_circles,
                              const double *d_landmarks_xyz, double motion_time_step, int64_t *n_residuals) {
    return cost_associate(ctx, d_kf_time, d_kf_circles, n_keyframes, n_circles, d_landmarks_xyz, motion_time_step, n_residuals, true);
}

static int cost_associate(ecb_ctx *ctx, const double *kf_time, const double *kf_circles, int n_keyframes, int n_circles,
                          const double *landmarks_xyz, double motion_time_step, int64_t *n_residuals, bool on_device) {
    if (!ctx || !ctx->cost) return ecb_fail(ctx, ECB_ERR_STATE, "ecb_cost_setup first");
    if (!kf_time || !kf_circles || !landmarks_xyz || n_keyframes < 1 || n_circles < 1) return ECB_ERR_ARG;
    if (ctx->n_events <= 0) return ecb_fail(ctx, ECB_ERR_STATE, "no events loaded");
    if (ctx->n_events > 0xFFFFFFFFll) return ecb_fail(ctx, ECB_ERR_UNSUPPORTED, "association over more than 2^32-1 events per context");
    cudaSetDevice(ctx->device);
    CostState *st = (CostState *) ctx->cost;
    int rc;
    const int64_t n = ctx->n_events;
    const int nb = (int) ((n + AS_THREADS - 1) / AS_THREADS);
    if (!on_device) {
        if ((rc = ecb_reserve(ctx, st->kf_t, (size_t) n_keyframes * 8))) return rc;
        if ((rc = ecb_reserve(ctx, st->kf_circ, (size_t) n_keyframes * n_circles * 24))) return rc;
    }
    if ((rc = ecb_reserve(ctx, st->lm_tab, (size_t) n_circles * 24))) return rc;
    const int n_circ32 = (n_circles + 1) & ~1;
    if ((rc = ecb_reserve(ctx, st->kf_c32, (size_t) n_keyframes * n_circ32 * 8))) return rc;
    if (n_circles > 254 || st->n_splines > 255) return ecb_fail(ctx, ECB_ERR_UNSUPPORTED, "more than 254 circles or 255 spline segments");
    if ((rc = ecb_reserve(ctx, st->ev_cnt, (size_t) nb * 4 + 16))) return rc;
    if ((rc = ecb_reserve(ctx, st->ev_tag, (size_t) n * 2 + 16))) return rc;
    if ((rc = ecb_reserve(ctx, st->ev_flag, (size_t) nb * 8 + 16))) return rc;
    // key-frame tables: uploaded from host arrays, or used where they lie when the caller keeps them on the device (the
    // landmark table is copied either way: the evaluation kernels read it long after this call)
    const double *d_kf_t = kf_time, *d_kf_circ = kf_circles;
    if (!on_device) {
        if ((rc = ecb_h2d(ctx, st->kf_t.p, kf_time, (size_t) n_keyframes * 8))) return rc;
        if ((rc = ecb_h2d(ctx, st->kf_circ.p, kf_circles, (size_t) n_keyframes * n_circles * 24))) return rc;
        if ((rc = ecb_h2d(ctx, st->lm_tab.p, landmarks_xyz, (size_t) n_circles * 24))) return rc;
        d_kf_t = (const double *) st->kf_t.p;
        d_kf_circ = (const double *) st->kf_circ.p;
    } else {
        ECB_CUDA(ctx, cudaMemcpyAsync(st->lm_tab.p, landmarks_xyz, (size_t) n_circles * 24, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    AssocArgs a;
    a.ev_t = (const double *) ctx->ev_t.p;
    a.ev_xyp = (const uint32_t *) ctx->ev_xyp.p;
    a.n_ev = n;
    a.knots = (const double *) st->d_knots.p;
    a.knot_off = (const int *) st->d_knot_off.p;
    a.ncp = (const int *) st->d_ncp.p;
    a.n_splines = st->n_splines;
    a.kf_t = d_kf_t;
    a.kf_circ = d_kf_circ;
    a.lm_tab = (const double *) st->lm_tab.p;
    a.kf_c32 = (const float2 *) st->kf_c32.p;
    a.K = n_keyframes;
    a.n_circ = n_circles;
    a.n_circ32 = n_circ32;
    a.gate2 = 5 * motion_time_step * 5 * motion_time_step;  // EventCalibSpline.cpp:168
    ECB_PROF_BEGIN(ctx, ECB_STAGE_ASSOC);
    k_circ32<<<(n_keyframes * n_circ32 + 255) / 256, 256, 0, ctx->stream>>>(d_kf_circ, n_keyframes, n_circles,
                                                                           n_circ32, (float2 *) st->kf_c32.p);
    ECB_LAUNCHED(ctx);
    k_assoc_count<<<nb, AS_THREADS, 0, ctx->stream>>>(a, (uint16_t *) st->ev_tag.p, (uint32_t *) st->ev_cnt.p);
    ECB_LAUNCHED(ctx);
    int64_t *d_total = (int64_t *) st->ev_flag.p + nb;
    k_scan_blocks<<<1, 1024, 0, ctx->stream>>>((uint32_t *) st->ev_cnt.p, nb, (int64_t *) st->ev_flag.p, d_total);
    ECB_LAUNCHED(ctx);
    int64_t total = 0;
    if ((rc = ecb_d2h(ctx, &total, d_total, 8))) return rc;
    const size_t nn = (size_t) std::max<int64_t>(total, 1);
    if ((rc = ecb_reserve(ctx, st->obs, nn * 16))) return rc;
    if ((rc = ecb_reserve(ctx, st->lm, nn * 24))) return rc;
    if ((rc = ecb_reserve(ctx, st->tt, nn * 8))) return rc;
    if ((rc = ecb_reserve(ctx, st->spl, nn * 4))) return rc;
    if ((rc = ecb_reserve(ctx, st->sel_event, nn * 8))) return rc;
    if ((rc = ecb_reserve(ctx, st->sel_circle, nn * 4))) return rc;
    // time-sorted events + spline ranges in ascending, disjoint order => records come out ordered by (spline, time): the
    // basis pass (k_prepare) is folded into the record write; otherwise k_prepare runs and checks the order
    bool fused = true;
    {
        // The reference adds one residual block per spline whose [front, back] holds the event (EventCalibSpline.cpp:157-192);
        // assoc_one emits at most one.  The two agree while the ranges are disjoint — the reference's own segmentation always
        // is (segments split at gaps > 50 steps, extended by 3 steps) — so overlapping ranges are refused instead of silently
        // dropping the second residual.
        std::vector<std::pair<double, double>> rg;
        size_t ko2 = 0;
        for (int q = 0; q < st->n_splines; ++q) {
            rg.emplace_back(st->knots[ko2], st->knots[ko2 + (size_t) st->n_cp[(size_t) q] + 3]);
            ko2 += (size_t) st->n_cp[(size_t) q] + 4;
        }
        std::sort(rg.begin(), rg.end());
        for (size_t q = 1; q < rg.size(); ++q)
            if (!(rg[q].first > rg[q - 1].second))
                return ecb_fail(ctx, ECB_ERR_UNSUPPORTED,
                                "spline knot ranges overlap ([%g, %g] and [%g, %g]): an event inside both would need one residual "
                                "per spline", rg[q - 1].first, rg[q - 1].second, rg[q].first, rg[q].second);
        size_t ko = 0;
        double prev_end = -1e300;
        for (int q = 0; q < st->n_splines; ++q) {
            const double b = st->knots[ko], e = st->knots[ko + (size_t) st->n_cp[(size_t) q] + 3];
            if (!(b > prev_end)) fused = false;
            prev_end = e;
            ko += (size_t) st->n_cp[(size_t) q] + 4;
        }
    }
    a.cp_off = (const int *) st->d_cp_off.p;
    a.span_off = (const int *) st->d_span_off.p;
    a.basis = nullptr;
    a.cp0 = a.span = nullptr;
    if (fused) {
        if ((rc = ecb_reserve(ctx, st->basis, nn * 32))) return rc;
        if ((rc = ecb_reserve(ctx, st->cp0, nn * 4))) return rc;
        if ((rc = ecb_reserve(ctx, st->span, nn * 4))) return rc;
        a.basis = (double *) st->basis.p;
        a.cp0 = (int *) st->cp0.p;
        a.span = (int *) st->span.p;
    }
    k_assoc_write<<<nb, AS_THREADS, 0, ctx->stream>>>(a, (const uint16_t *) st->ev_tag.p, (const int64_t *) st->ev_flag.p, (double *) st->obs.p, (double *) st->lm.p,
                                                      (double *) st->tt.p, (int *) st->spl.p, (int64_t *) st->sel_event.p,
                                                      (int *) st->sel_circle.p);
    ECB_LAUNCHED(ctx);
    ECB_PROF_END(ctx, ECB_STAGE_ASSOC);
    if ((rc = ecb_check(ctx, cudaGetLastError(), "association kernels"))) return rc;
    st->n_res = total;
    st->use_circ = fused;
    if (n_residuals) *n_residuals = total;
    return prepare_records(ctx, st, fused);
}

int ecb_cost_get_association(ecb_ctx *ctx, int64_t *event_index, int32_t *circle_id, int64_t cap) {
    if (!ctx || !ctx->cost) return ECB_ERR_STATE;
    CostState *st = (CostState *) ctx->cost;
    cudaSetDevice(ctx->device);
    const int64_t n = std::min<int64_t>(cap, st->n_res);
    if (n > 0 && event_index)
        ECB_CUDA(ctx, cudaMemcpyAsync(event_index, st->sel_event.p, (size_t) n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (n > 0 && circle_id)
        ECB_CUDA(ctx, cudaMemcpyAsync(circle_id, st->sel_circle.p, (size_t) n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return ecb_check(ctx, ecb_stream_sync(ctx), "association copy");
}

// ---- launches shared by the host-parameter entry points and the device-side LM loop (ecb_lmdev.cu) ----
}  // extern "C"

static long long exchange_timeout_clocks(ecb_ctx *ctx) {
    static const double secs = getenv("ECB_EXCHANGE_TIMEOUT_S") ? atof(getenv("ECB_EXCHANGE_TIMEOUT_S")) : 30.0;
    // cudaDevAttrClockRate is a live driver query (~1 ms, much more while nvidia-smi polls the device): asked once per device,
    // not per exchange — it was 60 % of a multi-GPU LM iteration (profiles/r2s_lm_exchange_timing.md)
    static int khz_of[64] = {0};
    const int d = ctx->device & 63;
    if (khz_of[d] == 0) {
        int khz = 1965000;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
        khz_of[d] = khz > 0 ? khz : 1965000;
    }
    return (long long) (secs * 1e3 * (double) khz_of[d]);
}

static void fill_ne_args(CostState *st, NeArgs &a, const double *d_params, const int *d_go) {
    a.obs = (const double *) st->obs.p;
    a.lm = (const double *) st->lm.p;
    a.circ = st->use_circ ? (const int *) st->sel_circle.p : nullptr;
    a.lm_tab = (const double *) st->lm_tab.p;
    a.basis = (const double *) st->basis.p;
    a.cp0 = (const int *) st->cp0.p;
    a.items = (const Item *) st->items.p;
    a.n_items = st->n_items;
    a.params = d_params;
    a.total_cp = st->total_cp;
    a.radius = st->radius;
    a.huber = st->huber;
    a.part = (double *) st->part.p;
    a.go = d_go;
}

// k_normal_eq over the residual set with the parameters at d_params (device: intr 9 | rot 4C | trans 3C)
static int launch_normal_eq_kernel(ecb_ctx *ctx, CostState *st, const double *d_params, const int *d_go) {
    NeArgs a;
    fill_ne_args(st, a, d_params, d_go);
    const size_t smem = (size_t) NE_WARPS * TILE_ROWS * TILE_LD * 8;
    // 2 CTAs x 8 warps per SM at 128 registers.  More resident warps for the latency-bound residual / Jacobian phase were tried
    // with 4-warp CTAs: 5 per SM at 96 registers (more spills) 2.10 ms, 4 per SM at 128 registers 1.87 ms, against 1.81 ms
    // (profiles/r2s_ab_normal_eq.jsonl).
    static const int variant = getenv("ECB_NE_VARIANT") ? atoi(getenv("ECB_NE_VARIANT")) : 2;
    void (*kern)(const NeArgs) = st->so3 ? k_normal_eq<1, true> : (variant == 1 ? k_normal_eq<1, false> : k_normal_eq<2, false>);
    ECB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    const int grid = std::min((st->n_items + NE_WARPS - 1) / NE_WARPS, ctx->sm_count * ((variant == 1 || st->so3) ? 1 : 2));
    kern<<<grid, NE_THREADS, smem, ctx->stream>>>(a);
    ECB_LAUNCHED(ctx);
    return ECB_OK;
}

static const int *item_start_of(CostState *st) {
    return (const int *) ((const char *) st->items.p + (size_t) std::max(st->n_items, 1) * sizeof(Item));
}

// the span range this rank's residuals contribute to
static void span_range(CostState *st, int &lo, int &hi) {
    lo = 0;
    hi = -1;
    for (int s = 0; s < st->total_spans; ++s)
        if (st->h_span_start[(size_t) s + 1] > st->h_span_start[(size_t) s]) {
            if (hi < 0) lo = s;
            hi = s;
        }
}

// cost-only evaluation at d_params -> *d_cost (device scalar); async
int ecb_cost_dev_eval(ecb_ctx *ctx, const double *d_params, const int *d_go, double *d_cost) {
    CostState *st = (CostState *) ctx->cost;
    int rc;
    if ((rc = ecb_reserve(ctx, st->cost_part, 4096 * 8))) return rc;
    if (st->n_res == 0) {
        ECB_CUDA(ctx, cudaMemsetAsync(d_cost, 0, 8, ctx->stream));
        return ECB_OK;
    }
    const int grid = (int) std::min<int64_t>((st->n_res + 255) / 256, (int64_t) ctx->sm_count * 8);
    ECB_PROF_BEGIN(ctx, ECB_STAGE_COST);
    static const int cost_minb = getenv("ECB_COST_MINB") ? atoi(getenv("ECB_COST_MINB")) : 4;  // 64 registers, 32 warps / SM: 0.35 -> 0.33 ms
    (st->so3 ? k_cost<true, 2> : (cost_minb >= 4 ? k_cost<false, 4> : k_cost<false, 3>))<<<grid, 256, 0, ctx->stream>>>(
        (const double *) st->obs.p, (const double *) st->lm.p, st->use_circ ? (const int *) st->sel_circle.p : nullptr,
        (const double *) st->lm_tab.p, (const double *) st->basis.p, (const int *) st->cp0.p, st->n_res, d_params, st->total_cp,
        st->radius, st->huber, (double *) st->cost_part.p, d_go);
    ECB_LAUNCHED(ctx);
    k_reduce_cost<<<1, 256, 0, ctx->stream>>>((const double *) st->cost_part.p, grid, 1, 0, d_cost, d_go);
    ECB_LAUNCHED(ctx);
    ECB_PROF_END(ctx, ECB_STAGE_COST);
    return ecb_check(ctx, cudaGetLastError(), "cost kernels");
}

// normal equations at d_params -> d_out (packed, this rank's residuals only); async
int ecb_cost_dev_normal_eq(ecb_ctx *ctx, const double *d_params, const int *d_go, double *d_out) {
    CostState *st = (CostState *) ctx->cost;
    if (st->n_items <= 0) return ECB_OK;
    int rc;
    ECB_PROF_BEGIN(ctx, ECB_STAGE_NORMAL_EQ);
    if ((rc = launch_normal_eq_kernel(ctx, st, d_params, d_go))) return rc;
    k_reduce_spans<<<st->total_spans, 192, 0, ctx->stream>>>((const double *) st->part.p, item_start_of(st), st->total_spans, d_out, d_go);
    ECB_LAUNCHED(ctx);
    k_reduce_cost<<<1, 256, 0, ctx->stream>>>((const double *) st->part.p, st->n_items, PART_STRIDE, PART_SC + 3,
                                              d_out + (size_t) st->total_spans * OUT_STRIDE, d_go);
    ECB_LAUNCHED(ctx);
    ECB_PROF_END(ctx, ECB_STAGE_NORMAL_EQ);
    return ecb_check(ctx, cudaGetLastError(), "normal equation kernels");
}

// the same with the inter-GPU sum fused in (all ranks' residuals -> d_out on every rank); the epoch is a device counter
// (d_epoch, already advanced for this exchange) or a host value.  phases as in ecb_cost_normal_eq_exchange.  d_err: sticky
// timeout bits (device word).
int ecb_cost_dev_normal_eq_exchange(ecb_ctx *ctx, const double *d_params, const int *d_go, int rank, int n_ranks,
                                    void *const *recv_buffers, const unsigned long long *d_epoch, unsigned long long epoch,
                                    int phases, double *d_out, unsigned *d_err) {
    CostState *st = (CostState *) ctx->cost;
    int rc;
    if ((rc = ecb_reserve(ctx, st->cost_part, 4096 * 8))) return rc;
    XPeers peers;
    memset(&peers, 0, sizeof peers);
    peers.n_ranks = n_ranks;
    peers.rank = rank;
    for (int p = 0; p < n_ranks; ++p) {
        if (!recv_buffers[p]) return ECB_ERR_ARG;
        peers.base[p] = (double *) recv_buffers[p];
    }
    XCtl x;
    x.go = d_go;
    x.d_epoch = d_epoch;
    x.epoch = epoch;
    x.slot = xslot_doubles(st->total_spans);
    x.n_ranks = n_ranks;
    x.rank = rank;
    int lo, hi;
    span_range(st, lo, hi);
    double *d_cost = (double *) st->cost_part.p + 4001;
    if (phases & ECB_EXCHANGE_SEND) {
        ECB_CUDA(ctx, cudaMemsetAsync(d_cost, 0, 8, ctx->stream));  // (harmless when the launch is gated off)
        if (st->n_items > 0 && hi >= lo) {
            ECB_PROF_BEGIN(ctx, ECB_STAGE_NORMAL_EQ);
            if ((rc = launch_normal_eq_kernel(ctx, st, d_params, d_go))) return rc;
            k_reduce_spans_push<<<hi - lo + 1, 192, 0, ctx->stream>>>((const double *) st->part.p, item_start_of(st), lo, hi, peers, x);
            ECB_LAUNCHED(ctx);
            k_reduce_cost<<<1, 256, 0, ctx->stream>>>((const double *) st->part.p, st->n_items, PART_STRIDE, PART_SC + 3, d_cost, d_go);
            ECB_LAUNCHED(ctx);
            ECB_PROF_END(ctx, ECB_STAGE_NORMAL_EQ);
        }
        k_exchange_signal<<<1, 32, 0, ctx->stream>>>(peers, x, lo, hi, d_cost);
        ECB_LAUNCHED(ctx);
    }
    if (phases & ECB_EXCHANGE_RECV) {
        k_exchange_wait<<<1, 32, 0, ctx->stream>>>(peers.base[rank], x, exchange_timeout_clocks(ctx), d_err);
        ECB_LAUNCHED(ctx);
        k_exchange_sum<<<st->total_spans + 1, 192, 0, ctx->stream>>>(peers.base[rank], x, st->total_spans, d_out, d_err);
        ECB_LAUNCHED(ctx);
    }
    return ecb_check(ctx, cudaGetLastError(), "exchange kernels");
}

// scalar sum over the ranks (device-side LM loop); *d_value: this rank's part in, the sum out
int ecb_cost_dev_scalar_exchange(ecb_ctx *ctx, const int *d_go, int rank, int n_ranks, void *const *recv_buffers,
                                 const unsigned long long *d_epoch, double *d_value, unsigned *d_err) {
    CostState *st = (CostState *) ctx->cost;
    XPeers peers;
    memset(&peers, 0, sizeof peers);
    peers.n_ranks = n_ranks;
    peers.rank = rank;
    for (int p = 0; p < n_ranks; ++p) peers.base[p] = (double *) recv_buffers[p];
    XCtl x;
    x.go = d_go;
    x.d_epoch = d_epoch;
    x.epoch = 0;
    x.slot = xslot_doubles(st->total_spans);
    x.n_ranks = n_ranks;
    x.rank = rank;
    k_exchange_scalar<<<1, 32, 0, ctx->stream>>>(peers, x, 2 * (size_t) n_ranks * x.slot, d_value, exchange_timeout_clocks(ctx), d_err);
    ECB_LAUNCHED(ctx);
    return ecb_check(ctx, cudaGetLastError(), "scalar exchange");
}

double *ecb_cost_params_buffer(ecb_ctx *ctx, size_t *doubles) {  // the context's parameter upload buffer (intr | rot | trans)
    CostState *st = (CostState *) ctx->cost;
    const size_t n = 9 + 7 * (size_t) st->total_cp;
    if (ecb_reserve(ctx, st->params, n * 8)) return nullptr;
    if (doubles) *doubles = n;
    return (double *) st->params.p;
}

extern "C" {

int ecb_cost_eval(ecb_ctx *ctx, const double *intrinsics, const double *rot_cp, const double *trans_cp, double *cost) {
    if (!ctx || !ctx->cost || !intrinsics || !rot_cp || !trans_cp || !cost) return ECB_ERR_ARG;
    cudaSetDevice(ctx->device);
    CostState *st = (CostState *) ctx->cost;
    int rc;
    if ((rc = upload_params(ctx, st, intrinsics, rot_cp, trans_cp))) return rc;
    if ((rc = ecb_reserve(ctx, st->cost_part, 4096 * 8))) return rc;
    *cost = 0.0;
    if (st->n_res == 0) return ECB_OK;
    if ((rc = ecb_cost_dev_eval(ctx, (const double *) st->params.p, nullptr, (double *) st->cost_part.p + 4000))) return rc;
    return ecb_d2h(ctx, cost, (double *) st->cost_part.p + 4000, 8);
}

// Packed result (device or host): per span s  [H 33x33 full symmetric | g 33], then [cost, 0].
// d_out (device pointer, may be NULL -> internal buffer); h_out (host, may be NULL).
int ecb_cost_normal_eq(ecb_ctx *ctx, const double *intrinsics, const double *rot_cp, const double *trans_cp, void *d_out,
                       double *h_out, double *cost) {
    if (!ctx || !ctx->cost || !intrinsics || !rot_cp || !trans_cp) return ECB_ERR_ARG;
    cudaSetDevice(ctx->device);
    CostState *st = (CostState *) ctx->cost;
    int rc;
    if ((rc = upload_params(ctx, st, intrinsics, rot_cp, trans_cp))) return rc;
    const size_t n_out = (size_t) st->total_spans * OUT_STRIDE + 2;
    if ((rc = ecb_reserve(ctx, st->out, (n_out + 8) * 8))) return rc;
    double *out = d_out ? (double *) d_out : (double *) st->out.p;
    ECB_CUDA(ctx, cudaMemsetAsync(out, 0, n_out * 8, ctx->stream));
    if ((rc = ecb_cost_dev_normal_eq(ctx, (const double *) st->params.p, nullptr, out))) return rc;
    if (h_out) {
        if ((rc = ecb_d2h(ctx, h_out, out, n_out * 8))) return rc;
        if (cost) *cost = h_out[(size_t) st->total_spans * OUT_STRIDE];
    } else if (cost) {
        return ecb_d2h(ctx, cost, out + (size_t) st->total_spans * OUT_STRIDE, 8);
    }
    return ECB_OK;
}

// ---- fused reduce + exchange (multi-GPU) ----
size_t ecb_exchange_buffer_bytes(ecb_ctx *ctx, int n_ranks) {
    if (!ctx || !ctx->cost || n_ranks < 1) return 0;
    CostState *st = (CostState *) ctx->cost;
    // two parity blocks of n_ranks span slots, then two parity blocks of n_ranks (epoch, value) scalar slots
    return (2 * (size_t) n_ranks * xslot_doubles(st->total_spans) + 4 * (size_t) n_ranks + 8) * 8;
}

int ecb_cost_normal_eq_exchange(ecb_ctx *ctx, const double *intrinsics, const double *rot_cp, const double *trans_cp, int rank,
                                int n_ranks, void *const *recv_buffers, uint64_t epoch, int phases, void *d_out, double *cost) {
    if (!ctx || !ctx->cost || !intrinsics || !rot_cp || !trans_cp || !recv_buffers || !d_out) return ECB_ERR_ARG;
    if (!(phases & 3)) return ECB_ERR_ARG;
    if (n_ranks < 1 || n_ranks > ECB_MAX_PEERS || rank < 0 || rank >= n_ranks || epoch == 0) return ECB_ERR_ARG;
    cudaSetDevice(ctx->device);
    CostState *st = (CostState *) ctx->cost;
    int rc;
    if ((rc = upload_params(ctx, st, intrinsics, rot_cp, trans_cp))) return rc;
    if ((rc = ecb_reserve(ctx, st->cost_part, 4096 * 8))) return rc;
    // Epochs on the wire never repeat for a receive buffer: a caller that restarts its sequence (a second LM run on the same
    // buffers begins at 1 again) starts a new generation, so a header left over from the previous run cannot satisfy the wait.
    // Every rank sees the same epoch sequence, hence the same generations.  (SEND and RECV of one exchange carry one epoch.)
    if ((phases & ECB_EXCHANGE_SEND) && epoch <= st->xch_last_epoch) ++st->xch_generation;
    if (phases & ECB_EXCHANGE_SEND) st->xch_last_epoch = epoch;
    const unsigned long long wire = ((unsigned long long) st->xch_generation << 40) | (unsigned long long) (epoch & 0xFFFFFFFFFFull);
    // the parity must follow the caller's epoch (double buffering), generations keep it: shift left by one, parity in bit 0
    const unsigned long long wire_epoch = (wire << 1) | (epoch & 1ull);
    unsigned *d_err = (unsigned *) ((double *) st->cost_part.p + 4002);
    if (phases & ECB_EXCHANGE_SEND) ECB_CUDA(ctx, cudaMemsetAsync(d_err, 0, 4, ctx->stream));
    if ((rc = ecb_cost_dev_normal_eq_exchange(ctx, (const double *) st->params.p, nullptr, rank, n_ranks, recv_buffers, nullptr,
                                              wire_epoch, phases, (double *) d_out, d_err)))
        return rc;
    if (!(phases & ECB_EXCHANGE_RECV)) return ECB_OK;
    if (cost) {
        double hc[2];
        if ((rc = ecb_d2h(ctx, hc, (double *) d_out + (size_t) st->total_spans * OUT_STRIDE, 8))) return rc;
        unsigned herr = 0;
        if ((rc = ecb_d2h(ctx, &herr, d_err, 4))) return rc;
        if (herr) return ecb_fail(ctx, ECB_ERR_STATE, "normal-equation exchange timed out waiting for ranks (mask 0x%x)", herr);
        *cost = hc[0];
    }
    // cost == NULL: nothing is read back here; a timeout leaves NaN in the cost slot of d_out (k_exchange_sum), which every
    // consumer of the packed buffer sees
    return ECB_OK;
}

}  // extern "C"
