// k_pair — candidate circles of one window from the kept clusters of both polarities
// (CirclesEventFrame::extractFeatures, CirclesEventFrame.cpp:137-312) and the batched Kasa circle fit
// (CirclesEventFrame::fitCircle, CirclesEventFrame.cpp:361-415).  One CTA per window, one warp per
// positive cluster ("warp-per-cluster batched fit"); k-NN over <= max_k medians by warp arg-min
// (replaces the nanoflann trees, :160-168; ties -> lowest index).
//
// This translation unit is compiled with -fmad=false: the reference is built without FMA contraction
// (CMakeLists.txt:6-19, x86-64 baseline), and the 3x3 LU / fit-error arithmetic below follows it operation
// by operation.  The 9 moment sums are exact integers (int64 in k_cluster), so A and b are bit-identical.
#include <float.h>

#include "ecb_window.cuh"

namespace {

constexpr int PAIR_THREADS = 256;
constexpr int KNN_MAX = 8;

// Eigen PartialPivLU<Matrix3d>::solve, unblocked, first-max pivot [external: Eigen, not in /root/reference]
__device__ void lu_solve3(double A[3][3], const double b_in[3], double x[3]) {
    int perm[3] = {0, 1, 2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int piv = k;
        double big = fabs(A[k][k]);
        for (int i = k + 1; i < 3; ++i)
            if (fabs(A[i][k]) > big) {
                big = fabs(A[i][k]);
                piv = i;
            }
        if (big != 0.0) {
            if (piv != k) {
                for (int j = 0; j < 3; ++j) {
                    double t = A[k][j];
                    A[k][j] = A[piv][j];
                    A[piv][j] = t;
                }
                int t = perm[k];
                perm[k] = perm[piv];
                perm[piv] = t;
            }
            for (int i = k + 1; i < 3; ++i) A[i][k] /= A[k][k];
        }
        for (int i = k + 1; i < 3; ++i)
            for (int j = k + 1; j < 3; ++j) A[i][j] -= A[i][k] * A[k][j];
    }
    double y[3] = {b_in[perm[0]], b_in[perm[1]], b_in[perm[2]]};
    y[1] -= A[1][0] * y[0];
    y[2] -= (A[2][0] * y[0] + A[2][1] * y[1]);
    y[2] /= A[2][2];
    y[1] -= A[1][2] * y[2];
    y[1] /= A[1][1];
    y[0] -= (A[0][1] * y[1] + A[0][2] * y[2]);
    y[0] /= A[0][0];
    x[0] = y[0];
    x[1] = y[1];
    x[2] = y[2];
}

// m = Sx Sy Sxx Syy Sxy Sxxx Syyy Sxyy Sxxy, N = sample count   (CirclesEventFrame.cpp:406-414)
__device__ void fit_from_moments(const double *m, double N, double &cx, double &cy, double &r) {
    double A[3][3] = {{2 * m[0], 2 * m[1], N}, {2 * m[2], 2 * m[4], m[0]}, {2 * m[4], 2 * m[3], m[1]}};
    double b[3] = {m[2] + m[3], m[5] + m[7], m[8] + m[6]};
    double x[3];
    lu_solve3(A, b, x);
    cx = x[0];
    cy = x[1];
    r = sqrt(x[0] * x[0] + x[1] * x[1] + x[2]);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the members of |‖p-c‖ - r|   (CirclesEventFrame.cpp:209-214,300-305)
// DIRECT: `mem` already holds the members' packed pixels (staged in shared memory); else it holds pids into `pts`
template <bool DIRECT>
__device__ double warp_abs_dev(const uint32_t *pts, const uint32_t *mem, int sz, double cx, double cy, double r) {
    double s = 0;
    for (int i = threadIdx.x & 31; i < sz; i += 32) {
        const uint32_t p = DIRECT ? mem[i] : pts[mem[i]];
        const double dx = (double) ECB_PIX_X(p) - cx, dy = (double) ECB_PIX_Y(p) - cy;
        s += fabs(sqrt(dx * dx + dy * dy) - r);
    }
    return warp_sum(s);
}

// k nearest medians (ascending squared distance, ties by index), warp cooperative
__device__ void warp_knn(const int *mx, const int *my, int n, int qx, int qy, int k, int *idx, unsigned long long *d2) {
    unsigned long long last = 0;
    bool first = true;
    for (int j = 0; j < k; ++j) {
        unsigned long long best = ~0ull;
        for (int i = threadIdx.x & 31; i < n; i += 32) {
            const long long dx = qx - mx[i], dy = qy - my[i];
            const unsigned long long key = ((unsigned long long) (dx * dx + dy * dy) << 32) | (unsigned) i;
            if ((first || key > last) && key < best) best = key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
            best = t < best ? t : best;
        }
        idx[j] = best == ~0ull ? -1 : (int) (best & 0xFFFFFFFFu);
        d2[j] = best >> 32;
        last = best;
        first = false;
    }
}

template <bool DIRECT>
__global__ void __launch_bounds__(PAIR_THREADS, 4) k_pair(const PairArgs a) {
    extern __shared__ __align__(16) uint32_t sm_pix[];  // DIRECT: member pixels of both polarities
    __shared__ int mx[2][ECB_MAXK_LIMIT], my[2][ECB_MAXK_LIMIT];
    __shared__ int acc_ni[ECB_MAXK_LIMIT];
    __shared__ double acc_c[ECB_MAXK_LIMIT][3];
    __shared__ uint32_t ws[33];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarp = PAIR_THREADS >> 5;

    for (int w = blockIdx.x; w < a.n_win; w += gridDim.x) {
        const ProbDesc dn = a.prob[2 * w], dp = a.prob[2 * w + 1];
        const ProbHdr hn = a.hdr[2 * w], hp = a.hdr[2 * w + 1];
        const KeptCluster *kn = a.ktab + (size_t) (2 * w) * a.max_k, *kp = a.ktab + (size_t) (2 * w + 1) * a.max_k;
        const int nkn = hn.n_kept, nkp = hp.n_kept;
        __syncthreads();
        for (int i = tid; i < nkn; i += PAIR_THREADS) {
            mx[0][i] = kn[i].med_x;
            my[0][i] = kn[i].med_y;
        }
        for (int i = tid; i < nkp; i += PAIR_THREADS) {
            mx[1][i] = kp[i].med_x;
            my[1][i] = kp[i].med_y;
            acc_ni[i] = -1;
        }
        __syncthreads();
        const bool enough0 = dn.n > 0 && dp.n > 0 && (uint32_t) nkp >= a.rows_cols && (uint32_t) nkn >= a.rows_cols;
        const bool enough = enough0;
        const uint32_t *ptsP = a.pts[1] + dp.off, *ptsN = a.pts[0] + dn.off;
        const uint32_t *memP = a.kmem[1] + dp.off, *memN = a.kmem[0] + dn.off;
        if (DIRECT && enough0) {
            // stage the kept clusters' member pixels once (coalesced list read, gathered pixel read), then every
            // fit-error loop runs out of shared memory
            const int totP = nkp ? kp[nkp - 1].mem_off + kp[nkp - 1].size : 0, totN = nkn ? kn[nkn - 1].mem_off + kn[nkn - 1].size : 0;
            for (int i = tid; i < totN; i += PAIR_THREADS) sm_pix[i] = ptsN[memN[i]];
            for (int i = tid; i < totP; i += PAIR_THREADS) sm_pix[a.smem_cap + i] = ptsP[memP[i]];
            memN = sm_pix;
            memP = sm_pix + a.smem_cap;
            __syncthreads();
        }
        const double gate = 4 * a.rthr * a.rthr;
        if (enough) {
            for (int pi = wid; pi < nkp; pi += nwarp) {
                int nidx[KNN_MAX], pidx[KNN_MAX];
                unsigned long long d2[KNN_MAX];
                if (!a.fit_circle) {  // CirclesEventFrame.cpp:282-312
                    warp_knn(mx[0], my[0], nkn, mx[1][pi], my[1][pi], 1, nidx, d2);
                    if ((double) d2[0] > gate) continue;
                    const int n0 = nidx[0];
                    warp_knn(mx[1], my[1], nkp, mx[0][n0], my[0][n0], 1, pidx, d2);
                    if (pidx[0] != pi) continue;
                    const double px = mx[1][pi], py = my[1][pi], qx = mx[0][n0], qy = my[0][n0];
                    const double cx = (px + qx) / 2, cy = (py + qy) / 2;
                    const double ddx = px - qx, ddy = py - qy;
                    const double r = sqrt(ddx * ddx + ddy * ddy) / 2;
                    double e = warp_abs_dev<DIRECT>(ptsP, memP + kp[pi].mem_off, kp[pi].size, cx, cy, r);
                    // the reference accumulates + members then - members into one sum
                    e += warp_abs_dev<DIRECT>(ptsN, memN + kn[n0].mem_off, kn[n0].size, cx, cy, r);
                    e /= (double) (kp[pi].size + kn[n0].size) * r;
                    if (e < 10 / r && lane == 0) {
                        acc_ni[pi] = n0;
                        acc_c[pi][0] = cx;
                        acc_c[pi][1] = cy;
                        acc_c[pi][2] = r;
                    }
                } else {  // CirclesEventFrame.cpp:180-281
                    const int K = min(min(a.knn_num, KNN_MAX), min(nkn, nkp));
                    double ferr[KNN_MAX], fr[KNN_MAX], fcx[KNN_MAX], fcy[KNN_MAX];
                    warp_knn(mx[0], my[0], nkn, mx[1][pi], my[1][pi], K, nidx, d2);
                    int real = K;
                    for (int oi = 0; oi < K; ++oi)
                        if ((double) d2[oi] > (double) d2[0] * 4 || (double) d2[oi] > gate) {
                            real = oi;
                            break;
                        }
                    if (real == 0) continue;
                    auto fit_pair = [&](int p, int n, double &cx, double &cy, double &r) -> double {
                        double m[9];
#pragma unroll
                        for (int q = 0; q < 9; ++q) m[q] = kp[p].m[q] + kn[n].m[q];
                        fit_from_moments(m, (double) (kp[p].size + kn[n].size), cx, cy, r);
                        const double ddx = (double) mx[1][p] - (double) mx[0][n], ddy = (double) my[1][p] - (double) my[0][n];
                        const double approx = sqrt(ddx * ddx + ddy * ddy) / 2;
                        if (r > a.rthr || r > 2 * approx) return DBL_MAX;
                        double e = warp_abs_dev<DIRECT>(ptsP, memP + kp[p].mem_off, kp[p].size, cx, cy, r);
                        e += warp_abs_dev<DIRECT>(ptsN, memN + kn[n].mem_off, kn[n].size, cx, cy, r);
                        return e / ((double) (kp[p].size + kn[n].size) * r);
                    };
                    for (int j = 0; j < real; ++j) ferr[j] = fit_pair(pi, nidx[j], fcx[j], fcy[j], fr[j]);
                    int nmin = 0;
                    for (int j = 1; j < real; ++j)
                        if (ferr[j] < ferr[nmin]) nmin = j;
                    if (!(ferr[nmin] < 2 / fr[nmin])) continue;
                    const int nsel = nidx[nmin];
                    warp_knn(mx[1], my[1], nkp, mx[0][nsel], my[0][nsel], K, pidx, d2);
                    real = K;
                    for (int oi = 0; oi < K; ++oi)
                        if ((double) d2[oi] > (double) d2[0] * 4 || (double) d2[oi] > gate) {
                            real = oi;
                            break;
                        }
                    if (real == 0) continue;
                    for (int i = 0; i < real; ++i) ferr[i] = fit_pair(pidx[i], nsel, fcx[i], fcy[i], fr[i]);
                    int pmin = 0;
                    for (int i = 1; i < real; ++i)
                        if (ferr[i] < ferr[pmin]) pmin = i;
                    if (pidx[pmin] == pi && lane == 0) {
                        acc_ni[pi] = nsel;
                        acc_c[pi][0] = fcx[pmin];
                        acc_c[pi][1] = fcy[pmin];
                        acc_c[pi][2] = fr[pmin];
                    }
                }
            }
        }
        __syncthreads();
        // ordered compaction (candidates are pushed in ascending pi, CirclesEventFrame.cpp:274-278,306-310)
        uint32_t run = 0;
        double *out = a.cand + (size_t) w * a.cand_stride * 5;
        for (int c0 = 0; c0 < nkp; c0 += PAIR_THREADS) {
            const int pi = c0 + tid;
            const bool ok = enough && pi < nkp && acc_ni[pi] >= 0;
            uint32_t tot;
            const uint32_t ex = block_excl_scan(ok ? 1u : 0u, ws, &tot);
            if (ok && run + ex < (uint32_t) a.cand_stride) {
                double *o = out + (size_t) (run + ex) * 5;
                o[0] = pi;
                o[1] = acc_ni[pi];
                o[2] = acc_c[pi][0];
                o[3] = acc_c[pi][1];
                o[4] = acc_c[pi][2];
            }
            run += tot;
        }
        if (tid == 0) {
            ecb_window_summary s;
            s.ev_lo = a.lohi[2 * w];
            s.ev_hi = a.lohi[2 * w + 1] > a.lohi[2 * w] ? a.lohi[2 * w + 1] : a.lohi[2 * w];
            s.n_points[0] = dn.n;
            s.n_points[1] = dp.n;
            s.n_clusters[0] = hn.n_clusters;
            s.n_clusters[1] = hp.n_clusters;
            s.n_kept[0] = nkn;
            s.n_kept[1] = nkp;
            s.n_candidates = (int32_t) run;
            s.status = hn.status | hp.status;
            s.point_offset[0] = dn.off;
            s.point_offset[1] = dp.off;
            a.summary[w] = s;
        }
    }
}

// batched fit of explicit point sets: one warp per set (ecb_fit_circles)
__global__ void k_fit(const double *__restrict__ xy, const int64_t *__restrict__ off, int n_sets, double *__restrict__ out) {
    const int set = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (set >= n_sets) return;
    const int64_t b = off[set], e = off[set + 1];
    double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = b + lane; i < e; i += 32) {
        const double x = xy[2 * i], y = xy[2 * i + 1];
        const double xx = x * x, yy = y * y, xyv = x * y;
        m[0] += x;
        m[1] += y;
        m[2] += xx;
        m[3] += yy;
        m[4] += xyv;
        m[5] += xx * x;
        m[6] += yy * y;
        m[7] += xyv * y;
        m[8] += x * xyv;
    }
#pragma unroll
    for (int q = 0; q < 9; ++q) m[q] = warp_sum(m[q]);
    if (lane == 0) {
        double cx, cy, r;
        fit_from_moments(m, (double) (e - b), cx, cy, r);
        out[3 * set] = cx;
        out[3 * set + 1] = cy;
        out[3 * set + 2] = r;
    }
}

}  // namespace

int ecb_launch_pair(ecb_ctx *ctx, PairArgs &a) {
    if (a.n_win <= 0) return ECB_OK;
    int grid = a.n_win < ctx->sm_count * 8 ? a.n_win : ctx->sm_count * 8;
    // stage member pixels in shared memory when both polarities of the largest window fit
    const size_t smem = (size_t) 2 * a.smem_cap * 4;
    const bool direct = a.smem_cap > 0 && smem <= 96 * 1024;
    ECB_PROF_BEGIN(ctx, ECB_STAGE_PAIR);
    if (direct) {
        ECB_CUDA(ctx, cudaFuncSetAttribute(k_pair<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)  /* constant: race-free */);
        k_pair<true><<<grid, PAIR_THREADS, smem, ctx->stream>>>(a);
    } else {
        k_pair<false><<<grid, PAIR_THREADS, 0, ctx->stream>>>(a);
    }
    ECB_PROF_END(ctx, ECB_STAGE_PAIR);
    ECB_LAUNCHED(ctx);
    return ecb_check(ctx, cudaGetLastError(), "k_pair launch");
}

int ecb_launch_fit(ecb_ctx *ctx, const double *d_xy, const int64_t *d_off, int n_sets, double *d_out) {
    if (n_sets <= 0) return ECB_OK;
    const int thr = 128;
    k_fit<<<(n_sets * 32 + thr - 1) / thr, thr, 0, ctx->stream>>>(d_xy, d_off, n_sets, d_out);
    ECB_LAUNCHED(ctx);
    return ecb_check(ctx, cudaGetLastError(), "k_fit launch");
}
