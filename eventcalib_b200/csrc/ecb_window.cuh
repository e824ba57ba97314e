// Launch interfaces of the ingest / window / pair kernels.
#pragma once
#include "ecb_cluster.cuh"

struct WindowArgs {
    const uint32_t *xyp;      // packed events
    const int64_t *lohi;      // [n_win][2] event index range
    const int64_t *ptoff;     // [n_win] offset of the window's point slots (capacity hi-lo per polarity)
    uint32_t *arrive[2];      // distinct pixels per polarity in first-arrival order (before cancellation)
    uint32_t *pts[2];         // surviving pixels in pid order
    ProbDesc *prob;           // [2*n_win] clustering problems (index 2*w+pol)
    uint32_t *max_n;          // device words: [0] max points of any problem (sizes the cluster kernel's arrays),
                              //               [1] max distinct pixels before cancellation (sizes k_uset_order)
    int n_win, W, H, RW;
    uint32_t *gplanes;        // per-CTA bit planes in L2 scratch when they exceed shared memory (else NULL)
    int order_mode;           // 1: mark cancelled pixels in `arrive` (bit 31) for k_uset_order instead of compacting
};

// k_uset_order (ecb_order.cu): libstdc++ unordered_set iteration order of every (window, polarity) pixel set
struct OrderArgs {
    const ProbDesc *prob;     // pad0 = arrival count
    int n_prob;
    unsigned *work_counter;
    const uint32_t *arrive[2];
    uint32_t *pts[2];
    uint32_t *gscratch;
    size_t gscratch_stride;
    int arrays_in_smem, m_cap, b_cap;
    const unsigned long long *htab;  // [max(W, H)] hash_double((double) v) + 0x9e3779b9 (NULL: hash computed per pixel)
};

struct PairArgs {
    const ProbDesc *prob;
    const ProbHdr *hdr;
    const KeptCluster *ktab;
    const uint32_t *pts[2];
    const uint32_t *kmem[2];
    const int64_t *lohi;
    ecb_window_summary *summary;
    double *cand;             // [n_win][cand_stride][5]
    int n_win, max_k, cand_stride;
    int smem_cap;             // max points of any problem (capacity of the shared-memory member staging), 0 = off
    int fit_circle, knn_num;
    uint32_t rows_cols;
    double rthr;
    unsigned *win_counter;    // non-null: windows are handed out dynamically (zeroed by the caller), grid = resident CTAs
    double *gtab;             // per-CTA kept-cluster tables in global scratch when max_k is too large for shared memory
    size_t gtab_stride;       // doubles per CTA
};

int ecb_launch_ingest(ecb_ctx *ctx, const void *d_raw, int64_t n);
int ecb_launch_bounds(ecb_ctx *ctx, const double *d_win, int n_win, int64_t *d_lohi);
int ecb_launch_window(ecb_ctx *ctx, WindowArgs &a);
int ecb_launch_order(ecb_ctx *ctx, OrderArgs &a, int max_m);
int ecb_launch_pair(ecb_ctx *ctx, PairArgs &a);
int ecb_launch_rectify(ecb_ctx *ctx, const int32_t *d_win, int n_frames, int n_feat, const double *d_img, double *d_out,
                       const PairArgs &pa, const int32_t *const labels[2], double thr);
int ecb_launch_fit(ecb_ctx *ctx, const double *d_xy, const int64_t *d_off, int n_sets, double *d_out);
