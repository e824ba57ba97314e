// Shared device helpers and the context layout of libecb (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/eventcalib_b200.h"

#define ECB_MAX_EPS 15      // stencil rows fit a 64-bit funnel window (2*15+1 = 31 bits)
#define ECB_NONE 0xFFFFFFFFu
#define ECB_GH_MAXD 4       // largest point dimension of the general (grid-hash) DBSCAN path
#define ECB_GH_NBUF 32      // device buffers of that path

// packed event / pixel word:  x bits 0..14, y bits 15..29, bit 30 invalid, bit 31 polarity
#define ECB_PIX_X(p) ((p) & 0x7FFFu)
#define ECB_PIX_Y(p) (((p) >> 15) & 0x7FFFu)
#define ECB_PIX_POL(p) ((p) >> 31)
#define ECB_PIX_XY(p) ((p) & 0x3FFFFFFFu)
#define ECB_PIX_INVALID 0x40000000u

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct ecb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    uint64_t launches = 0;
    int sm_count = 148;
    int smem_optin = 0;
    int width = 0, height = 0;
    bool prof = false;
    // How the host waits for the stream (ECB_BLOCKING_SYNC, read at creation): 0 cudaStreamSynchronize (spins), 1 sleep on a
    // blocking event, 2 poll an event with sched_yield() in between — for many contexts / ranks on few host cores (8 ranks x 4
    // slice threads on a 32-core box starve each other when every wait spins)
    int sync_mode = 0;
    cudaEvent_t sync_ev = nullptr;
    cudaEvent_t pev[ECB_N_STAGES][2] = {};
    bool pev_used[ECB_N_STAGES] = {};

    // events (SoA)
    int64_t n_events = 0;
    DevBuf ev_raw, ev_t, ev_xyp, ev_flag;
    // front end
    int n_win = 0;
    ecb_frontend_params fp{};
    DevBuf win_t, win_lohi, win_ptoff, summary, arrive, pts[2], labels[2], scratch, ktab, kmem, cand, status, pair_tab;
    int max_k_auto = 128;     // kept-cluster table capacity of the automatic mode (max_clusters = 0): grows with the data, sticky
    std::vector<int64_t> h_lohi, h_ptoff;
    int64_t total_points = 0;
    int cand_stride = 0;
    // exact member order / medians (ecb_bfs.cu): exported kd-tree, claim keys, work items, frontier scratch
    DevBuf kd_tree, bfs_key, bfs_items, bfs_front, bfs_tab;
    // k_uset_order: std::hash<double> of every integer coordinate of the sensor (+ the hash_combine constant)
    DevBuf ord_htab;
    int ord_htab_n = 0;
    // dbscan boundary
    DevBuf db_pix, db_off, db_labels, db_hdr, db_scratch, db_dims, db_hdr_b, db_ktab, db_counter;
    DevBuf gh[ECB_GH_NBUF];   // general path (ecb_gridhash.cu)
    // fit
    DevBuf fit_in, fit_off, fit_out;
    // pinned staging (device -> host), and mapped pinned staging the device pulls small uploads from
    void *pinned = nullptr;
    size_t pinned_cap = 0;
    void *up = nullptr;
    size_t up_cap = 0, up_off = 0;
    // cost evaluation (ecb_cost.cu)
    void *cost = nullptr;
};

int ecb_fail(ecb_ctx *ctx, int code, const char *fmt, ...);
int ecb_reserve(ecb_ctx *ctx, DevBuf &b, size_t bytes);
int ecb_check(ecb_ctx *ctx, cudaError_t e, const char *what);
cudaError_t ecb_stream_sync(ecb_ctx *ctx);  // wait for the context's stream (spinning or sleeping, see blocking_sync)
int ecb_d2h(ecb_ctx *ctx, void *dst, const void *src, size_t bytes);  // pinned-staged copy + stream sync
// Small host -> device upload that does not touch the copy engines: the data is staged in mapped pinned memory and a
// kernel on the context stream pulls it over PCIe, so it cannot queue behind another context's bulk record upload.
int ecb_h2d(ecb_ctx *ctx, void *dst, const void *src, size_t bytes);
#define ECB_CUDA(ctx, call)                                   \
    do {                                                      \
        int _rc = ecb_check((ctx), (call), #call);            \
        if (_rc) return _rc;                                  \
    } while (0)
#define ECB_LAUNCHED(ctx) ((ctx)->launches++)
// stage timing: ECB_PROF_BEGIN/END bracket a kernel launch with events when profiling is on
#define ECB_PROF_BEGIN(ctx, st) do { if ((ctx)->prof) cudaEventRecord((ctx)->pev[st][0], (ctx)->stream); } while (0)
#define ECB_PROF_END(ctx, st) do { if ((ctx)->prof) { cudaEventRecord((ctx)->pev[st][1], (ctx)->stream); (ctx)->pev_used[st] = true; } } while (0)

// ---------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (unsigned) o) v += t;
    }
    return v;
}

// Block-wide exclusive scan of one value per thread; returns the exclusive prefix, *total = block sum.
// `ws` is a shared array of at least 33 words. All threads of the block must call it.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *ws, uint32_t *total) {
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    uint32_t inc = warp_incl_scan(v);
    __syncthreads();  // protect ws from a previous use
    if (lane == 31) ws[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < nw ? ws[lane] : 0;
        uint32_t wi = warp_incl_scan(w);
        ws[lane] = wi - w;
        if (lane == 31) ws[32] = wi;
    }
    __syncthreads();
    *total = ws[32];
    return ws[wid] + inc - v;
}

// bits [x0, x0+len) of a bitmap row (len <= 32), row has at least one spare word after the last used one
__device__ __forceinline__ uint32_t row_bits(const uint32_t *row, int x0, int len) {
    const int wi = x0 >> 5, sh = x0 & 31;
    uint64_t v = (uint64_t) row[wi] | ((uint64_t) row[wi + 1] << 32);
    return (uint32_t) (v >> sh) & (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u));
}
__device__ __forceinline__ uint32_t test_bit(const uint32_t *row, int x) { return (row[x >> 5] >> (x & 31)) & 1u; }
#endif
