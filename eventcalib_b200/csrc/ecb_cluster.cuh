// Device-side structures shared by the window, cluster and pair kernels.
#pragma once
#include "ecb_common.cuh"

#define ECB_CL_THREADS 512
#define ECB_MAXK_LIMIT 512

// one clustering problem = one (window, polarity) point set, or one ecb_dbscan_run input
struct ProbDesc {
    int64_t off;      // offset of the problem's points in pix[pol] / labels[pol] / kmem[pol]
    int32_t n;        // number of points (pid = 0..n-1)
    int32_t pol;      // which flat array
    int32_t x0, y0;   // origin subtracted before indexing the bitmap
    int32_t pad0, pad1;
};

struct ProbHdr {
    int32_t n_clusters;  // raw clusters (discovery order ids 0..n_clusters-1)
    int32_t n_kept;      // clusters with size >= cluster_min (clipped to max_k)
    int32_t n_core;
    uint32_t status;     // ECB_PB_*
};

struct KeptCluster {
    int32_t raw_id;
    int32_t size;
    int32_t med_pid;   // member with the median norm (slot size/2 of the members sorted by (norm^2, pid))
    int32_t med_x, med_y;
    int32_t mem_off;   // offset of the member list inside the problem's kmem slice
    double m[9];       // Sx Sy Sxx Syy Sxy Sxxx Syyy Sxyy Sxxy (exact integers)
};

struct ClusterArgs {
    const ProbDesc *prob;
    int n_prob;
    unsigned *work_counter;
    const uint32_t *pix[2];
    int32_t *labels[2];
    uint32_t *kmem[2];
    ProbHdr *hdr;
    KeptCluster *ktab;
    int max_k;
    uint32_t *gscratch;         // per-CTA global scratch when the per-point arrays do not fit shared memory
    size_t gscratch_stride;     // words per CTA
    int arrays_in_smem;
    int n_cap;                  // region size (>= max n over the batch, >= 64)
    int W, H, E, PW, PH;        // bitmap: W x H pixels, E = floor(eps) padding, PW words per padded row, PH rows
    int eps_int;                // eps if it is an integer (the kd tie rule can fire), else -1
    uint32_t min_pts, cluster_min;
    int8_t halfw[ECB_MAX_EPS + 1];  // half width of the eps-disc at |dy|
};

size_t ecb_cluster_smem_bytes(int PW, int PH, int n_cap, bool arrays_in_smem, bool rank32);
int ecb_launch_cluster(ecb_ctx *ctx, ClusterArgs &a, int max_n);
