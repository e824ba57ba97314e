// Measures the FP64 peaks of the GPU the cost-evaluation kernel is bounded by:
//   DFMA  (CUDA-core fma.rn.f64, 8 independent chains per thread)
//   DMMA  (mma.sync.aligned.m8n8k4.row.col.f64, 8 independent accumulator tiles per warp)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, int iters) {
    double a[8], b = 1.0000001, c = 1e-9;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double *out, int iters) {
    double acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Do the two pipes overlap?  Even warps run the DMMA loop, odd warps the DFMA loop, each with its solo iteration count:
// if DMMA and DFMA shared one pipe the mixed kernel would take (t_dmma + t_dfma) / 2, if they are independent max() / ... of
// the halves.  (ncu lists them as separate pipes: sm__pipe_fp64_cycles_active vs sm__pipe_tensor_subpipe_dmma_cycles_active.)
__global__ void k_mixed(double *out, int iters_dfma, int iters_dmma) {
    const int warp = threadIdx.x >> 5;
    double s = 0;
    if (warp & 1) {
        double a[8], b = 1.0000001, c = 1e-9;
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
        for (int it = 0; it < iters_dfma; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) s += a[i];
    } else {
        double acc[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.0;
        double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
        for (int it = 0; it < iters_dmma; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 4, threads = 512, iters = 20000;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_dfma<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8 * iters * (double) blocks * threads;
        printf("DFMA: %.3f ms  %.2f TFLOP/s\n", ms, fl / ms / 1e9);
    }
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_dmma<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double) blocks * (threads / 32);
        printf("DMMA m8n8k4: %.3f ms  %.2f TFLOP/s\n", ms, fl / ms / 1e9);
    }
    // mixed: half the warps each; iteration counts chosen so that each half alone would take about the same time
    const int it_dfma = 20000, it_dmma = 20000 / 8;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_mixed<<<blocks, threads>>>(out, it_dfma, 0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms_f, ms_m, ms_b;
        cudaEventElapsedTime(&ms_f, e0, e1);
        cudaEventRecord(e0);
        k_mixed<<<blocks, threads>>>(out, 0, it_dmma);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms_m, e0, e1);
        cudaEventRecord(e0);
        k_mixed<<<blocks, threads>>>(out, it_dfma, it_dmma);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms_b, e0, e1);
        const double fl = 2.0 * 8 * it_dfma * (double) blocks * (threads / 2) + 2.0 * 8 * 8 * 4 * 8 * it_dmma * (double) blocks * (threads / 64);
        printf("mixed (half the warps each): DFMA half alone %.3f ms, DMMA half alone %.3f ms, both %.3f ms  -> %.2f TFLOP/s combined "
               "(shared pipe would take %.3f ms, independent pipes %.3f ms)\n", ms_f, ms_m, ms_b, fl / ms_b / 1e9, ms_f + ms_m,
               ms_f > ms_m ? ms_f : ms_m);
    }
    printf("device: %s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    return 0;
}
