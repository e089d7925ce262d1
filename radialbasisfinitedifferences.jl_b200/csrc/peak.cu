// peak.cu -- live measurement of the FP64 roofline denominator (MEASURED_PEAKS.json records HBM and bf16 only).
// Two dependent-chain micro-kernels at full occupancy: scalar DFMA and DMMA (mma.sync m8n8k4 f64), best of 5,
// timed with CUDA events on the context's stream.  Same kernels as tools/fp64_peak.cu.
#include "common.cuh"

namespace {

template <int ILP>
__global__ void peak_dfma_kernel(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void peak_dmma_kernel(double* out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int rbffd_measure_fp64_peak(rbffd_context* ctx, double* dfma_tflops, double* dmma_tflops) {
    if (!ctx) return RBFFD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int threads = 256, blocks = ctx->sm_count * 8, iters = 8192;
    DevBuf<double> out;
    CUDA_TRY(ctx, out.alloc((size_t)threads * blocks, st));
    double best[2] = {0, 0};
    for (int which = 0; which < 2; ++which)
        for (int rep = 0; rep < 7; ++rep) {
            CUDA_TRY(ctx, cudaEventRecord(ctx->ev[6], st));
            if (which == 0) peak_dfma_kernel<8><<<blocks, threads, 0, st>>>(out.p, iters, 1.0000001, 1e-9);
            else peak_dmma_kernel<4><<<blocks, threads, 0, st>>>(out.p, iters, 1.0000001, 1e-9);
            CUDA_TRY(ctx, cudaEventRecord(ctx->ev[7], st));
            CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev[7]));
            float ms;
            CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]));
            double flops = which == 0 ? 2.0 * 8 * iters * (double)threads * blocks
                                      : 2.0 * 256 * 4 * iters * (double)threads * blocks / 32;
            if (rep >= 2) best[which] = std::max(best[which], flops / ms * 1e-9);
        }
    if (dfma_tflops) *dfma_tflops = best[0];
    if (dmma_tflops) *dmma_tflops = best[1];
    return RBFFD_OK;
}
