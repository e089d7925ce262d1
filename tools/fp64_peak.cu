// FP64 pipe micro-benchmark for B200 (sm_100a): DFMA and DMMA (mma.sync f64) issue rates.
// Used to obtain the MEASURED FP64 roofline denominator (MEASURED_PEAKS.json has none).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
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

// mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 : 256 FMA per warp instruction
template <int ILP>
__global__ void dmma884_kernel(double* out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

#ifdef TRY_M16
// m16n8k8: 1024 FMA per warp instruction
template <int ILP>
__global__ void dmma1688_kernel(double* out, int iters, double a, double b) {
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#endif

// mixed: DFMA interleaved with LDS.128 broadcast loads and integer ops, to see co-issue headroom
template <int ILP>
__global__ void dfma_lds_kernel(double* out, int iters, double a, double b) {
    __shared__ double sh[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = i * 1e-6;
    __syncthreads();
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    int base = (threadIdx.x & 3) * 8;
    for (int it = 0; it < iters; ++it) {
        double2 v0 = *reinterpret_cast<double2*>(&sh[(base + it * 2) & 1022]);
        double2 v1 = *reinterpret_cast<double2*>(&sh[(base + 64 + it * 2) & 1022]);
#pragma unroll
        for (int i = 0; i < ILP; i += 4) {
            acc[i] = fma(acc[i], v0.x, b);
            acc[i + 1] = fma(acc[i + 1], v0.y, b);
            acc[i + 2] = fma(acc[i + 2], v1.x, b);
            acc[i + 3] = fma(acc[i + 3], v1.y, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
double time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", prop.name, sms, prop.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * 1024 * 2048));
    const int iters = 4096;
    for (int wps = 4; wps <= 32; wps *= 2) {       // warps per SM
        int threads = 256, blocks = sms * wps * 32 / threads;
        {
            constexpr int ILP = 8;
            double ms = time_ms([&] { dfma_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double flops = 2.0 * ILP * iters * (double)blocks * threads;
            printf("DFMA   warps/SM %2d ILP %d : %8.3f ms  %7.2f TFLOP/s\n", wps, ILP, ms, flops / ms * 1e-9);
        }
        {
            constexpr int ILP = 8;
            double ms = time_ms([&] { dfma_lds_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double flops = 2.0 * ILP * iters * (double)blocks * threads;
            printf("DFMA+LDS warps/SM %2d ILP %d : %8.3f ms  %7.2f TFLOP/s\n", wps, ILP, ms, flops / ms * 1e-9);
        }
        {
            constexpr int ILP = 4;
            double ms = time_ms([&] { dmma884_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double flops = 2.0 * 256 * ILP * iters * (double)blocks * threads / 32;
            printf("DMMA884 warps/SM %2d ILP %d : %8.3f ms  %7.2f TFLOP/s\n", wps, ILP, ms, flops / ms * 1e-9);
        }
#ifdef TRY_M16
        {
            constexpr int ILP = 4;
            double ms = time_ms([&] { dmma1688_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double flops = 2.0 * 1024 * ILP * iters * (double)blocks * threads / 32;
            printf("DMMA1688 warps/SM %2d ILP %d : %8.3f ms  %7.2f TFLOP/s\n", wps, ILP, ms, flops / ms * 1e-9);
        }
#endif
    }
    // sustained: ~2 s of back-to-back DFMA
    {
        constexpr int ILP = 8;
        int threads = 256, blocks = sms * 16 * 32 / threads;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int n = 400;
        for (int i = 0; i < n; ++i) dfma_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * ILP * iters * (double)blocks * threads * n;
        printf("DFMA sustained %d launches: %8.1f ms  %7.2f TFLOP/s\n", n, ms, flops / ms * 1e-9);
    }
    return 0;
}
