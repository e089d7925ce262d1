// Dependent-issue latency micro-benchmark (one warp per SM, one dependent chain): cycles per instruction.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
constexpr int ITER = 4096;

template <int MODE>
__global__ void lat_kernel(double* out, long long* cyc, int src) {
    __shared__ double sh[64];
    sh[threadIdx.x & 63] = threadIdx.x;
    __syncthreads();
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0000001;
    double c0 = a, c1 = b;
    unsigned u = threadIdx.x + 1;
    float f = a;
    int idx = threadIdx.x & 31;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < ITER; ++i) {
        if (MODE == 0) a = fma(a, b, 1e-9);                                   // DFMA chain
        if (MODE == 1) a = a * b;                                             // DMUL chain
        if (MODE == 2) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(b), "d"(b));
        if (MODE == 3) u = __reduce_max_sync(0xffffffffu, u + 1);             // CREDUX chain
        if (MODE == 4) a = __shfl_sync(0xffffffffu, a, (src + i) & 31);       // SHFL 64-bit chain
        if (MODE == 5) { asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(a) : "d"(a)); }   // MUFU.RCP64H chain
        if (MODE == 6) { idx = (int)sh[idx] & 31; }                           // LDS.64 pointer chase (+cvt)
        if (MODE == 7) f = fmaf(f, 1.0000001f, 1e-9f);                        // FFMA chain
        if (MODE == 8) u = (u * 3u + 1u) ^ (u >> 3);                          // integer ALU chain (IMAD + SHF + LOP3)
        if (MODE == 9) { a = a > 0.5 ? a * b : a + b; }                       // DSETP + select + op
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + c0 + c1 + u + f + idx;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, double* out, long long* cyc, int per_iter) {
    lat_kernel<MODE><<<148, 32>>>(out, cyc, 1);
    CK(cudaDeviceSynchronize());
    lat_kernel<MODE><<<148, 32>>>(out, cyc, 1);
    CK(cudaDeviceSynchronize());
    long long h; CK(cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    printf("%-40s %7.1f cycles per iteration (%d dependent op%s)\n", name, (double)h / ITER, per_iter, per_iter > 1 ? "s" : "");
}

int main() {
    double* out; long long* cyc;
    CK(cudaMalloc(&out, sizeof(double) * 148 * 32)); CK(cudaMalloc(&cyc, 8));
    run<0>("DFMA dependent chain", out, cyc, 1);
    run<1>("DMUL dependent chain", out, cyc, 1);
    run<2>("DMMA m8n8k4 dependent chain", out, cyc, 1);
    run<3>("CREDUX.MAX (+IADD) chain", out, cyc, 2);
    run<4>("SHFL.IDX 64-bit chain", out, cyc, 2);
    run<5>("MUFU.RCP64H chain", out, cyc, 1);
    run<6>("LDS.64 + F2I pointer chase", out, cyc, 3);
    run<7>("FFMA dependent chain", out, cyc, 1);
    run<8>("IMAD+SHF+LOP3 chain", out, cyc, 3);
    run<9>("DSETP+DMUL/DADD+select chain", out, cyc, 3);
    return 0;
}
