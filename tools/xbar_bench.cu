// Crossbar micro-benchmark for B200: cost (SM clocks per warp-instruction, SM-wide) of the shared-memory / shuffle
// patterns the weight kernel can use, measured at full occupancy.  Guides the kernel design.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int ITER = 2048;
constexpr int UNR = 8;

// mode: 0 LDS.64 conflict-free (lane*8B), 1 LDS.128 conflict-free (lane*16B), 2 LDS.64 uniform address,
// 3 LDS.128 uniform address, 4 LDS.64 4-distinct (lane&3), 5 LDS.128 4-distinct, 6 SHFL.IDX 32-bit, 7 SHFL 64-bit (2x),
// 8 REDUX.max u32, 9 STS.64 conflict-free, 10 STS.128 conflict-free, 11 LDS.32 conflict-free, 12 LDS.64 stride 50 doubles (row-per-lane),
// 13 LDS.128 stride 50 doubles, 14 DFMA only, 15 LDS.32 uniform
template <int MODE>
__global__ void xbar_kernel(double* out, int src_lane) {
    __shared__ __align__(16) double sh[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = i * 1e-3;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc = lane;
    double2 acc2 = make_double2(0, 0);
    unsigned ui = lane + 1;
    float facc = 0;
    int base = warp * 8;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int off = (base + u * 2 + it) & 255;
            if (MODE == 0) acc += sh[off + lane];
            if (MODE == 1) { double2 v = *reinterpret_cast<double2*>(&sh[(off & 254) + lane * 2]); acc2.x += v.x; acc2.y += v.y; }
            if (MODE == 2) acc += sh[off];
            if (MODE == 3) { double2 v = *reinterpret_cast<double2*>(&sh[off & 254]); acc2.x += v.x; acc2.y += v.y; }
            if (MODE == 4) acc += sh[off + (lane & 3) * 8];
            if (MODE == 5) { double2 v = *reinterpret_cast<double2*>(&sh[(off & 254) + (lane & 3) * 8]); acc2.x += v.x; acc2.y += v.y; }
            if (MODE == 6) ui += __shfl_sync(0xffffffffu, ui, (src_lane + u) & 31);
            if (MODE == 7) acc += __shfl_sync(0xffffffffu, acc, (src_lane + u) & 31);
            if (MODE == 8) ui += __reduce_max_sync(0xffffffffu, ui + u);
            if (MODE == 9) sh[off + lane] = acc + u;
            if (MODE == 10) *reinterpret_cast<double2*>(&sh[(off & 254) + lane * 2]) = make_double2(acc, acc + u);
            if (MODE == 11) facc += reinterpret_cast<float*>(sh)[off + lane];
            if (MODE == 12) acc += sh[(off & 31) + lane * 50];
            if (MODE == 13) { double2 v = *reinterpret_cast<double2*>(&sh[(off & 30) + lane * 50]); acc2.x += v.x; acc2.y += v.y; }
            if (MODE == 14) acc = fma(acc, 1.0000001, 1e-9);
            if (MODE == 15) facc += reinterpret_cast<float*>(sh)[off];
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + acc2.x + acc2.y + ui + facc;
}

template <int MODE>
void run(const char* name, double* out, int sms, double clk_ghz) {
    const int threads = 256, blocks = sms * 4;    // 32 warps per SM
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    xbar_kernel<MODE><<<blocks, threads>>>(out, 3);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        xbar_kernel<MODE><<<blocks, threads>>>(out, 3);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    // warp-instructions per SM = 32 warps * ITER * UNR
    double instr = 32.0 * ITER * UNR;
    double clk = best * 1e-3 * clk_ghz * 1e9;
    printf("%-44s %8.3f ms  %6.2f SM-clk per warp-instruction\n", name, best, clk / instr);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    double* out; CK(cudaMalloc(&out, sizeof(double) * 1024 * 1024));
    double ghz = prop.clockRate * 1e-6;
    printf("%s  sms=%d  clock=%.3f GHz (nominal max; divide by the real clock under load)\n", prop.name, prop.multiProcessorCount, ghz);
    int sms = prop.multiProcessorCount;
    run<14>("DFMA (dependent chain, 32 warps/SM)", out, sms, ghz);
    run<0>("LDS.64  conflict-free", out, sms, ghz);
    run<1>("LDS.128 conflict-free", out, sms, ghz);
    run<11>("LDS.32  conflict-free", out, sms, ghz);
    run<2>("LDS.64  uniform address", out, sms, ghz);
    run<3>("LDS.128 uniform address", out, sms, ghz);
    run<15>("LDS.32  uniform address", out, sms, ghz);
    run<4>("LDS.64  4 distinct addresses", out, sms, ghz);
    run<5>("LDS.128 4 distinct addresses", out, sms, ghz);
    run<12>("LDS.64  row-per-lane stride 50 doubles", out, sms, ghz);
    run<13>("LDS.128 row-per-lane stride 50 doubles", out, sms, ghz);
    run<9>("STS.64  conflict-free", out, sms, ghz);
    run<10>("STS.128 conflict-free", out, sms, ghz);
    run<6>("SHFL.IDX 32-bit", out, sms, ghz);
    run<7>("SHFL.IDX 64-bit (2 SHFL)", out, sms, ghz);
    run<8>("REDUX.MAX u32", out, sms, ghz);
    return 0;
}
