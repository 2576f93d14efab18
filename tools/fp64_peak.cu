// FP64 peak microbenchmark for sm_100a: DFMA (vector pipe) and DMMA.8x8x4
// (mma.sync.aligned.m8n8k4.f64, the only FP64 tensor instruction sm_100a has;
// m16n8k{4,8,16}.f64 compile to sequences of DMMA.8x8x4).
// Prints one JSON line per variant.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ACC>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a0, double b0) {
    double c0[ACC], c1[ACC];
#pragma unroll
    for (int i = 0; i < ACC; i++) { c0[i] = threadIdx.x; c1[i] = i; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ACC; i++) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACC; i++) s += c0[i] + c1[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a0, double b0) {
    double c[ACC];
#pragma unroll
    for (int i = 0; i < ACC; i++) c[i] = threadIdx.x + i;
    double a = a0, b = b0 + threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ACC; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACC; i++) s += c[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F launch, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * 256 * sms * 8));
    const int iters = 20000;
    int bps_list[] = {1, 2, 4};
    for (int bi = 0; bi < 3; bi++) {
        int bps = bps_list[bi];
        int grid = sms * bps;
        {
            float ms = time_it([&] { k_dmma<8><<<grid, 256>>>(out, iters, 1.0000001, 0.5); }, 5);
            double fl = (double)grid * 8 /*warps*/ * iters * 8 * 512.0;
            printf("{\"kind\": \"dmma884\", \"acc_per_warp\": 8, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n",
                   8 * bps, ms, fl / ms / 1e9);
        }
        {
            float ms = time_it([&] { k_dmma<16><<<grid, 256>>>(out, iters, 1.0000001, 0.5); }, 5);
            double fl = (double)grid * 8 * iters * 16 * 512.0;
            printf("{\"kind\": \"dmma884\", \"acc_per_warp\": 16, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n",
                   8 * bps, ms, fl / ms / 1e9);
        }
        {
            float ms = time_it([&] { k_dfma<16><<<grid, 256>>>(out, iters, 1.0000001, 0.5); }, 5);
            double fl = (double)grid * 256 * iters * 16 * 2.0;
            printf("{\"kind\": \"dfma\", \"acc_per_thread\": 16, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n",
                   8 * bps, ms, fl / ms / 1e9);
        }
    }
    CK(cudaDeviceSynchronize());
    printf("{\"sms\": %d, \"clock_khz\": %d}\n", sms, p.clockRate);
    return 0;
}
