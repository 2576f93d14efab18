// Fused residual/Jacobian kernels for the synthetic workloads (see
// include/blsq_models.h): user-side callbacks, one thread per residual.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/blsq.h"
#include "../../include/blsq_models.h"

namespace {

__global__ void __launch_bounds__(256)
expdecay2_kernel(int64_t A, const int64_t* __restrict__ idx, int m,
                 const double* __restrict__ t, const double* __restrict__ X,
                 const double* __restrict__ y, double* __restrict__ F,
                 double* __restrict__ J) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= A * m) return;
    int64_t s = g / m;
    int r = (int)(g % m);
    int64_t pid = idx ? idx[s] : s;
    const double4 x = *reinterpret_cast<const double4*>(X + s * 4);
    const double tr = t[r];
    // same operation order as synthetic.ExpDecay2.fun_t / jac_t
    const double e1 = exp(-x.y * tr);
    const double e2 = exp(-x.w * tr);
    F[g] = x.x * e1 + x.z * e2 - __ldcs(y + pid * m + r);
    if (J) {
        double2* jp = reinterpret_cast<double2*>(J + g * 4);
        __stcs(jp, make_double2(e1, -x.x * tr * e1));
        __stcs(jp + 1, make_double2(e2, -x.z * tr * e2));
    }
}

__global__ void __launch_bounds__(256)
gausspeak_kernel(int64_t A, const int64_t* __restrict__ idx, int m,
                 const double* __restrict__ t, const double* __restrict__ X,
                 const double* __restrict__ y, double* __restrict__ F) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= A * m) return;
    int64_t s = g / m;
    int r = (int)(g % m);
    int64_t pid = idx ? idx[s] : s;
    const double* x = X + s * 6;
    const double tr = t[r];
    // same operation order as synthetic.GaussPeak.fun_t
    const double z = (tr - x[1]) / x[2];
    F[g] = x[0] * exp(-0.5 * z * z) + x[3] + x[4] * tr + x[5] * tr * tr -
           __ldcs(y + pid * m + r);
}

}  // namespace

extern "C" {

int blsq_model_expdecay2(int64_t A, const int64_t* idx, int m, const double* t,
                         const double* X, const double* y, double* F, double* J,
                         void* stream) {
    if (A < 0 || m < 1 || !t || !X || !y || !F) return BLSQ_E_BADARG;
    if (A == 0) return 0;
    int64_t blocks = (A * m + 255) / 256;
    if (blocks > 0x7fffffff) return BLSQ_E_UNSUPPORTED;
    expdecay2_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        A, idx, m, t, X, y, F, J);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

int blsq_model_gausspeak(int64_t A, const int64_t* idx, int m, const double* t,
                         const double* X, const double* y, double* F,
                         void* stream) {
    if (A < 0 || m < 1 || !t || !X || !y || !F) return BLSQ_E_BADARG;
    if (A == 0) return 0;
    int64_t blocks = (A * m + 255) / 256;
    if (blocks > 0x7fffffff) return BLSQ_E_UNSUPPORTED;
    gausspeak_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        A, idx, m, t, X, y, F);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

}  // extern "C"
