// Tall mode: the n x n tail of a trust-region round as ONE CTA (latency bound).
// All of the mathematics is in blsq_tall_core.cuh; this file carves the shared
// memory, runs the requested phase and exposes the C ABI.
//
//   phase 0  init     trf.py:201 / dogbox.py:131
//   phase 1  judge    trf.py:310-344 / dogbox.py:222-251 (+ accept 346-352 / 253-261)
//   phase 2  propose  trf.py:238-308 / dogbox.py:164-220
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/blsq.h"
#include "blsq_tall_core.cuh"

using namespace blsq_tall;

#define BLSQ_LAUNCH_CHECK()                                  \
    do {                                                     \
        cudaError_t e_ = cudaGetLastError();                 \
        if (e_ != cudaSuccess) return (int)e_;               \
    } while (0)

namespace {

constexpr int ROUND_SMEM_MAX_N = 128;     // Jacobi work matrix in shared memory up to here

struct RoundSmem {
    size_t vec, red, rowbuf, A, ints, total;
    __host__ __device__ RoundSmem(int n, int nwarps) {
        vec = 0;
        red = vec + TallWork::doubles(n) * 8;
        rowbuf = red + (size_t)4 * 32 * 8;
        A = rowbuf + (size_t)nwarps * n * 8;
        ints = A + (n <= ROUND_SMEM_MAX_N ? (size_t)n * n * 8 : 0);
        total = ints + ((size_t)5 * n + 1 + 64) * 4;
    }
};

__global__ void __launch_bounds__(1024, 1)
tall_round_kernel(int phase, TallParams P, int first, int new_lin, int nranks,
                  const double* __restrict__ ssq_parts, const double* __restrict__ fac,
                  const double* __restrict__ x0, const double* __restrict__ lb,
                  const double* __restrict__ ub, const double* __restrict__ scaling,
                  double* __restrict__ st, int* __restrict__ ist, double* __restrict__ gwork) {
    extern __shared__ __align__(16) unsigned char rsm[];
    const int n = P.n;
    Blk B;
    B.tid = threadIdx.x;
    B.nt = blockDim.x;
    B.lane = threadIdx.x & 31;
    B.warp = threadIdx.x >> 5;
    B.nwarps = blockDim.x >> 5;
    B.lanes = 32;
    const RoundSmem SM(n, B.nwarps);
    B.red = reinterpret_cast<double*>(rsm + SM.red);
    int* ints = reinterpret_cast<int*>(rsm + SM.ints);
    B.ired = ints + 5 * n + 1;

    if (phase == 0) {
        tall_init(B, P.method, n, x0, lb, ub, st, ist);
        return;
    }
    if (phase == 1) {
        double obj_new = 0.0;
        for (int r = 0; r < nranks; r++) obj_new += ssq_parts[r];     // rank order
        tall_judge(B, P, obj_new, first, lb, ub, st, ist);
        return;
    }
    TallWork W;
    W.carve(reinterpret_cast<double*>(rsm + SM.vec), n);
    W.rowbuf = reinterpret_cast<double*>(rsm + SM.rowbuf);
    W.hits = ints;
    W.flags = ints + n;
    W.fr = ints + 2 * n;
    W.marks = ints + 3 * n;
    W.prog = ints + 4 * n;
    double* A = (n <= ROUND_SMEM_MAX_N) ? reinterpret_cast<double*>(rsm + SM.A) : gwork;
    W.A = A;
    for (int i = B.tid; i < n; i += B.nt) { W.lb[i] = lb[i]; W.ub[i] = ub[i]; }
    B.sync();
    if (P.method == BLSQ_METHOD_TRF)
        tall_trf_propose(B, P, W, fac, x0, scaling, first, new_lin, A, st, ist);
    else
        tall_dogbox_propose(B, P, W, fac, x0, scaling, first, new_lin, A, st, ist);
}

}  // namespace

extern "C" {

int blsq_tall_layout(int n, int64_t* out) {
    if (!out) return BLSQ_E_BADARG;
    if (n < 2 || n > BLSQ_MAX_TALL_N) return BLSQ_E_UNSUPPORTED;
    const TallLayout L(n);
    const FacLayout FL(n);
    out[0] = L.SIZE;  out[1] = L.ISIZE; out[2] = L.X;    out[3] = L.XNEW;
    out[4] = TS_OBJ;  out[5] = TS_DELTA; out[6] = TS_GNORM; out[7] = L.ONB;
    out[8] = FL.SIZE; out[9] = FL.R;    out[10] = FL.QTF; out[11] = FL.G;
    out[12] = FL.OBJ; out[13] = FL.INFO; out[14] = FL.RINVP; out[15] = L.SCALE;
    out[16] = FL.REFINE;
    return 0;
}

int blsq_tall_round(int method, int phase, int n, int64_t m_total, int nranks,
                    const double* ssq_parts, const double* fac, const double* x0,
                    const double* lb, const double* ub, const double* scaling, double ftol,
                    double xtol, double gtol, int max_nfev, int first, int new_lin,
                    double* state, int32_t* istate, double* work, void* stream) {
    if (method != BLSQ_METHOD_TRF && method != BLSQ_METHOD_DOGBOX) return BLSQ_E_BADARG;
    if (phase < 0 || phase > 2 || !x0 || !lb || !ub || !state || !istate) return BLSQ_E_BADARG;
    if (n < 2 || n > BLSQ_MAX_TALL_N) return BLSQ_E_UNSUPPORTED;
    if (phase == 1 && (!ssq_parts || nranks < 1)) return BLSQ_E_BADARG;
    if (phase == 2 && (!fac || (n > ROUND_SMEM_MAX_N && !work))) return BLSQ_E_BADARG;
    TallParams P;
    P.ftol = ftol; P.xtol = xtol; P.gtol = gtol;
    P.max_nfev = max_nfev; P.m = (double)m_total;
    P.jac_scaling = scaling ? 0 : 1; P.n = n; P.method = method;
    const int threads = (n > 32) ? 1024 : 256;
    const RoundSmem SM(n, threads / 32);
    cudaError_t e = cudaFuncSetAttribute(tall_round_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)RoundSmem(ROUND_SMEM_MAX_N, 32).total);
    if (e != cudaSuccess) return (int)e;
    tall_round_kernel<<<1, threads, SM.total, (cudaStream_t)stream>>>(
        phase, P, first, new_lin, nranks, ssq_parts, fac, x0, lb, ub, scaling, state, istate,
        work);
    BLSQ_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
