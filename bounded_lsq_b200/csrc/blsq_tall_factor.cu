// Tall mode, the n x n factor step of CholeskyQR2 (one CTA, latency bound).
//
//   pass 1:  G = sum over ranks of J_s^T J_s (all rows, or a sample of the row
//            tiles)  ->  R1 = chol(G) (upper), R1^-1: the preconditioner
//   pass 2:  G2 = sum over ranks of Y^T Y, Y = J R1^-1 (every row)  ->
//            R2 = chol(G2), R = R2 R1 (J = Q R), Q^T f = R2^-T (Y^T f),
//            g = J^T f (trf.py:244 / dogbox.py:170), f.f; and the check that
//            G2 is close enough to the identity for ONE Cholesky pass to be
//            accurate (Gershgorin on the unit-diagonal scaling: off-diagonal
//            row sums <= 1/2  =>  scaled condition number <= 3).  If
//            not, fac.refine = 1 and the driver runs
//   pass 3:  R1 <- R, R1^-1  (no Cholesky), then gram(2) + pass 2 again.
//
// The rank partials are summed in rank order, so every rank of a row-sharded
// run computes bit-identical factors (no broadcast needed).
//
// `fac` layout: blsq_tall_common.cuh (FacLayout).  R1^-1 is emitted in the
// DMMA fragment order the pass-2 Gram kernel consumes: for every upper 8x8
// block (kb <= jb), two k-halves of 32 lanes, lane -> element
// (8 kb + 4 half + lane%4, 8 jb + lane/4).
// info = 0, or 1000*pass + k + 1 when the Cholesky pivot k was not positive
// (J numerically rank deficient: CholeskyQR2 does not apply).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/blsq.h"
#include "blsq_tall_common.cuh"

#define BLSQ_LAUNCH_CHECK()                                  \
    do {                                                     \
        cudaError_t e_ = cudaGetLastError();                 \
        if (e_ != cudaSuccess) return (int)e_;               \
    } while (0)

namespace {

constexpr int FAC_THREADS = 1024;
constexpr int FAC_SMEM_MAX_N = 128;      // n x n doubles in shared memory up to here

// sqrt(d) and 1/sqrt(d) from one reciprocal square root and a Newton step each
// (within ~1 ulp; d > 0)
__device__ __forceinline__ void sqrt_terms(double d, double& r, double& inv_r) {
    double y = rsqrt(d);
    double rr = d * y;
    rr = fma(fma(-rr, rr, d), 0.5 * y, rr);
    y = fma(fma(-rr, y, 1.0), y, y);
    r = rr;
    inv_r = y;
}

// in-place upper Cholesky of the upper triangle of M (n x n, row-major):
// M = R^T R.  diag0[k] = the diagonal of the matrix before any shift.  A pivot
// that has collapsed to rounding level (<= 8 n eps diag0[k]) stops the sweep:
// returns k + 1 (0 = success) and the caller retries with a diagonal shift.
// 32 x 32 threads over the trailing block (no index division), two barriers
// per column: the pivot itself stays in place until the sweep is over (dg
// collects the diagonal of R), so nobody has to wait for the others to have
// read it.
__device__ int chol_upper(double* M, const double* diag0, double* dg, int n) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int tx = tid & 31, ty = tid >> 5, ny = nt >> 5;
    const double tol = 8.0 * n * 2.220446049250313e-16;
    for (int k = 0; k < n; k++) {
        const double piv = M[k * n + k];
        if (!(piv > tol * diag0[k])) return k + 1;   // uniform: everyone reads the same value
        double r, inv_r;
        sqrt_terms(piv, r, inv_r);
        if (tid == 0) dg[k] = r;
        for (int j = k + 1 + tid; j < n; j += nt) M[k * n + j] *= inv_r;
        __syncthreads();
        for (int i = k + 1 + ty; i < n; i += ny) {
            const double ri = M[k * n + i];
            for (int j = k + 1 + tx; j < n; j += 32)
                if (j >= i) M[i * n + j] = fma(-ri, M[k * n + j], M[i * n + j]);
        }
        __syncthreads();
    }
    for (int k = tid; k < n; k += nt) M[k * n + k] = dg[k];
    __syncthreads();
    return 0;
}

// X = R^-1 for upper triangular R; one warp per column (back substitution).
// The column being built stays in registers -- lane l owns its entries
// k = l, l + 32, ... (n <= 256: 8 per lane) -- so that a step never waits for
// a value the previous step has just written to memory; invd[i] = 1 / R_ii.
// Entries below the diagonal are zeroed.
__device__ void inv_upper(const double* R, double* X, double* invd, int n) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < n; i += blockDim.x) invd[i] = 1.0 / R[i * n + i];
    __syncthreads();
    for (int j = warp; j < n; j += nw) {
        double xc[8];
#pragma unroll
        for (int q = 0; q < 8; q++) xc[q] = (lane + 32 * q == j) ? invd[j] : 0.0;
        for (int i = j - 1; i >= 0; i--) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int k = lane + 32 * q;
                if (k > i && k <= j) s = fma(R[i * n + k], xc[q], s);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            const double xi = -s * invd[i];
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (lane + 32 * q == i) xc[q] = xi;
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int k = lane + 32 * q;
            if (k < n) X[k * n + j] = xc[q];          // zero below the diagonal
        }
    }
}

// dense upper-triangular X (n x n) -> the DMMA fragment order of pass 2
__device__ void pack_rinv(const double* X, double* rinvp, int n, int nb) {
    const int nblk = nb * (nb + 1) / 2;
    for (int e = threadIdx.x; e < nblk * 64; e += blockDim.x) {
        const int q = e >> 6, half = (e >> 5) & 1, lane = e & 31;
        int kb = 0, qq = q;
        while (qq >= nb - kb) { qq -= nb - kb; kb++; }
        const int jb = kb + qq;
        const int r = 8 * kb + 4 * half + (lane & 3), c = 8 * jb + (lane >> 2);
        rinvp[e] = (r < n && c < n && c >= r) ? X[r * n + c] : 0.0;
    }
}

__global__ void __launch_bounds__(FAC_THREADS, 1)
tall_factor_kernel(int pass, int n, int nranks, int64_t gstride, const double* __restrict__ grams,
                   double* __restrict__ fac, int use_smem) {
    extern __shared__ __align__(16) double fsm[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int n2 = n * n;
    const blsq_tall::FacLayout FL(n);
    double* R1 = fac + FL.R1;
    double* R = fac + FL.R;
    double* scratch = fac + FL.SCR;
    double* qtf = fac + FL.QTF;
    double* g = fac + FL.G;
    double* obj = fac + FL.OBJ;
    double* info = fac + FL.INFO;
    double* rinvp = fac + FL.RINVP;
    double* shifts = fac + FL.SHIFT;
    double* refine = fac + FL.REFINE;
    double* M = use_smem ? fsm : scratch;
    __shared__ double diag0[256];
    __shared__ double wrk[256];              // diag(R) / reciprocals / 1/sqrt(diag G)
    __shared__ double dmax_s;

    if (pass == 3) {
        // the factor found so far becomes the preconditioner of another pass
        for (int e = tid; e < n2; e += nt) { R1[e] = R[e]; M[e] = R[e]; }
        __syncthreads();
        double* X = use_smem ? scratch : R;
        inv_upper(M, X, wrk, n);
        __syncthreads();
        pack_rinv(X, rinvp, n, FL.nb);
        if (tid == 0) { refine[0] = 0.0; info[0] = 0.0; }
        return;
    }
    // Cholesky of the rank-order sum of the records.  A rank-deficient (or
    // worse than ~1e7 conditioned) Jacobian makes a pivot collapse; the sweep
    // is then redone on G + shift*I, shift = 16 n eps max(diag G) (x10 per
    // retry).  With both passes shifted R^T R = J^T J (1 + O(shift2)) +
    // O(shift1 shift2) I, i.e. the null directions of J come out with a
    // singular value of ~1e-13 |J| instead of 0 and everything else is
    // unchanged to ~1e-14 (shifted CholeskyQR, Fukaya et al. 2020).
    double shift = 0.0;
    int bad = 0;
    for (int attempt = 0; attempt < 12; attempt++) {
        const int nvec = (pass == 2 && attempt == 0) ? 2 * n + 1 : 0;
        for (int e = tid; e < n2 + nvec; e += nt) {
            double s = 0.0;
            for (int r = 0; r < nranks; r++) s += grams[(size_t)r * gstride + e];
            if (e < n2) {
                const int i = e / n, j = e % n;
                if (i == j) { if (attempt == 0) diag0[i] = s; s += shift; }
                M[e] = s;
            } else if (e < n2 + n) {
                qtf[e - n2] = s;                     // Y^T f for now
            } else if (e == n2 + n) {
                obj[0] = s;
            } else {
                g[e - n2 - n - 1] = s;
            }
        }
        __syncthreads();
        if (attempt == 0 && tid == 0) {
            double d = 0.0;
            for (int i = 0; i < n; i++) d = diag0[i] > d ? diag0[i] : d;
            dmax_s = d;
        }
        __syncthreads();
        if (attempt == 0 && pass == 2) {
            // is G2 close to (a multiple of) the identity?  Columns whose
            // diagonal is at rounding level are the null directions of a rank
            // deficient J (their Y column is noise): no pass can fix those,
            // they are left out of the test.
            const double tiny = 1e-8 * dmax_s;
            for (int i = tid; i < n; i += nt) wrk[i] = (diag0[i] > tiny) ? rsqrt(diag0[i]) : 0.0;
            __syncthreads();
            int far = 0;
            const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
            for (int i = warp; i < n; i += nw) {       // one warp per row
                double rs = 0.0;
                for (int j = lane; j < n; j += 32) {
                    if (j == i) continue;
                    const double gij = (j > i) ? M[i * n + j] : M[j * n + i];
                    rs += fabs(gij) * wrk[j];
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
                if (wrk[i] > 0.0 && !(rs * wrk[i] <= 0.5)) far = 1;
            }
            // (the spread of the diagonal itself does not matter: Gram and
            // Cholesky errors scale with sqrt(G_ii G_jj) componentwise)
            far = __syncthreads_or(far);
            if (tid == 0) refine[0] = far ? 1.0 : 0.0;
        }
        __syncthreads();
        bad = chol_upper(M, diag0, wrk, n);
        if (!bad) break;
        if (!(dmax_s > 0.0) || dmax_s != dmax_s) break;      // zero or NaN Jacobian
        shift = (shift == 0.0) ? 16.0 * n * 2.220446049250313e-16 * dmax_s : shift * 10.0;
        __syncthreads();
    }
    if (tid == 0) {
        shifts[pass - 1] = shift;
        // a pivot that needed a shift means J itself is rank deficient:
        // another pass cannot improve on that
        if (pass == 2 && shift != 0.0) refine[0] = 0.0;
    }
    if (bad) {
        if (tid == 0) info[0] = 1000.0 * pass + bad;
        return;
    }
    if (pass == 1) {
        for (int e = tid; e < n2; e += nt) {
            const int i = e / n, j = e % n;
            R1[e] = (j >= i) ? M[e] : 0.0;
        }
        __syncthreads();
        // dense inverse into a free n x n region, then the fragment order
        double* X = use_smem ? scratch : R;
        inv_upper(M, X, wrk, n);
        __syncthreads();
        pack_rinv(X, rinvp, n, FL.nb);
        if (tid == 0) info[0] = 0.0;
        return;
    }
    // pass 2: R = R2 R1 (upper x upper) by all warps but the last, which does
    // Q^T f = R2^-T z (forward substitution with R2^T) at the same time
    for (int i = tid; i < n; i += nt) wrk[i] = 1.0 / M[i * n + i];
    __syncthreads();
    const int nprod = nt - 32;
    if (tid < nprod) {
        for (int e = tid; e < n2; e += nprod) {
            const int i = e / n, j = e % n;
            double s = 0.0;
            if (j >= i)
                for (int k = i; k <= j; k++) s = fma(M[i * n + k], R1[k * n + j], s);
            R[e] = s;
        }
    } else {
        // lane l owns entries k = l, l + 32, ... of the solution (registers)
        const int lane = tid & 31;
        double qc[8];
#pragma unroll
        for (int q = 0; q < 8; q++) qc[q] = (lane + 32 * q < n) ? qtf[lane + 32 * q] : 0.0;
        // right-looking: once entry i is final it is taken out of the later
        // ones along ROW i of R2 (contiguous in memory, no reduction)
        for (int i = 0; i < n; i++) {
            double v = 0.0;
#pragma unroll
            for (int q = 0; q < 8; q++)
                if ((i >> 5) == q) v = qc[q] * wrk[i];
            v = __shfl_sync(0xffffffffu, v, i & 31);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int k = lane + 32 * q;
                if (k == i) qc[q] = v;
                else if (k > i && k < n) qc[q] = fma(-M[i * n + k], v, qc[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 8; q++)
            if (lane + 32 * q < n) qtf[lane + 32 * q] = qc[q];
    }
    if (tid == 0) info[0] = 0.0;
}

}  // namespace

extern "C" {

int64_t blsq_tall_fac_size(int n) {
    if (n < 2 || n > 256) return BLSQ_E_UNSUPPORTED;
    return blsq_tall::FacLayout(n).SIZE;
}

int blsq_tall_factor(int pass, int n, int nranks, int64_t gstride, const double* grams,
                     double* fac, void* stream) {
    if (pass < 1 || pass > 3 || nranks < 1 || !fac || (pass != 3 && !grams)) return BLSQ_E_BADARG;
    if (n < 2 || n > 256) return BLSQ_E_UNSUPPORTED;
    if (pass != 3 && gstride < (int64_t)n * n + 2 * n + 1) return BLSQ_E_BADARG;
    const int use_smem = n <= FAC_SMEM_MAX_N;
    const size_t smem = use_smem ? (size_t)n * n * 8 : 0;
    cudaError_t e = cudaFuncSetAttribute(tall_factor_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         FAC_SMEM_MAX_N * FAC_SMEM_MAX_N * 8);
    if (e != cudaSuccess) return (int)e;
    tall_factor_kernel<<<1, FAC_THREADS, smem, (cudaStream_t)stream>>>(pass, n, nranks, gstride,
                                                                        grams, fac, use_smem);
    BLSQ_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
