// Elementwise bound-geometry passes and finite-difference points: pure
// HBM-bandwidth kernels, bit-exact with bounds.py / dogbox.py:9-35 / scipy
// _numdiff (see blsq_core.cuh for the arithmetic, include/blsq.h for the ABI).
//
// Layout: x is (B, n) row-major; one thread per element so a warp touches 32
// consecutive doubles.  The two row reductions (step_size_to_bound's min and
// in_bounds' all) give each row to a power-of-two group of lanes and reduce
// with shuffles.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/blsq.h"
#include "blsq_core.cuh"

using namespace blsq;

#define BLSQ_LAUNCH_CHECK()                                  \
    do {                                                     \
        cudaError_t e_ = cudaGetLastError();                 \
        if (e_ != cudaSuccess) return (int)e_;               \
    } while (0)

namespace {

__device__ __forceinline__ int64_t gtid() {
    return (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
}

int group_for(int n) {
    int G = 1;
    while (G < 32 && G < n) G <<= 1;
    return G;
}

__global__ void step_size_kernel(int64_t B, int n, int G,
                                 const double* __restrict__ x,
                                 const double* __restrict__ d,
                                 const double* __restrict__ lb,
                                 const double* __restrict__ ub, int bstride,
                                 double* __restrict__ step,
                                 int64_t* __restrict__ hits) {
    int64_t t = gtid();
    int64_t b = t / G;
    int lane = (int)(t % G);
    bool valid = b < B;
    // pass 1: row minimum (NaN-propagating like np.min)
    // NumPy's sequential minimum keeps the LATER element on ties, which
    // decides the sign of a zero step; carry the index to reproduce it
    double tmin = dinf();
    int imin = -1;
    bool has_nan = false;
    if (valid) {
        for (int i = lane; i < n; i += G) {
            double di = d[b * n + i];
            double ti = dinf();
            if (di != 0) {
                double xi = x[b * n + i];
                ti = np_max((lb[b * bstride + i] - xi) / di,
                            (ub[b * bstride + i] - xi) / di);
            }
            if (ti != ti) has_nan = true;
            if (!(tmin < ti)) { tmin = ti; imin = i; }
        }
    }
    for (int off = G >> 1; off > 0; off >>= 1) {
        double o = __shfl_xor_sync(0xffffffffu, tmin, off, 32);
        int oi = __shfl_xor_sync(0xffffffffu, imin, off, 32);
        bool on = __shfl_xor_sync(0xffffffffu, (int)has_nan, off, 32);
        if (o < tmin || (o == tmin && oi > imin)) { tmin = o; imin = oi; }
        has_nan = has_nan || on;
    }
    if (has_nan) tmin = dnan();
    if (!valid) return;
    if (lane == 0) step[b] = tmin;
    if (hits) {
        for (int i = lane; i < n; i += G) {
            double di = d[b * n + i];
            double ti = dinf();
            if (di != 0) {
                double xi = x[b * n + i];
                ti = np_max((lb[b * bstride + i] - xi) / di,
                            (ub[b * bstride + i] - xi) / di);
            }
            hits[b * n + i] = (ti == tmin) ? isign(di) : 0;
        }
    }
}

__global__ void in_bounds_kernel(int64_t B, int n, int G,
                                 const double* __restrict__ x,
                                 const double* __restrict__ lb,
                                 const double* __restrict__ ub, int bstride,
                                 uint8_t* __restrict__ ok) {
    int64_t t = gtid();
    int64_t b = t / G;
    int lane = (int)(t % G);
    bool valid = b < B;
    int good = 1;
    if (valid) {
        for (int i = lane; i < n; i += G) {
            double xi = x[b * n + i];
            good &= (int)((xi >= lb[b * bstride + i]) & (xi <= ub[b * bstride + i]));
        }
    }
    for (int off = G >> 1; off > 0; off >>= 1)
        good &= __shfl_xor_sync(0xffffffffu, good, off, 32);
    if (valid && lane == 0) ok[b] = (uint8_t)good;
}

__global__ void active_kernel(int64_t total, int n, const double* __restrict__ x,
                              const double* __restrict__ lb,
                              const double* __restrict__ ub, int bstride,
                              double rtol, int64_t* __restrict__ mask) {
    int64_t t = gtid();
    if (t >= total) return;
    int64_t b = t / n;
    int i = (int)(t % n);
    mask[t] = active_constraint(x[t], lb[b * bstride + i], ub[b * bstride + i], rtol);
}

__global__ void feasible_kernel(int64_t total, int n, const double* __restrict__ x,
                                const double* __restrict__ lb,
                                const double* __restrict__ ub, int bstride,
                                double rstep, double* __restrict__ out) {
    int64_t t = gtid();
    if (t >= total) return;
    int64_t b = t / n;
    int i = (int)(t % n);
    out[t] = strictly_feasible(x[t], lb[b * bstride + i], ub[b * bstride + i], rstep);
}

__global__ void cl_kernel(int64_t total, int n, const double* __restrict__ x,
                          const double* __restrict__ g,
                          const double* __restrict__ lb,
                          const double* __restrict__ ub, int bstride,
                          double* __restrict__ v, double* __restrict__ jv) {
    int64_t t = gtid();
    if (t >= total) return;
    int64_t b = t / n;
    int i = (int)(t % n);
    double vv, jj;
    cl_scaling(x[t], g[t], lb[b * bstride + i], ub[b * bstride + i], vv, jj);
    v[t] = vv;
    jv[t] = jj;
}

__global__ void intersection_kernel(int64_t total, int n,
                                    const double* __restrict__ x,
                                    const double* __restrict__ tr,
                                    const double* __restrict__ lb,
                                    const double* __restrict__ ub, int bstride,
                                    double* __restrict__ lo,
                                    double* __restrict__ hi,
                                    uint8_t* __restrict__ flags) {
    int64_t t = gtid();
    if (t >= total) return;
    int64_t b = t / n;
    int i = (int)(t % n);
    double l, h;
    int f = find_intersection(x[t], tr[t], lb[b * bstride + i], ub[b * bstride + i], l, h);
    lo[t] = l;
    hi[t] = h;
    flags[t] = (uint8_t)f;
}

// one thread per (i, slot, k): Xp[i, slot, k] = x_k (+ h_i when k == i)
__global__ void fd2_points_kernel(int64_t A, const int32_t* __restrict__ idx,
                                  int n, const double* __restrict__ x,
                                  const double* __restrict__ lb,
                                  const double* __restrict__ ub, int bstride,
                                  double rel_step, double* __restrict__ Xp,
                                  double* __restrict__ dx) {
    int64_t t = gtid();
    int64_t total = A * n * n;
    if (t >= total) return;
    int k = (int)(t % n);
    int64_t slot = (t / n) % A;
    int i = (int)(t / ((int64_t)n * A));
    double xk = x[slot * n + k];
    if (k == i) {
        int64_t pid = idx ? idx[slot] : slot;
        double h = fd2_step(xk, lb[pid * bstride + k], ub[pid * bstride + k], rel_step);
        double xp = xk + h;
        dx[slot * n + i] = xp - xk;
        xk = xp;
    }
    Xp[t] = xk;
}

// 3-point scheme: batches 2i and 2i+1 are the two evaluation points of
// coordinate i -- (x - h, x + h) central, (x + h, x + 2h) one sided;
// dxo[slot, i] = the denominator, dxo[slot, n + i] = 1.0 when one sided
__global__ void fd3_points_kernel(int64_t A, const int32_t* __restrict__ idx,
                                  int n, const double* __restrict__ x,
                                  const double* __restrict__ lb,
                                  const double* __restrict__ ub, int bstride,
                                  double rel_step, double* __restrict__ Xp,
                                  double* __restrict__ dxo) {
    int64_t t = gtid();
    int64_t total = A * n * n;
    if (t >= total) return;
    int k = (int)(t % n);
    int64_t slot = (t / n) % A;
    int i = (int)(t / ((int64_t)n * A));
    double xk = x[slot * n + k];
    double x1 = xk, x2 = xk;
    if (k == i) {
        int64_t pid = idx ? idx[slot] : slot;
        bool one;
        double h = fd3_step(xk, lb[pid * bstride + k], ub[pid * bstride + k], rel_step, one);
        if (one) {
            x1 = xk + h;
            x2 = xk + 2 * h;
            dxo[slot * 2 * n + i] = x2 - xk;
        } else {
            x1 = xk - h;
            x2 = xk + h;
            dxo[slot * 2 * n + i] = x2 - x1;
        }
        dxo[slot * 2 * n + n + i] = one ? 1.0 : 0.0;
    }
    Xp[((int64_t)(2 * i) * A + slot) * n + k] = x1;
    Xp[((int64_t)(2 * i + 1) * A + slot) * n + k] = x2;
}

}  // namespace

// x_covariance (least_squares.py:248-252: the inverse of J^T J) from the
// triangular factor J = Q R that the solve already holds: (R^T R)^-1 =
// R^-1 R^-T.  One thread per problem, R given as the packed upper triangle
// (row i holds columns i..n-1) at `r_off` of a record of `stride` doubles, or
// dense row-major (packed = 0).  A zero pivot ("the inverse doesn't exist")
// gives a NaN matrix.
__global__ void covariance_kernel(int64_t B, int n, const double* __restrict__ rec,
                                  int64_t stride, int r_off, int packed,
                                  double* __restrict__ cov) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double* R = rec + b * stride + r_off;
    double* C = cov + b * (int64_t)n * n;
    auto r = [&](int i, int j) -> double {
        return packed ? R[i * n - (i * (i - 1)) / 2 + (j - i)] : R[(int64_t)i * n + j];
    };
    // pivots at rounding level of the largest one: J^T J has no inverse
    double dmax = 0.0;
    for (int i = 0; i < n; i++) dmax = fabs(r(i, i)) > dmax ? fabs(r(i, i)) : dmax;
    bool singular = false;
    for (int i = 0; i < n; i++)
        singular = singular || !(fabs(r(i, i)) > 16.0 * n * 2.220446049250313e-16 * dmax);
    if (singular) {
        for (int e = 0; e < n * n; e++) C[e] = nan("");
        return;
    }
    // T = R^-1 (upper) stored in C, column by column (back substitution)
    for (int j = 0; j < n; j++) {
        for (int i = n - 1; i >= 0; i--) {
            double v = 0.0;
            if (i == j) v = 1.0 / r(i, i);
            else if (i < j) {
                double acc = 0.0;
                for (int k = i + 1; k <= j; k++) acc = fma(r(i, k), C[(int64_t)k * n + j], acc);
                v = -acc / r(i, i);
            }
            C[(int64_t)i * n + j] = v;
        }
    }
    // C = T T^T in place, row by row from the top: entry (i, j), i <= j, needs
    // rows i and j of T from column j on, which are not overwritten yet when
    // the lower triangle is filled by symmetry afterwards
    for (int i = 0; i < n; i++) {
        for (int j = i; j < n; j++) {
            double acc = 0.0;
            for (int k = j; k < n; k++) acc = fma(C[(int64_t)i * n + k], C[(int64_t)j * n + k], acc);
            C[(int64_t)i * n + j] = acc;          // (i, j) is read no more: k >= j > ... only later columns
        }
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++) C[(int64_t)i * n + j] = C[(int64_t)j * n + i];
}

extern "C" {

int blsq_version(void) { return BLSQ_VERSION; }

const char* blsq_error_string(int code) {
    if (code == 0) return "ok";
    if (code == BLSQ_E_BADARG) return "invalid argument";
    if (code == BLSQ_E_UNSUPPORTED) return "unsupported size";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown error";
}

#define BLSQ_CHECK_COMMON(B, n, bstride)                        \
    if ((B) < 0 || (n) < 1) return BLSQ_E_BADARG;               \
    if ((bstride) != 0 && (bstride) != (n)) return BLSQ_E_BADARG; \
    if ((B) == 0) return 0;

int blsq_step_size_to_bound(int64_t B, int n, const double* x, const double* d,
                            const double* lb, const double* ub, int bstride,
                            double* step, int64_t* hits, void* stream) {
    if (!x || !d || !lb || !ub || !step) return BLSQ_E_BADARG;
    BLSQ_CHECK_COMMON(B, n, bstride)
    int G = group_for(n);
    int64_t blocks = (B * G + 255) / 256;
    step_size_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        B, n, G, x, d, lb, ub, bstride, step, hits);
    BLSQ_LAUNCH_CHECK();
    return 0;
}

int blsq_in_bounds(int64_t B, int n, const double* x, const double* lb,
                   const double* ub, int bstride, uint8_t* ok, void* stream) {
    if (!x || !lb || !ub || !ok) return BLSQ_E_BADARG;
    BLSQ_CHECK_COMMON(B, n, bstride)
    int G = group_for(n);
    int64_t blocks = (B * G + 255) / 256;
    in_bounds_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        B, n, G, x, lb, ub, bstride, ok);
    BLSQ_LAUNCH_CHECK();
    return 0;
}

int blsq_find_active_constraints(int64_t B, int n, const double* x,
                                 const double* lb, const double* ub,
                                 int bstride, double rtol, int64_t* mask,
                                 void* stream) {
    if (!x || !lb || !ub || !mask) return BLSQ_E_BADARG;
    BLSQ_CHECK_COMMON(B, n, bstride)
    int64_t total = B * n;
    active_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        total, n, x, lb, ub, bstride, rtol, mask);
    BLSQ_LAUNCH_CHECK();
    return 0;
}

int blsq_make_strictly_feasible(int64_t B, int n, const double* x,
                                const double* lb, const double* ub, int bstride,
                                double rstep, double* out, void* stream) {
    if (!x || !lb || !ub || !out) return BLSQ_E_BADARG;
    BLSQ_CHECK_COMMON(B, n, bstride)
    int64_t total = B * n;
    feasible_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        total, n, x, lb, ub, bstride, rstep, out);
    BLSQ_LAUNCH_CHECK();
    return 0;
}

int blsq_scaling_vector(int64_t B, int n, const double* x, const double* g,
                        const double* lb, const double* ub, int bstride,
                        double* v, double* jv, void* stream) {
    if (!x || !g || !lb || !ub || !v || !jv) return BLSQ_E_BADARG;
    BLSQ_CHECK_COMMON(B, n, bstride)
    int64_t total = B * n;
    cl_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        total, n, x, g, lb, ub, bstride, v, jv);
    BLSQ_LAUNCH_CHECK();
    return 0;
}

int blsq_find_intersection(int64_t B, int n, const double* x, const double* tr,
                           const double* lb, const double* ub, int bstride,
                           double* lo, double* hi, uint8_t* flags,
                           void* stream) {
    if (!x || !tr || !lb || !ub || !lo || !hi || !flags) return BLSQ_E_BADARG;
    BLSQ_CHECK_COMMON(B, n, bstride)
    int64_t total = B * n;
    intersection_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        total, n, x, tr, lb, ub, bstride, lo, hi, flags);
    BLSQ_LAUNCH_CHECK();
    return 0;
}

int blsq_fd2_points(int64_t A, const int32_t* idx, int n, const double* x,
                    const double* lb, const double* ub, int bstride,
                    double rel_step, double* Xp, double* dx, void* stream) {
    if (!x || !lb || !ub || !Xp || !dx) return BLSQ_E_BADARG;
    BLSQ_CHECK_COMMON(A, n, bstride)
    int64_t total = A * n * n;
    fd2_points_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        A, idx, n, x, lb, ub, bstride, rel_step, Xp, dx);
    BLSQ_LAUNCH_CHECK();
    return 0;
}

int blsq_fd3_points(int64_t A, const int32_t* idx, int n, const double* x,
                    const double* lb, const double* ub, int bstride,
                    double rel_step, double* Xp, double* dxo, void* stream) {
    if (!x || !lb || !ub || !Xp || !dxo) return BLSQ_E_BADARG;
    BLSQ_CHECK_COMMON(A, n, bstride)
    int64_t total = A * n * n;
    fd3_points_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        A, idx, n, x, lb, ub, bstride, rel_step, Xp, dxo);
    BLSQ_LAUNCH_CHECK();
    return 0;
}

int blsq_covariance(int64_t B, int n, const double* rec, int64_t stride, int r_off,
                    int packed, double* cov, void* stream) {
    if (B < 0 || n < 1 || n > 256 || !rec || !cov || stride < 0 || r_off < 0) return BLSQ_E_BADARG;
    if (B == 0) return 0;
    int64_t blocks = (B + 127) / 128;
    covariance_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(B, n, rec, stride, r_off,
                                                                          packed, cov);
    BLSQ_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
