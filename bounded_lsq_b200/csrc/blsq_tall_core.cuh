// Tall mode: the n x n tail of one trust-region round for general n
// (8 < n <= 256), written once for a whole thread block.
//
// This is the block-parallel restatement of trf_round / dogbox_round in
// blsq_core.cuh (same stage boundaries, same quirks, same reference lines):
// vectors of n live in (shared) memory and are processed with
// `for (i = tid; i < n; i += nt)` loops, reductions go through blk_* helpers,
// matrices (R from the CholeskyQR2 factor record, the Jacobi work matrix) are
// swept one row per warp.  On the device a Blk describes the CUDA block; on the
// host (tests/host_emul, TEST ONLY) it is a single "thread" with one lane, so
// the very same control flow runs serially and can be checked against the
// oracle without a GPU.  Summation orders differ between the two builds (and
// from the reference's BLAS), the branch logic does not.
//
// Reference being restated (nmayorov/bounded-lsq):
//   trf.py:201-358          trf (judge = 310-344, propose = 238-308)
//   dogbox.py:131-272       dogbox
//   trust_region.py:11-152  intersect_trust_region, solve_lsq_trust_region
//   bounds.py:24-149        step_size_to_bound, make_strictly_feasible,
//                           scaling_vector
#pragma once
#include "blsq_core.cuh"
#include "blsq_tall_common.cuh"

#if defined(__CUDA_ARCH__)
#define BLSQ_TALL_DEV 1
#else
#define BLSQ_TALL_DEV 0
#endif

namespace blsq_tall {

using namespace blsq;

// ---- state records -----------------------------------------------------------
enum { TS_OBJ = 0, TS_DELTA, TS_ALPHA, TS_PRED, TS_CORR, TS_NSTEPH, TS_NSTEP, TS_GNORM,
       TS_GNSTEP,        // |p_gn| of the cached Gauss-Newton step (TI_GN = 1)
       TS_NSCAL = 16 };
enum { TI_STATUS = 0, TI_NFEV, TI_NJEV, TI_ACCEPT, TI_PENDING, TI_TRHIT, TI_SWEEPS,
       TI_GN,            // 1: the state holds the Gauss-Newton step (in SUF), no SVD yet
       TI_NSCAL = 8 };

struct TallLayout {
    int n;
    int64_t X, XNEW, SCALE, S, SUF, VT, SIZE;     // doubles
    int ONB, MARKS, FREE, ISIZE;                  // ints
    BLSQ_HD explicit TallLayout(int n_) : n(n_) {
        X = TS_NSCAL;
        XNEW = X + n;
        SCALE = XNEW + n;
        S = SCALE + n;
        SUF = S + n;
        VT = SUF + n;
        SIZE = VT + (int64_t)n * n;
        ONB = TI_NSCAL;
        MARKS = ONB + n;
        FREE = MARKS + n;
        ISIZE = FREE + n;
    }
};

struct TallParams {
    double ftol, xtol, gtol;
    int max_nfev;
    double m;            // total number of residuals over all ranks
    int jac_scaling;
    int n;
    int method;
};

// ---- the block ---------------------------------------------------------------
struct Blk {
    int tid, nt, lane, warp, nwarps, lanes;
    double* red;         // 4 * nwarps doubles
    int* ired;           // nwarps + 8 ints
    BLSQ_HD void sync() const {
#if BLSQ_TALL_DEV
        __syncthreads();
#endif
    }
};

BLSQ_HD double warp_sum(double v) {
#if BLSQ_TALL_DEV
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
#endif
    return v;
}

// all threads of the block call these together; every thread gets the result
BLSQ_HD double blk_sum(const Blk& B, double v) {
#if BLSQ_TALL_DEV
    v = warp_sum(v);
    if (B.lane == 0) B.red[B.warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < B.nwarps; w++) s += B.red[w];
    __syncthreads();
    return s;
#else
    return v;
#endif
}
BLSQ_HD void blk_sum3(const Blk& B, double& a, double& b, double& c) {
#if BLSQ_TALL_DEV
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (B.lane == 0) { B.red[3 * B.warp] = a; B.red[3 * B.warp + 1] = b; B.red[3 * B.warp + 2] = c; }
    __syncthreads();
    double sa = 0.0, sb = 0.0, sc = 0.0;
    for (int w = 0; w < B.nwarps; w++) { sa += B.red[3 * w]; sb += B.red[3 * w + 1]; sc += B.red[3 * w + 2]; }
    __syncthreads();
    a = sa; b = sb; c = sc;
#endif
}
BLSQ_HD double blk_min(const Blk& B, double v) {      // plain minimum (+inf identity)
#if BLSQ_TALL_DEV
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double o = __shfl_xor_sync(0xffffffffu, v, off);
        v = o < v ? o : v;
    }
    if (B.lane == 0) B.red[B.warp] = v;
    __syncthreads();
    double s = B.red[0];
    for (int w = 1; w < B.nwarps; w++) s = B.red[w] < s ? B.red[w] : s;
    __syncthreads();
    return s;
#else
    return v;
#endif
}
BLSQ_HD double blk_max(const Blk& B, double v) {
    return -blk_min(B, -v);
}
BLSQ_HD bool blk_any(const Blk& B, bool f) {
#if BLSQ_TALL_DEV
    return __syncthreads_or(f ? 1 : 0) != 0;
#else
    return f;
#endif
}

// out[i] = sum_{j in [lo_i, n)} M[i*n + j] * v[j], lo_i = i (upper) or 0 (full);
// one warp per row, lanes over j.  Caller syncs before using `out`.
BLSQ_HD void rows_dot(const Blk& B, const double* M, int n, bool upper, const double* v,
                      double* out) {
    for (int i = B.warp; i < n; i += B.nwarps) {
        double s = 0.0;
        for (int j = (upper ? i : 0) + B.lane; j < n; j += B.lanes) s = fma(M[(size_t)i * n + j], v[j], s);
        s = warp_sum(s);
        if (B.lane == 0) out[i] = s;
    }
}
// out[i] = sum_j Mt[j*n + i] * w[j]   (V w from Vt), one thread per i
BLSQ_HD void cols_dot(const Blk& B, const double* Mt, int n, const double* w, double* out,
                      double sign) {
    for (int i = B.tid; i < n; i += B.nt) {
        double s = 0.0;
        for (int j = 0; j < n; j++) s = fma(Mt[(size_t)j * n + i], w[j], s);
        out[i] = sign * s;
    }
}
BLSQ_HD double vdot(const Blk& B, const double* a, const double* b, int n) {
    double s = 0.0;
    for (int i = B.tid; i < n; i += B.nt) s = fma(a[i], b[i], s);
    return blk_sum(B, s);
}

// ---- work space (shared memory on the device) ------------------------------------
struct TallWork {
    int n;
    double *x, *g, *qtf, *lb, *ub, *scale, *v, *jv, *d, *g_h, *diag_h, *s, *suf, *w, *p_h, *p,
        *r_h, *r, *x_edge, *refl, *c_h, *ng, *t1, *t2, *t3, *step_h, *step, *tv, *b;
    int* hits;
    int* flags;
    int* fr;
    int* marks;
    double* A;           // n x n Jacobi work matrix (shared up to n = 128, else global)
    double* rowbuf;      // fold-in rows, one per warp
    volatile int* prog;  // fold-in progress (n + 1 ints)
    static constexpr int NVEC = 29;
    BLSQ_HD static size_t doubles(int n) { return (size_t)NVEC * n; }
    BLSQ_HD void carve(double* base, int n_) {
        n = n_;
        double** slots[NVEC] = {&x, &g, &qtf, &lb, &ub, &scale, &v, &jv, &d, &g_h, &diag_h, &s,
                                &suf, &w, &p_h, &p, &r_h, &r, &x_edge, &refl, &c_h, &ng, &t1,
                                &t2, &t3, &step_h, &step, &tv, &b};
        for (int k = 0; k < NVEC; k++) *slots[k] = base + (size_t)k * n;
    }
};

// bounds.py:24-48 over the block.  `msk` (nullable): coordinates with msk == 0
// are ignored (dogbox works on the free sub-vector).  tv = per-coordinate steps.
BLSQ_HD double tall_step_to_bound(const Blk& B, int n, const double* x, const double* dvec,
                                  const double* lo, const double* hi, const int* msk,
                                  double* tv, int* hits) {
    double tmin = dinf();
    bool has_nan = false;
    for (int i = B.tid; i < n; i += B.nt) {
        double t = dinf();
        if ((!msk || msk[i]) && dvec[i] != 0)
            t = np_max((lo[i] - x[i]) / dvec[i], (hi[i] - x[i]) / dvec[i]);
        tv[i] = t;
        if (t != t) has_nan = true;
        else if (t < tmin) tmin = t;
    }
    tmin = blk_min(B, tmin);
    if (blk_any(B, has_nan)) tmin = dnan();           // np.min propagates NaN
    if (hits) {
        for (int i = B.tid; i < n; i += B.nt)
            hits[i] = ((!msk || msk[i]) && tv[i] == tmin) ? isign(dvec[i]) : 0;
    }
    B.sync();
    return tmin;
}

// y = R (dvec o s) for the upper-triangular R of the factor record, i.e.
// J_h s in the rotated frame (|J_h s| = |R diag(d) s|).  dvec nullable.
BLSQ_HD void tall_Rh_matvec(const Blk& B, const double* R, int n, const double* dvec,
                            const double* s, double* tmp, double* y) {
    for (int i = B.tid; i < n; i += B.nt) tmp[i] = dvec ? dvec[i] * s[i] : s[i];
    B.sync();
    rows_dot(B, R, n, true, tmp, y);
    B.sync();
}

// trf.py:79-102 for one step
BLSQ_HD double tall_eval_quadratic(const Blk& B, const double* R, const TallWork& W,
                                   const double* s) {
    const int n = W.n;
    tall_Rh_matvec(B, R, n, W.d, s, W.t1, W.t2);
    double vv = 0.0, sd = 0.0, sg = 0.0;
    for (int i = B.tid; i < n; i += B.nt) {
        vv = fma(W.t2[i], W.t2[i], vv);
        sd = fma(W.diag_h[i], s[i] * s[i], sd);
        sg = fma(s[i], W.g_h[i], sg);
    }
    blk_sum3(B, vv, sd, sg);
    return 0.5 * (vv + sd) + sg;
}

// trf.py:37-76
BLSQ_HD void tall_build_quadratic_1d(const Blk& B, const double* R, const TallWork& W,
                                     const double* s, const double* s0, double& a, double& b) {
    const int n = W.n;
    tall_Rh_matvec(B, R, n, W.d, s, W.t1, W.t2);          // t2 = J_h s
    double vv = 0.0, sd = 0.0, gs = 0.0;
    for (int i = B.tid; i < n; i += B.nt) {
        vv = fma(W.t2[i], W.t2[i], vv);
        sd = fma(s[i] * W.diag_h[i], s[i], sd);
        gs = fma(W.g_h[i], s[i], gs);
    }
    blk_sum3(B, vv, sd, gs);
    a = 0.5 * (vv + sd);
    b = gs;
    if (s0) {
        tall_Rh_matvec(B, R, n, W.d, s0, W.t1, W.t3);     // t3 = J_h s0
        double uv = 0.0, s0d = 0.0, z = 0.0;
        for (int i = B.tid; i < n; i += B.nt) {
            uv = fma(W.t3[i], W.t2[i], uv);
            s0d = fma(s0[i] * W.diag_h[i], s[i], s0d);
        }
        blk_sum3(B, uv, s0d, z);
        b += uv + s0d;
    }
}

// ---- one-sided Jacobi on the rows of A (n x n) with the same rotations on b ----
// Round-robin ordering: in step st of a sweep the N/2 pairs are disjoint.  A
// pair is handled by a GROUP of G lanes (G = 4..32, about n/8: each lane owns
// ~8 elements of both rows, so the three dot products cost log2(G) shuffle
// rounds), 32/G pairs per warp; only the warps that have pairs take part and
// they synchronise on a named barrier, the rest of the block waits at the end.
// Rows of A -> s_j v_j^T, b -> U^T b (see jacobi_rows in blsq_core.cuh).
BLSQ_HD double group_sum(double v, int G) {
#if BLSQ_TALL_DEV
    for (int off = G >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
#endif
    return v;
}

// `valid` = this group has a real pair (all lanes of the warp call this).
BLSQ_HD bool jacobi_pair(double* A, double* b, int n, int p, int q, bool valid, int gl, int G) {
    double* ap = A + (size_t)p * n;
    double* aq = A + (size_t)q * n;
    double al = 0.0, be = 0.0, ga = 0.0;
    if (valid) {
        for (int i = gl; i < n; i += G) {
            const double x = ap[i], y = aq[i];
            al = fma(x, x, al);
            be = fma(y, y, be);
            ga = fma(x, y, ga);
        }
    }
    al = group_sum(al, G); be = group_sum(be, G); ga = group_sum(ga, G);
    if (!valid || ga == 0.0 || ga * ga <= (EPS * EPS) * (al * be)) return false;
    // tan of the rotation angle, t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)) with
    // zeta = (be - al) / (2 ga), rearranged to one division and one square root
    const double dd = be - al;
    const double g2 = 2.0 * ga;
#if BLSQ_TALL_DEV
    // The angle only has to be accurate enough to make progress (Jacobi is
    // self correcting; the convergence test above is in double): evaluate the
    // tangent in single precision after scaling (dd, g2) by a power of two,
    // then build an exactly orthogonal (c, sn) pair in double.  This takes the
    // double division and square root off the per-step critical path.
    const double big = fmax(fabs(dd), fabs(g2));
    const int ex = (__double2hiint(big) >> 20) & 0x7ff;
    const double scl = __hiloint2double((2046 - ex) << 20, 0);      // 2^(1023 - ex)
    const float df = (float)(dd * scl), gf = (float)(g2 * scl);
    float tf = copysignf(gf, df * gf) / (fabsf(df) + sqrtf(fmaf(df, df, gf * gf)));
    if (df == 0.0f) tf = copysignf(1.0f, gf);
    const double t = (ex == 0 || ex == 0x7ff) ? copysign(g2, dd * ga) /
                         (fabs(dd) + sqrt(fma(dd, dd, g2 * g2)))
                                              : (double)tf;
#else
    const double t = copysign(g2, dd * ga) / (fabs(dd) + sqrt(fma(dd, dd, g2 * g2)));
#endif
    const double c = rsqrt_d(fma(t, t, 1.0));
    const double sn = c * t;
    for (int i = gl; i < n; i += G) {
        const double x = ap[i], y = aq[i];
        ap[i] = fma(c, x, -(sn * y));
        aq[i] = fma(sn, x, c * y);
    }
    if (gl == 0) {
        const double bp = b[p], bq = b[q];
        b[p] = fma(c, bp, -(sn * bq));
        b[q] = fma(sn, bp, c * bq);
    }
    return true;
}

// Returns the number of sweeps (diagnostics: istate[TI_SWEEPS]).
BLSQ_HD int tall_jacobi(const Blk& B, double* A, double* b, int n) {
    if (n < 2) return 0;
    const int N = (n + 1) & ~1;          // players; index n (if any) is a bye
    const int npairs = N / 2;
#if BLSQ_TALL_DEV
    int G = 4;
    while (G < 32 && G * 8 < n) G <<= 1;
    const int ppw = 32 / G;                              // pairs per warp
    int nact = (npairs + ppw - 1) / ppw;                 // warps that have pairs
    if (nact > B.nwarps) nact = B.nwarps;
    const int gl = B.lane & (G - 1);
    const int slot = B.warp * ppw + B.lane / G;          // first pair of this group
    const int nslots = nact * ppw;
    volatile int* flag = B.ired;                         // [0], [1]: "rotated" per sweep parity
    if (B.tid < 2) flag[B.tid] = 0;
    __syncthreads();
#else
    const int G = 1, gl = 0, slot = 0, nslots = 1, nact = 1;
#endif
    int sweep = 0;
    int sweeps = 0;
    if (B.warp < nact) {
        for (; sweep < 60; sweep++) {
            bool rotated = false;
            for (int st = 0; st < N - 1; st++) {
                for (int k0 = 0; k0 < npairs; k0 += nslots) {
                    const int k = k0 + slot;
                    int p = N - 1, q = st;
                    if (k != 0) {
                        p = st + k; if (p >= N - 1) p -= N - 1;
                        q = st - k; if (q < 0) q += N - 1;
                    }
                    if (p > q) { const int tmp = p; p = q; q = tmp; }
                    const bool valid = k < npairs && q < n;
                    rotated = jacobi_pair(A, b, n, valid ? p : 0, valid ? q : 0, valid, gl, G) ||
                              rotated;
                }
#if BLSQ_TALL_DEV
                asm volatile("bar.sync 1, %0;" ::"r"(nact * 32) : "memory");
                // everyone has read last sweep's flag before this barrier
                if (st == 0 && B.tid == 0) flag[(sweep + 1) & 1] = 0;
#endif
            }
#if BLSQ_TALL_DEV
            if (rotated) flag[sweep & 1] = 1;
            asm volatile("bar.sync 1, %0;" ::"r"(nact * 32) : "memory");
            rotated = flag[sweep & 1] != 0;
#endif
            if (!rotated) break;
        }
        sweeps = sweep + 1;
#if BLSQ_TALL_DEV
        if (B.tid == 0) B.ired[2] = sweeps;
#endif
    }
    B.sync();
#if BLSQ_TALL_DEV
    sweeps = B.ired[2];
#endif
    return sweeps;
}

// Fold row k of diag(sqrt(diag_h)) into the upper triangle A (and [b; 0]) by
// Givens rotations (hat_svd in blsq_core.cuh).  Fold k only touches rows
// i >= k of A, and needs the previous non-trivial fold to be done with row i
// first: the warps run the folds as a wavefront.  prog[k + 1] = number of rows
// fold k has finished with (prog[0] = n: "no predecessor"); waitix[k] = index
// into prog of the fold that k follows.
BLSQ_HD void tall_fold(const Blk& B, double* A, double* b, const double* diag_h, int n,
                       double* rowbuf, volatile int* prog, int* waitix) {
    for (int k = B.tid; k <= n; k += B.nt) prog[k] = (k == 0) ? n : 0;
    if (B.tid == 0) {
        int last = 0;                      // prog index of the last active fold
        for (int k = 0; k < n; k++) {
            waitix[k] = last;
            if (diag_h[k] != 0.0) last = k + 1;
        }
    }
    B.sync();
    double* row = rowbuf + (size_t)B.warp * n;
    for (int k = B.warp; k < n; k += B.nwarps) {
        const double e = sqrt(diag_h[k]);
        if (e == 0.0) continue;
#if BLSQ_TALL_DEV
        const int wix = waitix[k];
        __syncwarp();
#endif
        for (int j = B.lane; j < n; j += B.lanes) row[j] = (j == k) ? e : 0.0;
        double bz = 0.0;
#if BLSQ_TALL_DEV
        __syncwarp();
#endif
        for (int i = k; i < n; i++) {
#if BLSQ_TALL_DEV
            while (prog[wix] <= i) { }          // the fold before us has left row i
            __threadfence_block();
#endif
            const double xv = row[i];
            if (xv != 0.0) {
                double* ai = A + (size_t)i * n;
                const double a = ai[i];
                const double bi = b[i];
                const double rr = sqrt(fma(a, a, xv * xv));
                const double c = a / rr, sn = xv / rr;
#if BLSQ_TALL_DEV
                __syncwarp();
#endif
                for (int j = i + B.lane; j < n; j += B.lanes) {
                    const double aj = ai[j], rj = row[j];
                    ai[j] = fma(c, aj, sn * rj);
                    row[j] = fma(-sn, aj, c * rj);
                }
                if (B.lane == 0) b[i] = fma(c, bi, sn * bz);
                bz = fma(-sn, bi, c * bz);
            }
#if BLSQ_TALL_DEV
            __syncwarp();
            __threadfence_block();
#endif
            if (B.lane == 0) prog[k + 1] = i + 1;
        }
    }
    B.sync();
}

// Factorisation of the hat-space augmented matrix (trf.py:264-274) from the
// triangle: A = R diag(d) with diag(sqrt(diag_h)) folded in, then Jacobi.
// Outputs (global state): S, SUF = S * (U^T f_aug), VT (row j = v_j).
BLSQ_HD void tall_hat_fold(const Blk& B, const double* R, const TallWork& W, double* A) {
    const int n = W.n;
    for (int e = B.tid; e < n * n; e += B.nt) {
        const int i = e / n, j = e % n;
        A[e] = (j >= i) ? R[e] * W.d[j] : 0.0;
    }
    for (int i = B.tid; i < n; i += B.nt) W.b[i] = W.qtf[i];
    B.sync();
    tall_fold(B, A, W.b, W.diag_h, n, W.rowbuf, W.prog, W.flags);
}

BLSQ_HD int tall_hat_finish(const Blk& B, const TallWork& W, double* A, double* S, double* SUF,
                            double* VT) {
    const int n = W.n;
    const int sweeps = tall_jacobi(B, A, W.b, n);
    for (int j = B.warp; j < n; j += B.nwarps) {
        double nn = 0.0;
        for (int i = B.lane; i < n; i += B.lanes) nn = fma(A[(size_t)j * n + i], A[(size_t)j * n + i], nn);
        nn = warp_sum(nn);
        const double sj = sqrt(nn);
        const double inv = (sj > 0.0) ? 1.0 / sj : 0.0;
        for (int i = B.lane; i < n; i += B.lanes) VT[(size_t)j * n + i] = A[(size_t)j * n + i] * inv;
        if (B.lane == 0) { S[j] = sj; SUF[j] = sj * W.b[j]; }
    }
    B.sync();
    return sweeps;
}

BLSQ_HD int tall_hat_svd(const Blk& B, const double* R, const TallWork& W, double* A,
                         double* S, double* SUF, double* VT) {
    tall_hat_fold(B, R, W, A);
    return tall_hat_finish(B, W, A, S, SUF, VT);
}

// Gauss-Newton shortcut of solve_lsq_trust_region (trust_region.py:108-117) on
// the folded triangle A and b (see gn_shortcut in blsq_core.cuh): full rank is
// certified by cond(A) <= |A|_F |A^-1|_F with a factor 4 to spare against
// EPS m, the step is p = -A^-1 b.  Every warp back-substitutes whole columns
// of the identity (only their squared norms are kept) and one extra right-hand
// side, b; row i of a solve is a lane-parallel dot product over the part of the
// solution that exists already.  Returns true when the certificate holds; then
// pgn = -A^-1 b and *pnorm = |pgn|.
// A is n x n row major; the triangle that is solved is T[i][k] = A[i][cm[k]]
// (cm == nullptr: identity) for 0 <= i <= k < nn, the right-hand side W.b[0..nn).
BLSQ_HD bool tall_tri_try(const Blk& B, const double* A, const TallWork& W, int nn,
                          const int* cm, double em, double* pgn, double* pnorm) {
    const int n = W.n;
    double fa = 0.0;
    bool bad = false;
    for (int e = B.tid; e < nn * nn; e += B.nt) {
        const int i = e / nn, k = e % nn;
        if (k < i) continue;
        const double a = A[(size_t)i * n + (cm ? cm[k] : k)];
        fa = fma(a, a, fa);
        if (i == k && !(a != 0.0)) bad = true;
    }
    fa = blk_sum(B, fa);
    if (blk_any(B, bad)) return false;
    double ft = 0.0;
    double* t = W.rowbuf + (size_t)B.warp * n;       // this warp's solution vector
    for (int c = B.warp; c <= nn; c += B.nwarps) {   // c == nn: right-hand side b
        const int top = (c < nn) ? c : nn - 1;       // e_c has no entries below row c
        double nrm = 0.0;
        for (int i = top; i >= 0; i--) {
            const double* ai = A + (size_t)i * n;
            double acc = 0.0;
            for (int k = i + 1 + B.lane; k <= top; k += B.lanes)
                acc = fma(ai[cm ? cm[k] : k], t[k], acc);
            acc = warp_sum(acc);
            const double rhs = (c < nn) ? (i == c ? 1.0 : 0.0) : W.b[i];
            const double ti = (rhs - acc) / ai[cm ? cm[i] : i];
            if (B.lane == 0) t[i] = ti;
            nrm = fma(ti, ti, nrm);
#if BLSQ_TALL_DEV
            __syncwarp();
#endif
        }
        if (c < nn) {
            ft += nrm;                                // same value on every lane
        } else {
            for (int i = B.lane; i < nn; i += B.lanes) pgn[i] = -t[i];
            if (B.lane == 0) *pnorm = sqrt(nrm);
        }
    }
    // ft was accumulated identically by all lanes of a warp: count it once
    ft = blk_sum(B, B.lane == 0 ? ft : 0.0);
    B.sync();
    return fa * ft * (em * em) < 0.0625;              // false for NaN / inf too
}

BLSQ_HD bool tall_gn_try(const Blk& B, const double* A, const TallWork& W, double m,
                         double* pgn, double* pnorm) {
    if (m < W.n) return false;
    return tall_tri_try(B, A, W, W.n, nullptr, EPS * m, pgn, pnorm);
}

// Householder QR of a column staircase of A (n x n, row major): position p is
// matrix column cols[p], whose entries reach down to row ext[p] >= p.  The
// reflectors are also applied to the trailing staircase columns, to the nx
// extra columns cols[nc .. nc + nx) (all rows) and to W.b.  One thread per
// trailing column (coalesced along the rows of A), ~3 block barriers per column.
BLSQ_HD void tall_staircase_qr(const Blk& B, double* A, const TallWork& W, const int* cols,
                               const int* ext, int nc, int nx) {
    const int n = W.n;
    double* v = W.w;
    for (int c = 0; c < nc; c++) {
        const int jc = cols[c];
        const int len = ext[c] - c + 1;               // rows c .. ext[c]
        if (len <= 1) continue;                       // already triangular here
        double nn = 0.0;
        for (int r = B.tid; r < len; r += B.nt) {
            const double a = A[(size_t)(c + r) * n + jc];
            v[r] = a;
            nn = fma(a, a, nn);
        }
        nn = blk_sum(B, nn);
        const double v0 = v[0];
        const double below = nn - v0 * v0;
        if (!(below > 0.0)) continue;                 // nothing under the diagonal
        const double alpha = (v0 > 0.0) ? -sqrt(nn) : sqrt(nn);
        const double u0 = v0 - alpha;                 // v <- v - alpha e_0
        const double vtv = below + u0 * u0;
        const double beta = 2.0 / vtv;
        B.sync();
        if (B.tid == 0) v[0] = u0;
        B.sync();
        // trailing columns, extra columns and the right-hand side (index nc + nx)
        for (int cc = c + 1 + B.tid; cc <= nc + nx; cc += B.nt) {
            double w = 0.0;
            if (cc < nc + nx) {
                const int j = cols[cc];
                for (int r = 0; r < len; r++) w = fma(v[r], A[(size_t)(c + r) * n + j], w);
                w *= beta;
                for (int r = 0; r < len; r++)
                    A[(size_t)(c + r) * n + j] = fma(-w, v[r], A[(size_t)(c + r) * n + j]);
            } else {
                for (int r = 0; r < len; r++) w = fma(v[r], W.b[c + r], w);
                w *= beta;
                for (int r = 0; r < len; r++) W.b[c + r] = fma(-w, v[r], W.b[c + r]);
            }
        }
        if (B.tid == 0) A[(size_t)c * n + jc] = alpha;
        B.sync();
    }
}

// Minimum-norm least squares for a NUMERICALLY RANK-DEFICIENT triangle (the
// C4 / C5 benchmark start: two pairs of identical columns, and dogbox never
// breaks the symmetry, so every iterate is rank deficient) without singular
// values.  In: A holds the QR factor T of the free columns (positions 0..nfree,
// matrix columns fl[.]) from tall_dogbox_ls, W.b the rotated right-hand side.
//   1. D = positions whose diagonal collapsed (|T_dd| <= 1e-10 max |T_ii|; at
//      most KMAX of them), I = the others.  The columns of D go to the end:
//      [T_I | T_D] is a staircase again (steps of at most |D| rows), one more
//      pass of short reflectors gives  [R11 R12; 0 R22].
//   2. Accept only when the rank decision is unambiguous against numpy's
//      rcond = eps max(m, n_free) (dogbox.py:197 -> gelsd): R11 certified well
//      above the cut (tall_tri_try with a 100x margin), |R22|_F at least 100x
//      below it, |W|_F^2 <= 100 for W = R11^-1 R12.  Otherwise: false, and the
//      caller takes the SVD.
//   3. z_I = -R11^-1 c1 is a particular solution, N = [-W; I] spans the null
//      space; the minimum-norm solution is z - N (N^T N)^-1 N^T z:
//        y = (I + W^T W)^-1 W^T z_I,   z_I <- z_I - W y,   z_D = y.
// For an exact duplicate W = e_i and the step is split evenly, as gelsd does.
BLSQ_HD bool tall_dogbox_ls_deficient(const Blk& B, const TallWork& W, double* A, int nfree,
                                      double m, double* newton, double* pnorm_slot) {
    constexpr int KMAX = 8;
    const int n = W.n;
    // scratch for the K x K system: x_edge | refl | c_h are consecutive vectors
    // (TallWork::carve) and unused at this point of the propose: 3 n doubles
    int kcap = KMAX;
    while (kcap * kcap > 3 * n) kcap--;
    double* scr = W.x_edge;
    const int* fl = W.hits;               // free columns by position
    int* cols = W.flags;                  // I positions first, then D
    int* ext = W.marks;
    double dmax = 0.0;
    for (int c = B.tid; c < nfree; c += B.nt) {
        const double a = fabs(A[(size_t)c * n + fl[c]]);
        if (a > dmax) dmax = a;
    }
    dmax = blk_max(B, dmax);
    if (!(dmax > 0.0)) return false;
    B.sync();
    if (B.tid == 0) {
        int ni = 0, nd = 0;
        for (int c = 0; c < nfree; c++)
            if (fabs(A[(size_t)c * n + fl[c]]) <= 1e-10 * dmax) nd++;
        int pi = 0, pd = nfree - nd;
        if (nd >= 1 && nd <= kcap) {
            for (int c = 0; c < nfree; c++) {
                const bool dep = fabs(A[(size_t)c * n + fl[c]]) <= 1e-10 * dmax;
                const int p = dep ? pd++ : pi++;
                cols[p] = fl[c];
                ext[p] = c;               // T is upper triangular: rows 0 .. c
            }
        }
        ni = nfree - nd;
        B.ired[0] = nd;
        B.ired[1] = ni;
    }
    B.sync();
    const int nd = B.ired[0], r = B.ired[1];
    B.sync();
    if (nd < 1 || nd > kcap || r < 1) return false;
    // the dependent columns hold stale reflector data under their diagonal
    for (int e = 0; e < nd; e++) {
        const int j = cols[r + e], d = ext[r + e];
        for (int i = d + 1 + B.tid; i < nfree; i += B.nt) A[(size_t)i * n + j] = 0.0;
    }
    B.sync();
    tall_staircase_qr(B, A, W, cols, ext, r, nd);
    // |R22|_F and |T|_F (an upper bound of s_max; s_max >= |T|_F / sqrt(nfree))
    double r22 = 0.0, tf2 = 0.0;
    for (int e = B.tid; e < nfree * nfree; e += B.nt) {
        const int i = e / nfree, p = e % nfree;
        if (p < r && i > p) continue;
        const double a = A[(size_t)i * n + cols[p]];
        tf2 = fma(a, a, tf2);
        if (p >= r && i >= r) r22 = fma(a, a, r22);
    }
    double z0 = 0.0;
    blk_sum3(B, r22, tf2, z0);
    const double mx = m > nfree ? m : (double)nfree;
    const double cut_lo = EPS * mx * sqrt(tf2 / nfree);
    if (!(sqrt(r22) < 0.01 * cut_lo)) return false;
    double* sol = W.suf;
    if (!tall_tri_try(B, A, W, r, cols, 100.0 * EPS * mx, sol, pnorm_slot)) return false;
    // W = R11^-1 R12, one warp per dependent column, in place
    double* t = W.rowbuf + (size_t)B.warp * n;
    for (int e = B.warp; e < nd; e += B.nwarps) {
        const int je = cols[r + e];
        for (int i = r - 1; i >= 0; i--) {
            const double* ai = A + (size_t)i * n;
            double acc = 0.0;
            for (int k = i + 1 + B.lane; k < r; k += B.lanes) acc = fma(ai[cols[k]], t[k], acc);
            acc = warp_sum(acc);
            const double ti = (ai[je] - acc) / ai[cols[i]];
            if (B.lane == 0) t[i] = ti;
#if BLSQ_TALL_DEV
            __syncwarp();
#endif
        }
        for (int i = B.lane; i < r; i += B.lanes) A[(size_t)i * n + je] = t[i];
    }
    B.sync();
    // M = I + W^T W, h = W^T z_I
    double* Mh = W.tv;                    // y (K doubles)
    double* hv = W.t3;                    // W^T z_I (K doubles)
    double wf = 0.0;
    for (int a1 = 0; a1 < nd; a1++) {
        for (int a2 = a1; a2 <= nd; a2++) {
            double acc = 0.0;
            for (int i = B.tid; i < r; i += B.nt) {
                const double wa = A[(size_t)i * n + cols[r + a1]];
                acc = fma(wa, a2 < nd ? A[(size_t)i * n + cols[r + a2]] : sol[i], acc);
            }
            acc = blk_sum(B, acc);
            if (B.tid == 0) {
                if (a2 < nd) scr[a1 * kcap + a2] = acc;
                else hv[a1] = acc;
            }
            if (a2 == a1) wf += acc;
        }
    }
    B.sync();
    if (!(wf <= 100.0)) return false;
    if (B.tid == 0) {
        // Cholesky of the K x K matrix I + W^T W, then y
        double Mm[KMAX][KMAX], y[KMAX];
        for (int a1 = 0; a1 < nd; a1++)
            for (int a2 = 0; a2 < nd; a2++)
                Mm[a1][a2] = (a1 <= a2 ? scr[a1 * kcap + a2] : scr[a2 * kcap + a1]) +
                             (a1 == a2 ? 1.0 : 0.0);
        for (int k = 0; k < nd; k++) {
            Mm[k][k] = sqrt(Mm[k][k]);
            for (int i = k + 1; i < nd; i++) Mm[i][k] /= Mm[k][k];
            for (int j = k + 1; j < nd; j++)
                for (int i = j; i < nd; i++) Mm[i][j] -= Mm[i][k] * Mm[j][k];
        }
        for (int i = 0; i < nd; i++) {
            double sacc = hv[i];
            for (int k = 0; k < i; k++) sacc -= Mm[i][k] * y[k];
            y[i] = sacc / Mm[i][i];
        }
        for (int i = nd - 1; i >= 0; i--) {
            double sacc = y[i];
            for (int k = i + 1; k < nd; k++) sacc -= Mm[k][i] * y[k];
            y[i] = sacc / Mm[i][i];
        }
        for (int i = 0; i < nd; i++) Mh[i] = y[i];
    }
    B.sync();
    for (int i = B.tid; i < n; i += B.nt) newton[i] = 0.0;
    B.sync();
    double nn = 0.0;
    for (int i = B.tid; i < r; i += B.nt) {
        double zi = sol[i];
        for (int e = 0; e < nd; e++) zi = fma(-A[(size_t)i * n + cols[r + e]], Mh[e], zi);
        newton[cols[i]] = zi;
        nn = fma(zi, zi, nn);
    }
    for (int e = B.tid; e < nd; e += B.nt) {
        newton[cols[r + e]] = Mh[e];
        nn = fma(Mh[e], Mh[e], nn);
    }
    nn = blk_sum(B, nn);
    if (B.tid == 0) *pnorm_slot = sqrt(nn);
    B.sync();
    return true;
}

// dogbox.py:197, newton_step = lstsq(J_free, -f)[0], without singular values:
// Householder QR of the free columns of R (a staircase: column fl[c] has
// entries down to row fl[c] only), the same reflectors on Q^T f, then the
// certified triangular solve of tall_tri_try against numpy's
// rcond = eps * max(m, n_free); a triangle whose diagonal has collapsed goes
// through tall_dogbox_ls_deficient.  Returns false when neither applies (the
// caller then takes the SVD route, which rebuilds A).
BLSQ_HD bool tall_dogbox_ls(const Blk& B, const double* R, const TallWork& W, double* A,
                            int nfree, double m, double* newton, double* pnorm_slot) {
    const int n = W.n;
    int* fl = W.hits;                                 // free columns, ascending
    for (int e = B.tid; e < n * n; e += B.nt) {
        const int i = e / n, j = e % n;
        A[e] = (j >= i && W.fr[j]) ? R[e] : 0.0;
    }
    for (int i = B.tid; i < n; i += B.nt) W.b[i] = W.qtf[i];
    if (B.tid == 0) {
        int c = 0;
        for (int j = 0; j < n; j++)
            if (W.fr[j]) fl[c++] = j;
    }
    B.sync();
    tall_staircase_qr(B, A, W, fl, fl, nfree, 0);
    const double mx = m > nfree ? m : (double)nfree;
    double* sol = W.suf;
    if (!tall_tri_try(B, A, W, nfree, fl, EPS * mx, sol, pnorm_slot))
        return tall_dogbox_ls_deficient(B, W, A, nfree, m, newton, pnorm_slot);
    for (int i = B.tid; i < n; i += B.nt) newton[i] = 0.0;
    B.sync();
    for (int c = B.tid; c < nfree; c += B.nt) newton[fl[c]] = sol[c];
    B.sync();
    return true;
}

// trust_region.py:47-53 over the block
BLSQ_HD void tall_phi(const Blk& B, int n, double alpha, const double* suf, const double* s,
                      double Delta, double& phi, double& dphi) {
    double nn = 0.0, dd = 0.0, z = 0.0;
    for (int i = B.tid; i < n; i += B.nt) {
        const double den = s[i] * s[i] + alpha;
        const double q = suf[i] / den;
        nn = fma(q, q, nn);
        dd += (suf[i] * suf[i]) / (den * den * den);
    }
    blk_sum3(B, nn, dd, z);
    const double pn = sqrt(nn);
    phi = pn - Delta;
    dphi = -dd / pn;
}

// trust_region.py:56-152; S, SUF, VT from tall_hat_svd.  p_h -> W.p_h.
BLSQ_HD void tall_solve_tr(const Blk& B, const TallWork& W, double m, const double* S,
                           const double* SUF, const double* VT, double Delta, double& alpha,
                           bool zero_col) {
    const int n = W.n;
    double smin = dinf(), smax = -dinf();
    for (int i = B.tid; i < n; i += B.nt) {
        smin = S[i] < smin ? S[i] : smin;
        smax = S[i] > smax ? S[i] : smax;
    }
    smin = blk_min(B, smin);
    smax = blk_max(B, smax);
    const bool full_rank = (m >= n) && (smin > EPS * m * smax);
    if (full_rank) {
        for (int j = B.tid; j < n; j += B.nt) W.w[j] = (SUF[j] / S[j]) / S[j];
        B.sync();
        cols_dot(B, VT, n, W.w, W.p_h, -1.0);
        B.sync();
        const double pn = sqrt(vdot(B, W.p_h, W.p_h, n));
        if (pn <= Delta) { alpha = 0.0; return; }
    }
    double hi = sqrt(vdot(B, SUF, SUF, n)) / Delta;
    double lo = 0.0;
    double phi, dphi;
    if (full_rank) {
        tall_phi(B, n, 0.0, SUF, S, Delta, phi, dphi);
        lo = -phi / dphi;
    }
    if (!full_rank && alpha == 0) {
        const double a1 = 0.001 * hi, a2 = sqrt(lo * hi);
        alpha = a1 > a2 ? a1 : a2;
    }
    for (int it = 0; it < 10; it++) {
        if (alpha < lo || alpha > hi) {
            const double a1 = 0.001 * hi, a2 = sqrt(lo * hi);
            alpha = a1 > a2 ? a1 : a2;
        }
        tall_phi(B, n, alpha, SUF, S, Delta, phi, dphi);
        if (fabs(phi) < 0.01 * Delta) break;
        if (phi < 0) hi = alpha;
        const double q = phi / dphi;
        const double cand = alpha - q;
        lo = lo > cand ? lo : cand;
        alpha -= (phi + Delta) * q / Delta;
    }
    // exactly zero singular value: see solve_lsq_trust_region in blsq_core.cuh
    if (smin == 0.0 && !zero_col && !(alpha > 0.0)) {
        const double a1 = 0.001 * hi, a2 = sqrt(lo * hi);
        alpha = a1 > a2 ? a1 : a2;
    }
    B.sync();
    for (int j = B.tid; j < n; j += B.nt) W.w[j] = SUF[j] / (S[j] * S[j] + alpha);
    B.sync();
    cols_dot(B, VT, n, W.w, W.p_h, -1.0);
    B.sync();
    if (phi > 0) {
        const double sc = Delta / sqrt(vdot(B, W.p_h, W.p_h, n));
        for (int i = B.tid; i < n; i += B.nt) W.p_h[i] *= sc;
        B.sync();
    }
}

// 'jac' scaling (trf.py:216-221,239-242 / dogbox.py:141-146,165-168): the
// column norms of J are those of its triangular factor.
BLSQ_HD void tall_scale(const Blk& B, const TallWork& W, const double* R, const double* scaling,
                        double* SCALE, int jac_scaling, int first, int new_lin) {
    const int n = W.n;
    for (int j = B.tid; j < n; j += B.nt) {
        double sc;
        if (jac_scaling && new_lin) {
            double nn = 0.0;
            for (int i = 0; i <= j; i++) nn = fma(R[(size_t)i * n + j], R[(size_t)i * n + j], nn);
            double cn = sqrt(nn);
            if (first) {
                if (cn == 0) cn = 1.0;
                sc = 1.0 / cn;
            } else {
                sc = np_min(SCALE[j], 1.0 / cn);
            }
            SCALE[j] = sc;
        } else if (!jac_scaling && first) {
            sc = 1.0 / scaling[j];
            SCALE[j] = sc;
        } else {
            sc = SCALE[j];
        }
        W.scale[j] = sc;
    }
    B.sync();
}

// ---- judge: ratio test of the trial in flight (trf.py:310-344 / dogbox.py:
//      222-251), radius update, termination tests, accept -----------------------
// obj_new = ||f(x_new)||^2.  Sets ist[TI_ACCEPT]; on accept copies the trial
// into X (dogbox: with the bound hits snapped, dogbox.py:253-261) and counts the
// Jacobian the driver is about to evaluate.
BLSQ_HD void tall_judge(const Blk& B, const TallParams& P, double obj_new, int first,
                        const double* lb, const double* ub, double* st, int* ist) {
    const int n = P.n;
    const TallLayout L(n);
    if (ist[TI_STATUS] != ST_RUNNING) return;
    double* X = st + L.X;
    double* XNEW = st + L.XNEW;
    int status = ST_RUNNING;
    bool adopt;
    int nfev;
    if (first) {
        nfev = 1;
        adopt = true;
    } else {
        nfev = ist[TI_NFEV] + 1;
        const double obj = st[TS_OBJ];
        const double actual = obj - obj_new;
        const double pred = st[TS_PRED];
        double x_norm_term;
        bool x_ok;
        double ratio;
        if (P.method == BLSQ_METHOD_TRF) {
            ratio = (pred > 0) ? (actual - st[TS_CORR]) / pred : 0.0;
            const double nsh = st[TS_NSTEPH];
            const double Delta = st[TS_DELTA];
            const double xn = sqrt(vdot(B, X, X, n));
            x_norm_term = SQRT_EPS > xn ? SQRT_EPS : xn;
            x_ok = st[TS_NSTEP] < P.xtol * x_norm_term;
            B.sync();
            if (B.tid == 0) {
                if (ratio < 0.25) {
                    const double Dn = 0.25 * nsh;
                    st[TS_ALPHA] *= Delta / Dn;
                    st[TS_DELTA] = Dn;
                } else if (ratio > 0.75 && nsh > 0.95 * Delta) {
                    st[TS_DELTA] = Delta * 2.0;
                    st[TS_ALPHA] *= 0.5;
                }
            }
        } else {
            ratio = (pred > 0) ? actual / pred : 0.0;
            const bool tr_hit = ist[TI_TRHIT] != 0;
            double Delta = st[TS_DELTA];
            if (ratio < 0.25) Delta = 0.25 * st[TS_NSTEP];
            else if (ratio > 0.75 && tr_hit) Delta *= 2.0;
            // dogbox.py:241-242: on the NEW Delta and the pre-step x
            double xs = 0.0;
            bool xnan = false;
            const double* SC = st + L.SCALE;
            for (int i = B.tid; i < n; i += B.nt) {
                const double q = fabs(X[i] / SC[i]);
                if (q != q) xnan = true;
                else if (q > xs) xs = q;
            }
            xs = blk_max(B, xs);
            if (blk_any(B, xnan)) xs = dnan();
            x_norm_term = SQRT_EPS > xs ? SQRT_EPS : xs;
            x_ok = Delta < P.xtol * x_norm_term;
            B.sync();
            if (B.tid == 0) st[TS_DELTA] = Delta;
        }
        const bool f_ok = fabs(actual) < P.ftol * obj && ratio > 0.25;
        if (f_ok && x_ok) status = 4;
        else if (f_ok) status = 2;
        else if (x_ok) status = 3;
        adopt = actual > 0;
    }
    B.sync();
    if (adopt) {
        if (P.method == BLSQ_METHOD_TRF || first) {
            for (int i = B.tid; i < n; i += B.nt) X[i] = XNEW[i];
        } else {
            int* onb = ist + L.ONB;
            const int* marks = ist + L.MARKS;
            const int* fr = ist + L.FREE;
            for (int i = B.tid; i < n; i += B.nt) {
                const int ob = fr[i] ? marks[i] : onb[i];
                onb[i] = ob;
                double xi = XNEW[i];
                if (ob == -1) xi = lb[i];
                if (ob == 1) xi = ub[i];
                X[i] = xi;
            }
        }
    }
    B.sync();
    if (B.tid == 0) {
        ist[TI_NFEV] = nfev;
        ist[TI_ACCEPT] = adopt ? 1 : 0;
        ist[TI_PENDING] = status;
        if (first) { ist[TI_NJEV] = 0; st[TS_ALPHA] = 0.0; }
        if (adopt) { st[TS_OBJ] = obj_new; ist[TI_NJEV] += 1; }
        // trf.py:238,354-358 / dogbox.py:164,269-272: budget exhausted ->
        // status 0 whatever the inner loop decided (Q-T6)
        if (nfev >= P.max_nfev) {
            if (first) st[TS_GNORM] = dnan();
            ist[TI_STATUS] = 0;
        }
    }
    B.sync();
}

// ---- TRF propose (trf.py:238-308) ---------------------------------------------------
// fac: factor record of the Jacobian at X (R, Q^T f, g).  new_lin: the record
// is new since the last call (accepted step) -> redo scaling and the SVD.
BLSQ_HD void tall_trf_propose(const Blk& B, const TallParams& P, const TallWork& W,
                              const double* fac, const double* x0, const double* scaling,
                              int first, int new_lin, double* A, double* st, int* ist) {
    const int n = P.n;
    const TallLayout L(n);
    const FacLayout FL(n);
    if (ist[TI_STATUS] != ST_RUNNING) return;
    const double* R = fac + FL.R;
    double* S = st + L.S;
    double* SUF = st + L.SUF;
    double* VT = st + L.VT;
    for (int i = B.tid; i < n; i += B.nt) {
        W.x[i] = st[L.X + i];
        W.g[i] = fac[FL.G + i];
        W.qtf[i] = fac[FL.QTF + i];
    }
    B.sync();
    tall_scale(B, W, R, scaling, st + L.SCALE, P.jac_scaling, first, new_lin);
    double gmax = 0.0;
    bool gnan = false;
    for (int i = B.tid; i < n; i += B.nt) {
        double v, jv;
        cl_scaling(W.x[i], W.g[i], W.lb[i], W.ub[i], v, jv);
        W.v[i] = v;
        W.d[i] = sqrt(v) * W.scale[i];
        W.g_h[i] = W.d[i] * W.g[i];
        W.diag_h[i] = W.g[i] * jv * (W.scale[i] * W.scale[i]);
        const double gv = fabs(W.g[i] * v);
        if (gv != gv) gnan = true;
        else if (gv > gmax) gmax = gv;
    }
    double g_norm = blk_max(B, gmax);
    if (blk_any(B, gnan)) g_norm = dnan();
    if (first) {
        // trf.py:223-226: Delta from the ORIGINAL x0 (Q-T1)
        double qq = 0.0;
        for (int i = B.tid; i < n; i += B.nt) {
            const double q = x0[i] / (W.scale[i] * sqrt(W.v[i]));
            qq = fma(q, q, qq);
        }
        const double D0 = sqrt(blk_sum(B, qq));
        if (B.tid == 0) st[TS_DELTA] = (D0 == 0) ? 1.0 : D0;
    }
    int status = ist[TI_PENDING];
    if (g_norm < P.gtol) status = 1;                  // trf.py:252-254 (overrides, Q-T5)
    B.sync();
    if (B.tid == 0) st[TS_GNORM] = g_norm;
    if (status != ST_RUNNING) {
        if (B.tid == 0) ist[TI_STATUS] = status;
        B.sync();
        return;
    }
    const double Delta = st[TS_DELTA];
    double alpha = st[TS_ALPHA];
    int gn = ist[TI_GN];
    B.sync();
    if (new_lin) {
        // Gauss-Newton shortcut first (no singular values needed, ~75 % of the
        // C4 solves); the SVD only when the step is not certified or too long
        tall_hat_fold(B, R, W, A);
        gn = tall_gn_try(B, A, W, P.m, SUF, st + TS_GNSTEP) ? 1 : 0;
        if (!gn) {
            const int sweeps = tall_hat_finish(B, W, A, S, SUF, VT);
            if (B.tid == 0) ist[TI_SWEEPS] = sweeps;
        }
        B.sync();
        if (B.tid == 0) ist[TI_GN] = gn;
    }
    bool took_gn = false;
    if (gn) {
        // the decision is left to the SVD route when |p| is within 1e-9 of
        // Delta, so that it is always taken with the reference's arithmetic
        if (st[TS_GNSTEP] <= Delta * (1.0 - 1e-9)) {
            for (int i = B.tid; i < n; i += B.nt) W.p_h[i] = SUF[i];
            alpha = 0.0;                              // trust_region.py:117
            took_gn = true;
        } else {
            if (!new_lin) tall_hat_fold(B, R, W, A);  // a rejected trial shrank Delta
            B.sync();
            const int sweeps = tall_hat_finish(B, W, A, S, SUF, VT);
            if (B.tid == 0) { ist[TI_SWEEPS] = sweeps; ist[TI_GN] = 0; }
        }
    }
    double theta = 1.0 - g_norm;
    if (theta < 0.995) theta = 0.995;
    B.sync();
    if (!took_gn) {
        // exactly zero column of [R diag(d); diag(sqrt(diag_h))]?  (see
        // solve_lsq_trust_region in blsq_core.cuh)
        bool zc = false;
        for (int j = B.tid; j < n; j += B.nt) {
            if (W.diag_h[j] != 0.0) continue;
            bool z = true;
            if (W.d[j] != 0.0)
                for (int i = 0; i <= j && z; i++) z = R[(size_t)i * n + j] == 0.0;
            zc = zc || z;
        }
        const bool zero_col = blk_any(B, zc);
        tall_solve_tr(B, W, P.m, S, SUF, VT, Delta, alpha, zero_col);
    }
    B.sync();
    for (int i = B.tid; i < n; i += B.nt) W.p[i] = W.d[i] * W.p_h[i];
    B.sync();
    const double to_bound = tall_step_to_bound(B, n, W.x, W.p, W.lb, W.ub, nullptr, W.tv, W.hits);
    double qbest;
    if (to_bound >= 1) {
        const double tb = theta * to_bound;
        const double fsc = tb < 1 ? tb : 1;
        for (int i = B.tid; i < n; i += B.nt) W.step_h[i] = W.p_h[i] * fsc;
        B.sync();
        qbest = tall_eval_quadratic(B, R, W, W.step_h);
    } else {
        // find_reflected_step, trf.py:105-156 (hits are those of to_bound)
        const double stride_p = to_bound;
        for (int i = B.tid; i < n; i += B.nt) {
            W.r_h[i] = W.hits[i] ? -W.p_h[i] : W.p_h[i];
            W.r[i] = W.d[i] * W.r_h[i];
            W.p[i] *= stride_p;
            W.p_h[i] *= stride_p;
            W.x_edge[i] = W.x[i] + W.p[i];
        }
        B.sync();
        // intersect_trust_region(p_h, r_h, Delta), trust_region.py:11-44
        double a3 = 0.0, b3 = 0.0, c3 = 0.0;
        for (int i = B.tid; i < n; i += B.nt) {
            a3 = fma(W.r_h[i], W.r_h[i], a3);
            b3 = fma(W.p_h[i], W.r_h[i], b3);
            c3 = fma(W.p_h[i], W.p_h[i], c3);
        }
        blk_sum3(B, a3, b3, c3);
        c3 -= Delta * Delta;
        if (a3 == 0 || c3 > 0) {
            if (B.tid == 0) ist[TI_STATUS] = (a3 == 0) ? ST_ERR_TR_ZERO : ST_ERR_TR_OUTSIDE;
            B.sync();
            return;
        }
        const double disc = sqrt(b3 * b3 - a3 * c3);
        const double qv = -(b3 + copysign(disc, b3));
        const double r1 = qv / a3, r2 = c3 / qv;
        const double to_tr = r1 < r2 ? r2 : r1;
        double tb2 = tall_step_to_bound(B, n, W.x_edge, W.r, W.lb, W.ub, nullptr, W.tv, nullptr);
        tb2 *= theta;
        const double hi = tb2 < to_tr ? tb2 : to_tr;
        const double lo = (hi > 0) ? (1 - theta) * stride_p / hi : -1.0;
        bool have_r = false;
        if (lo <= hi) {
            double a, b;
            tall_build_quadratic_1d(B, R, W, W.r_h, W.p_h, a, b);
            const double t = minimize_quadratic(a, b, lo, hi);
            for (int i = B.tid; i < n; i += B.nt) W.refl[i] = W.p_h[i] + W.r_h[i] * t;
            have_r = true;
        }
        B.sync();
        for (int i = B.tid; i < n; i += B.nt) {
            W.p_h[i] *= theta;
            if (!have_r) W.refl[i] = W.p_h[i];
            // find_gradient_step, trf.py:159-170
            W.ng[i] = -W.g_h[i];
            W.r[i] = W.ng[i] * W.d[i];          // r is free again: -g_h * d
        }
        B.sync();
        double tbg = tall_step_to_bound(B, n, W.x, W.r, W.lb, W.ub, nullptr, W.tv, nullptr);
        tbg *= theta;
        const double ttr = Delta / sqrt(vdot(B, W.g_h, W.g_h, n));
        const double hig = tbg < ttr ? tbg : ttr;
        double ag, bg;
        tall_build_quadratic_1d(B, R, W, W.ng, nullptr, ag, bg);
        const double tg = minimize_quadratic(ag, bg, 0.0, hig);
        for (int i = B.tid; i < n; i += B.nt) W.c_h[i] = -tg * W.g_h[i];
        B.sync();
        // trf.py:300-305: argmin, first minimum wins
        const double q0 = tall_eval_quadratic(B, R, W, W.p_h);
        const double q1 = tall_eval_quadratic(B, R, W, W.refl);
        const double q2 = tall_eval_quadratic(B, R, W, W.c_h);
        int k = 0;
        qbest = q0;
        if (q1 < qbest) { k = 1; qbest = q1; }
        if (q2 < qbest) { k = 2; qbest = q2; }
        for (int i = B.tid; i < n; i += B.nt)
            W.step_h[i] = (k == 0) ? W.p_h[i] : (k == 1 ? W.refl[i] : W.c_h[i]);
        B.sync();
    }
    double corr = 0.0, nsh = 0.0, ns = 0.0;
    for (int i = B.tid; i < n; i += B.nt) {
        const double sh = W.step_h[i];
        const double stp = W.d[i] * sh;
        corr = fma(sh * W.diag_h[i], sh, corr);
        nsh = fma(sh, sh, nsh);
        ns = fma(stp, stp, ns);
        st[L.XNEW + i] = strictly_feasible(W.x[i] + stp, W.lb[i], W.ub[i], 0.0);
    }
    blk_sum3(B, corr, nsh, ns);
    if (B.tid == 0) {
        st[TS_ALPHA] = alpha;
        st[TS_PRED] = -2 * qbest;
        st[TS_CORR] = corr;
        st[TS_NSTEPH] = sqrt(nsh);
        st[TS_NSTEP] = sqrt(ns);
    }
    B.sync();
}

// ---- dogbox propose (dogbox.py:164-223) ------------------------------------------------
BLSQ_HD bool tall_in_box(const Blk& B, int n, const double* s, const double* lo,
                         const double* hi, const int* fr) {
    bool bad = false;
    for (int i = B.tid; i < n; i += B.nt)
        if (fr[i] && !((s[i] >= lo[i]) && (s[i] <= hi[i]))) bad = true;
    return !blk_any(B, bad);
}

// dogbox.py:46-57,84-95: bound hits -> marks, trust-region hits -> tr_hit
BLSQ_HD bool tall_hit_bookkeeping(const Blk& B, int n, const int* hits, const int* flags,
                                  const int* fr, int* marks) {
    bool tr = false;
    for (int i = B.tid; i < n; i += B.nt) {
        int mk = 0;
        if (fr[i]) {
            if (hits[i] < 0 && (flags[i] & 1)) mk = -1;
            if (hits[i] > 0 && (flags[i] & 2)) mk = 1;
            if ((hits[i] < 0 && (flags[i] & 4)) || (hits[i] > 0 && (flags[i] & 8))) tr = true;
        }
        marks[i] = mk;
    }
    return blk_any(B, tr);
}

BLSQ_HD void tall_dogbox_propose(const Blk& B, const TallParams& P, const TallWork& W,
                                 const double* fac, const double* x0, const double* scaling,
                                 int first, int new_lin, double* A, double* st, int* ist) {
    const int n = P.n;
    const TallLayout L(n);
    const FacLayout FL(n);
    if (ist[TI_STATUS] != ST_RUNNING) return;
    const double* R = fac + FL.R;
    double* NEWTON = st + L.S;          // kept across rejected trials (dogbox.py:197-199)
    double* CAUCHY = st + L.SUF;
    int* onb = ist + L.ONB;
    double* newton = W.p_h; double* cauchy = W.p; double* gf = W.g_h; double* Jg = W.t2;
    double* tr = W.r_h; double* lo = W.r; double* hi = W.x_edge; double* cz = W.refl;
    double* diff = W.c_h; double* step = W.step; double* zero = W.ng;
    for (int i = B.tid; i < n; i += B.nt) {
        W.x[i] = st[L.X + i];
        W.g[i] = fac[FL.G + i];
        W.qtf[i] = fac[FL.QTF + i];
        zero[i] = 0.0;
        if (first) {
            // dogbox.py:152-154: exact equality (Q-D1)
            int ob = 0;
            if (x0[i] == W.lb[i]) ob = -1;
            if (x0[i] == W.ub[i]) ob = 1;
            onb[i] = ob;
        }
    }
    B.sync();
    tall_scale(B, W, R, scaling, st + L.SCALE, P.jac_scaling, first, new_lin);
    if (first) {
        double D0 = 0.0;
        bool dn = false;
        for (int i = B.tid; i < n; i += B.nt) {
            const double q = fabs(x0[i] / W.scale[i]);
            if (q != q) dn = true;
            else if (q > D0) D0 = q;
        }
        D0 = blk_max(B, D0);
        if (blk_any(B, dn)) D0 = dnan();
        if (B.tid == 0) st[TS_DELTA] = (D0 == 0) ? 1.0 : D0;
    }
    int nfree_l = 0;
    double gmax = 0.0;
    bool gnan = false;
    for (int i = B.tid; i < n; i += B.nt) {
        const int f = !(onb[i] * W.g[i] < 0);
        W.fr[i] = f;
        gf[i] = f ? W.g[i] : 0.0;
        if (f) {
            nfree_l++;
            const double ga = fabs(W.g[i]);
            if (ga != ga) gnan = true;
            else if (ga > gmax) gmax = ga;
        }
    }
    const int nfree = (int)(blk_sum(B, (double)nfree_l) + 0.5);
    double g_norm = blk_max(B, gmax);
    if (blk_any(B, gnan)) g_norm = dnan();
    int status = ist[TI_PENDING];
    if (nfree == 0 || g_norm < P.gtol) status = 1;     // dogbox.py:182-190
    B.sync();
    if (B.tid == 0) st[TS_GNORM] = g_norm;             // all active -> 0.0
    if (status != ST_RUNNING) {
        if (B.tid == 0) ist[TI_STATUS] = status;
        B.sync();
        return;
    }
    if (new_lin) {
        // newton_step = lstsq(J_free, -f) (dogbox.py:197): minimum norm through
        // the SVD of R[:, free], numpy rcond = eps * max(m, n_free)
        if (!(P.m >= nfree && tall_dogbox_ls(B, R, W, A, nfree, P.m, newton, st + TS_GNSTEP))) {
        for (int e = B.tid; e < n * n; e += B.nt) {
            const int i = e / n, j = e % n;
            A[e] = (j >= i && W.fr[j]) ? R[e] : 0.0;
        }
        for (int i = B.tid; i < n; i += B.nt) W.b[i] = W.qtf[i];
        B.sync();
        const int sweeps = tall_jacobi(B, A, W.b, n);
        if (B.tid == 0) ist[TI_SWEEPS] = sweeps;
        double smax2 = 0.0;
        for (int j = B.warp; j < n; j += B.nwarps) {
            double nn = 0.0;
            for (int i = B.lane; i < n; i += B.lanes) nn = fma(A[(size_t)j * n + i], A[(size_t)j * n + i], nn);
            nn = warp_sum(nn);
            if (B.lane == 0) W.s[j] = nn;          // s_j^2
        }
        B.sync();
        for (int j = B.tid; j < n; j += B.nt) smax2 = W.s[j] > smax2 ? W.s[j] : smax2;
        smax2 = blk_max(B, smax2);
        const double mx = P.m > nfree ? P.m : (double)nfree;
        const double cut = EPS * mx * sqrt(smax2);
        for (int j = B.tid; j < n; j += B.nt) W.w[j] = (sqrt(W.s[j]) > cut) ? W.b[j] / W.s[j] : 0.0;
        B.sync();
        cols_dot(B, A, n, W.w, newton, -1.0);
        B.sync();
        for (int i = B.tid; i < n; i += B.nt) if (!W.fr[i]) newton[i] = 0.0;
        }
        B.sync();
        // cauchy = -(g.g)/(Jg.Jg) g (dogbox.py:198-199), |J_free g| = |R g_free| (Q-D5 unguarded)
        tall_Rh_matvec(B, R, n, nullptr, gf, W.t1, Jg);
        double gg = 0.0, jj = 0.0, z = 0.0;
        for (int i = B.tid; i < n; i += B.nt) {
            gg = fma(gf[i], gf[i], gg);
            jj = fma(Jg[i], Jg[i], jj);
        }
        blk_sum3(B, gg, jj, z);
        const double cc = -gg / jj;
        for (int i = B.tid; i < n; i += B.nt) {
            cauchy[i] = cc * gf[i];
            NEWTON[i] = newton[i];
            CAUCHY[i] = cauchy[i];
        }
    } else {
        for (int i = B.tid; i < n; i += B.nt) { newton[i] = NEWTON[i]; cauchy[i] = CAUCHY[i]; }
    }
    B.sync();

    // ---- dogleg_step (dogbox.py:38-75) on the free coordinates ----
    const double Delta = st[TS_DELTA];
    for (int i = B.tid; i < n; i += B.nt) {
        tr[i] = Delta * W.scale[i];
        if (W.fr[i]) W.flags[i] = find_intersection(W.x[i], tr[i], W.lb[i], W.ub[i], lo[i], hi[i]);
        else { W.flags[i] = 0; lo[i] = 0; hi[i] = 0; }
        W.marks[i] = 0;
    }
    B.sync();
    bool tr_hit = false;
    if (tall_in_box(B, n, newton, lo, hi, W.fr)) {
        for (int i = B.tid; i < n; i += B.nt) step[i] = newton[i];     // Q-D2
    } else {
        for (int i = B.tid; i < n; i += B.nt) cz[i] = cauchy[i];
        B.sync();
        if (!tall_in_box(B, n, cz, lo, hi, W.fr)) {
            const double beta = tall_step_to_bound(B, n, zero, cz, lo, hi, W.fr, W.tv, W.hits);
            for (int i = B.tid; i < n; i += B.nt) cz[i] = beta * cz[i];
        }
        B.sync();
        for (int i = B.tid; i < n; i += B.nt) diff[i] = newton[i] - cz[i];
        B.sync();
        const double t = tall_step_to_bound(B, n, cz, diff, lo, hi, W.fr, W.tv, W.hits);
        tr_hit = tall_hit_bookkeeping(B, n, W.hits, W.flags, W.fr, W.marks);
        for (int i = B.tid; i < n; i += B.nt) step[i] = W.fr[i] ? cz[i] + t * diff[i] : 0.0;
    }
    B.sync();
    // predicted reduction (dogbox.py:208-209): |J s|^2 = |R s|^2, Js.f = s.g
    tall_Rh_matvec(B, R, n, nullptr, step, W.t1, W.t3);
    double JsJs = 0.0, Jsf = 0.0, z2 = 0.0;
    for (int i = B.tid; i < n; i += B.nt) {
        JsJs = fma(W.t3[i], W.t3[i], JsJs);
        Jsf = fma(step[i], gf[i], Jsf);
    }
    blk_sum3(B, JsJs, Jsf, z2);
    const double pred = -JsJs - 2 * Jsf;
    if (pred <= 0) {
        // constrained_cauchy_step (dogbox.py:78-97); the stale Js keeps pred <= 0 (Q-D3)
        if (tall_in_box(B, n, cauchy, lo, hi, W.fr)) {
            tr_hit = false;
            for (int i = B.tid; i < n; i += B.nt) { step[i] = cauchy[i]; W.marks[i] = 0; }
        } else {
            const double beta = tall_step_to_bound(B, n, zero, cauchy, lo, hi, W.fr, W.tv, W.hits);
            tr_hit = tall_hit_bookkeeping(B, n, W.hits, W.flags, W.fr, W.marks);
            for (int i = B.tid; i < n; i += B.nt) step[i] = W.fr[i] ? beta * cauchy[i] : 0.0;
        }
        B.sync();
    }
    double ns = 0.0;
    bool nsn = false;
    for (int i = B.tid; i < n; i += B.nt) {
        ist[L.MARKS + i] = W.marks[i];
        ist[L.FREE + i] = W.fr[i];
        st[L.XNEW + i] = W.x[i] + step[i];
        const double q = fabs(step[i] / W.scale[i]);
        if (q != q) nsn = true;
        else if (q > ns) ns = q;
    }
    ns = blk_max(B, ns);
    if (blk_any(B, nsn)) ns = dnan();
    if (B.tid == 0) {
        ist[TI_TRHIT] = tr_hit ? 1 : 0;
        st[TS_PRED] = pred;
        st[TS_NSTEP] = ns;
    }
    B.sync();
}

// trf.py:201 (x = make_strictly_feasible(x0, rstep=1e-10)) / dogbox.py:131 (x = x0)
BLSQ_HD void tall_init(const Blk& B, int method, int n, const double* x0, const double* lb,
                       const double* ub, double* st, int* ist) {
    const TallLayout L(n);
    for (int i = B.tid; i < n; i += B.nt) {
        double x = x0[i];
        if (method == BLSQ_METHOD_TRF) x = strictly_feasible(x, lb[i], ub[i], 1e-10);
        st[L.X + i] = x;
        st[L.XNEW + i] = x;
    }
    for (int i = B.tid; i < TS_NSCAL; i += B.nt) st[i] = 0.0;
    for (int i = B.tid; i < L.ISIZE; i += B.nt) ist[i] = (i == TI_STATUS || i == TI_PENDING) ? ST_RUNNING : 0;
    B.sync();
}

}  // namespace blsq_tall
