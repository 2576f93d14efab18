// TEST INFRASTRUCTURE ONLY -- never loaded by the bounded_lsq_b200 package.
//
// Compiles bounded_lsq_b200/csrc/blsq_core.cuh (the per-problem mathematics
// the CUDA kernels run, written as __host__ __device__ templates) with g++ and
// exposes the same C ABI as include/blsq.h on HOST pointers, so the branch
// logic of the kernels can be checked against the oracle and the golden
// vectors in the GPU-less authoring container.  The linearisation here is a
// serial modified Gram-Schmidt over all rows (the CUDA kernel does the same
// recurrence with the rows spread over a lane group).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/blsq.h"
#include "../../bounded_lsq_b200/csrc/blsq_core.cuh"

using namespace blsq;

#define DISPATCH_N(n, CALL)                               \
    switch (n) {                                          \
        case 1: { constexpr int N_ = 1; CALL; } break;    \
        case 2: { constexpr int N_ = 2; CALL; } break;    \
        case 3: { constexpr int N_ = 3; CALL; } break;    \
        case 4: { constexpr int N_ = 4; CALL; } break;    \
        case 5: { constexpr int N_ = 5; CALL; } break;    \
        case 6: { constexpr int N_ = 6; CALL; } break;    \
        case 7: { constexpr int N_ = 7; CALL; } break;    \
        case 8: { constexpr int N_ = 8; CALL; } break;    \
        default: return BLSQ_E_UNSUPPORTED;               \
    }

template <int N>
static void lin_host(int m, const double* F, const double* J,
                     const double* const* Fp, int64_t slot, const double* dx,
                     int mode, double* out) {
    typedef LinRec<N> L;
    constexpr int C = N + 1;
    std::vector<double> a((size_t)m * C);
    for (int r = 0; r < m; r++) {
        double f0 = F[slot * m + r];
        for (int j = 0; j < N; j++) {
            if (mode == 0) a[r * C + j] = J[(slot * m + r) * N + j];
            else if (mode == 1) a[r * C + j] = (Fp[j][slot * m + r] - f0) / dx[slot * N + j];
            else {
                const double f1 = Fp[2 * j][slot * m + r], f2 = Fp[2 * j + 1][slot * m + r];
                const bool one = dx[slot * 2 * N + N + j] != 0.0;
                const double df = one ? ((-3.0 * f0 + 4 * f1) - f2) : (f2 - f1);
                a[r * C + j] = df / dx[slot * 2 * N + j];
            }
        }
        a[r * C + N] = f0;
    }
    for (int i = 0; i < L::SIZE; i++) out[i] = 0.0;
    for (int j = 0; j < N; j++) {
        double g = 0.0;
        for (int r = 0; r < m; r++) g = fma(a[r * C + j], a[r * C + N], g);
        out[L::G + j] = g;
    }
    double obj = 0.0;
    for (int r = 0; r < m; r++) obj = fma(a[r * C + N], a[r * C + N], obj);
    out[L::OBJ] = obj;
    for (int k = 0; k < N; k++) {
        double dts[C];
        for (int j = k; j < C; j++) {
            double acc = 0.0;
            for (int r = 0; r < m; r++) acc = fma(a[r * C + k], a[r * C + j], acc);
            dts[j] = acc;
        }
        double dk = dts[k], rkk = sqrt(dk);
        out[L::R + tri_index<N>(k, k)] = rkk;
        for (int j = k + 1; j < C; j++) {
            double coef = dk > 0 ? dts[j] / dk : 0.0;
            double rv = dk > 0 ? dts[j] / rkk : 0.0;
            if (j < N) out[L::R + tri_index<N>(k, j)] = rv;
            else out[L::QTF + k] = rv;
            for (int r = 0; r < m; r++)
                a[r * C + j] = fma(-coef, a[r * C + k], a[r * C + j]);
        }
    }
}

extern "C" {

int blsq_version(void) { return BLSQ_VERSION; }
const char* blsq_error_string(int code) {
    return code == 0 ? "ok" : (code == -1 ? "invalid argument" : "unsupported");
}

int blsq_state_layout(int method, int n, int* out) {
    DISPATCH_N(n, {
        if (method == BLSQ_METHOD_TRF) {
            typedef TrfState<N_> S;
            out[0] = S::SIZE; out[1] = S::X; out[2] = S::XNEW; out[3] = S::SCALE;
            out[4] = S::OBJ; out[5] = S::DELTA; out[6] = S::GNORM; out[7] = S::G;
            out[8] = S::ALPHA;
        } else {
            typedef DogState<N_> S;
            out[0] = S::SIZE; out[1] = S::X; out[2] = S::XNEW; out[3] = S::SCALE;
            out[4] = S::OBJ; out[5] = S::DELTA; out[6] = S::GNORM; out[7] = S::G;
            out[8] = -1;
        }
    });
    return 0;
}

int blsq_lin_record_size(int n) {
    DISPATCH_N(n, { return LinRec<N_>::SIZE; });
    return BLSQ_E_UNSUPPORTED;
}

int blsq_step_size_to_bound(int64_t B, int n, const double* x, const double* d,
                            const double* lb, const double* ub, int bs,
                            double* step, int64_t* hits, void*) {
    for (int64_t b = 0; b < B; b++) {
        double tmin = dinf();
        bool has_nan = false;
        std::vector<double> t(n);
        for (int i = 0; i < n; i++) {
            double di = d[b * n + i];
            t[i] = dinf();
            if (di != 0)
                t[i] = np_max((lb[b * bs + i] - x[b * n + i]) / di,
                              (ub[b * bs + i] - x[b * n + i]) / di);
            if (t[i] != t[i]) has_nan = true;
            if (!(tmin < t[i])) tmin = t[i];
        }
        if (has_nan) tmin = dnan();
        step[b] = tmin;
        if (hits)
            for (int i = 0; i < n; i++)
                hits[b * n + i] = (t[i] == tmin) ? isign(d[b * n + i]) : 0;
    }
    return 0;
}

int blsq_find_active_constraints(int64_t B, int n, const double* x,
                                 const double* lb, const double* ub, int bs,
                                 double rtol, int64_t* mask, void*) {
    for (int64_t b = 0; b < B; b++)
        for (int i = 0; i < n; i++)
            mask[b * n + i] = active_constraint(x[b * n + i], lb[b * bs + i],
                                                ub[b * bs + i], rtol);
    return 0;
}

int blsq_make_strictly_feasible(int64_t B, int n, const double* x,
                                const double* lb, const double* ub, int bs,
                                double rstep, double* out, void*) {
    for (int64_t b = 0; b < B; b++)
        for (int i = 0; i < n; i++)
            out[b * n + i] = strictly_feasible(x[b * n + i], lb[b * bs + i],
                                               ub[b * bs + i], rstep);
    return 0;
}

int blsq_scaling_vector(int64_t B, int n, const double* x, const double* g,
                        const double* lb, const double* ub, int bs, double* v,
                        double* jv, void*) {
    for (int64_t b = 0; b < B; b++)
        for (int i = 0; i < n; i++)
            cl_scaling(x[b * n + i], g[b * n + i], lb[b * bs + i],
                       ub[b * bs + i], v[b * n + i], jv[b * n + i]);
    return 0;
}

int blsq_in_bounds(int64_t B, int n, const double* x, const double* lb,
                   const double* ub, int bs, uint8_t* ok, void*) {
    for (int64_t b = 0; b < B; b++) {
        bool good = true;
        for (int i = 0; i < n; i++)
            good = good && x[b * n + i] >= lb[b * bs + i] &&
                   x[b * n + i] <= ub[b * bs + i];
        ok[b] = good;
    }
    return 0;
}

int blsq_find_intersection(int64_t B, int n, const double* x, const double* tr,
                           const double* lb, const double* ub, int bs,
                           double* lo, double* hi, uint8_t* flags, void*) {
    for (int64_t b = 0; b < B; b++)
        for (int i = 0; i < n; i++)
            flags[b * n + i] = (uint8_t)find_intersection(
                x[b * n + i], tr[b * n + i], lb[b * bs + i], ub[b * bs + i],
                lo[b * n + i], hi[b * n + i]);
    return 0;
}

int blsq_fd2_points(int64_t A, const int32_t* idx, int n, const double* x,
                    const double* lb, const double* ub, int bs, double rel,
                    double* Xp, double* dx, void*) {
    for (int i = 0; i < n; i++)
        for (int64_t s = 0; s < A; s++) {
            int64_t pid = idx ? idx[s] : s;
            for (int k = 0; k < n; k++) {
                double xk = x[s * n + k];
                if (k == i) {
                    double h = fd2_step(xk, lb[pid * bs + k], ub[pid * bs + k], rel);
                    double xp = xk + h;
                    dx[s * n + i] = xp - xk;
                    xk = xp;
                }
                Xp[((int64_t)i * A + s) * n + k] = xk;
            }
        }
    return 0;
}

int blsq_fd3_points(int64_t A, const int32_t* idx, int n, const double* x,
                    const double* lb, const double* ub, int bs, double rel,
                    double* Xp, double* dxo, void*) {
    for (int i = 0; i < n; i++)
        for (int64_t s = 0; s < A; s++) {
            int64_t pid = idx ? idx[s] : s;
            for (int k = 0; k < n; k++) {
                double xk = x[s * n + k], x1 = xk, x2 = xk;
                if (k == i) {
                    bool one;
                    double h = fd3_step(xk, lb[pid * bs + k], ub[pid * bs + k], rel, one);
                    if (one) { x1 = xk + h; x2 = xk + 2 * h; dxo[s * 2 * n + i] = x2 - xk; }
                    else { x1 = xk - h; x2 = xk + h; dxo[s * 2 * n + i] = x2 - x1; }
                    dxo[s * 2 * n + n + i] = one ? 1.0 : 0.0;
                }
                Xp[((int64_t)(2 * i) * A + s) * n + k] = x1;
                Xp[((int64_t)(2 * i + 1) * A + s) * n + k] = x2;
            }
        }
    return 0;
}

int blsq_init_batched(int method, int64_t B, int n, const double* x0,
                      const double* lb, const double* ub, int bs,
                      double* state, int32_t* istate, double* Xnew, void*) {
    int lay[9];
    int rc = blsq_state_layout(method, n, lay);
    if (rc) return rc;
    int SS = lay[0];
    for (int64_t b = 0; b < B; b++) {
        for (int i = 0; i < n; i++) {
            double x = x0[b * n + i];
            if (method == BLSQ_METHOD_TRF)
                x = strictly_feasible(x, lb[b * bs + i], ub[b * bs + i], 1e-10);
            state[b * SS + lay[2] + i] = x;
            state[b * SS + i] = x;
            Xnew[b * n + i] = x;
        }
        int32_t* ip = istate + b * IS_SIZE;
        for (int k = 0; k < IS_SIZE; k++) ip[k] = 0;
        ip[IS_STATUS] = ST_RUNNING;
    }
    return 0;
}

int blsq_linearise_batched(int64_t A, const int32_t* idx, int m, int n,
                           const double* F, const double* J,
                           const double* const* Fp, const double* dx, int mode,
                           const int32_t* istate, double* lin, void*) {
    DISPATCH_N(n, {
        for (int64_t s = 0; s < A; s++) {
            int64_t pid = idx ? idx[s] : s;
            if (istate[pid * IS_SIZE + IS_STATUS] != ST_RUNNING) continue;
            lin_host<N_>(m, F, J, Fp, s, dx, mode, lin + s * LinRec<N_>::SIZE);
        }
    });
    return 0;
}

int blsq_round_batched(int method, int64_t A, const int32_t* idx, int m, int n,
                       const double* lin, const double* x0, const double* lb,
                       const double* ub, int bs, const double* scaling,
                       double ftol, double xtol, double gtol, int max_nfev,
                       int first, double* state, int32_t* istate, double* Xnew,
                       double* Xjac, int32_t* /*work*/, int32_t* count, void*) {
    SolveParams P;
    P.ftol = ftol; P.xtol = xtol; P.gtol = gtol;
    P.max_nfev = max_nfev; P.m = m; P.jac_scaling = scaling ? 0 : 1;
    DISPATCH_N(n, {
        double sc[N_];
        for (int i = 0; i < N_; i++) sc[i] = scaling ? scaling[i] : 1.0;
        for (int64_t s = 0; s < A; s++) {
            int64_t pid = idx ? idx[s] : s;
            int32_t* ist = istate + pid * IS_SIZE;
            if (ist[IS_STATUS] != ST_RUNNING) continue;
            const double* ln = lin + s * LinRec<N_>::SIZE;
            bool go;
            double* st;
            int xnew_off = N_;
            if (method == BLSQ_METHOD_TRF) {
                xnew_off = TrfState<N_>::XNEW;
                st = state + pid * TrfState<N_>::SIZE;
                go = trf_round<N_>(st, ist, ln, x0 + pid * N_, lb + pid * bs,
                                   ub + pid * bs, sc, P, first);
            } else {
                st = state + pid * DogState<N_>::SIZE;
                go = dogbox_round<N_>(st, ist, ln, x0 + pid * N_, lb + pid * bs,
                                      ub + pid * bs, sc, P, first);
            }
            if (go) {
                for (int i = 0; i < N_; i++) {
                    double xn = st[xnew_off + i];
                    Xnew[s * N_ + i] = xn;
                    if (Xjac) {
                        double xj = xn;
                        if (method == BLSQ_METHOD_DOGBOX) {
                            int ob = ((ist[IS_FREE] >> i) & 1)
                                         ? get2(ist[IS_MARKS], i)
                                         : get2(ist[IS_ONB], i);
                            if (ob == -1) xj = lb[pid * bs + i];
                            if (ob == 1) xj = ub[pid * bs + i];
                        }
                        Xjac[s * N_ + i] = xj;
                    }
                }
            }
        }
    });
    if (count) {
        int c = 0;
        for (int64_t s = 0; s < A; s++)
            c += istate[(idx ? idx[s] : s) * IS_SIZE + IS_STATUS] == ST_RUNNING;
        count[2] = c;
    }
    return 0;
}

int blsq_covariance(int64_t B, int n, const double* rec, int64_t stride, int r_off,
                    int packed, double* cov, void*) {
    for (int64_t b = 0; b < B; b++) {
        const double* R = rec + b * stride + r_off;
        double* C = cov + b * (int64_t)n * n;
        auto r = [&](int i, int j) {
            return packed ? R[i * n - (i * (i - 1)) / 2 + (j - i)] : R[(int64_t)i * n + j];
        };
        // pivots at rounding level of the largest one: J^T J has no inverse
        double dmax = 0.0;
        for (int i = 0; i < n; i++) dmax = fabs(r(i, i)) > dmax ? fabs(r(i, i)) : dmax;
        bool singular = false;
        for (int i = 0; i < n; i++)
            singular = singular || !(fabs(r(i, i)) > 16.0 * n * 2.220446049250313e-16 * dmax);
        if (singular) {
            for (int e = 0; e < n * n; e++) C[e] = std::nan("");
            continue;
        }
        std::vector<double> T((size_t)n * n, 0.0);
        for (int j = 0; j < n; j++)
            for (int i = j; i >= 0; i--) {
                if (i == j) { T[(size_t)i * n + j] = 1.0 / r(i, i); continue; }
                double acc = 0.0;
                for (int k = i + 1; k <= j; k++) acc = fma(r(i, k), T[(size_t)k * n + j], acc);
                T[(size_t)i * n + j] = -acc / r(i, i);
            }
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                double acc = 0.0;
                for (int k = (i > j ? i : j); k < n; k++)
                    acc = fma(T[(size_t)i * n + k], T[(size_t)j * n + k], acc);
                C[(int64_t)i * n + j] = acc;
            }
    }
    return 0;
}

int blsq_dogbox_on_bound(int64_t B, int n, const int32_t* istate,
                         int64_t* mask, void*) {
    for (int64_t b = 0; b < B; b++)
        for (int i = 0; i < n; i++)
            mask[b * n + i] = get2(istate[b * IS_SIZE + IS_ONB], i);
    return 0;
}

int blsq_count_running(int64_t B, const int32_t* idx, const int32_t* istate,
                       int32_t* count, void*) {
    int c = 0;
    for (int64_t b = 0; b < B; b++)
        c += istate[(idx ? idx[b] : b) * IS_SIZE + IS_STATUS] == ST_RUNNING;
    *count = c;
    return 0;
}

int64_t blsq_compact_work_size(int64_t A) { return A < 0 ? BLSQ_E_BADARG : A / 1024 + 2; }

int blsq_compact_batched(int64_t A, const int32_t* idx, const int32_t* istate, int n,
                         const double* Xnew, const double* Xjac, int32_t* idx_out,
                         int64_t* idx64_out, double* Xnew_out, double* Xjac_out,
                         int32_t* work, void*) {
    int64_t pos = 0;
    for (int64_t s = 0; s < A; s++) {
        const int64_t pid = idx ? idx[s] : s;
        if (istate[pid * IS_SIZE + IS_STATUS] != ST_RUNNING) continue;
        idx_out[pos] = (int32_t)pid;
        if (idx64_out) idx64_out[pos] = pid;
        for (int i = 0; i < n; i++) {
            Xnew_out[pos * n + i] = Xnew[s * n + i];
            if (Xjac) Xjac_out[pos * n + i] = Xjac[s * n + i];
        }
        pos++;
    }
    if (work) work[blsq_compact_work_size(A) - 1] = (int32_t)pos;
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Tall mode on the host (TEST ONLY): the same C ABI, CholeskyQR2 in plain loops
// and blsq_tall_core.cuh run by a single "thread".
// ---------------------------------------------------------------------------
#include "../../bounded_lsq_b200/csrc/blsq_tall_core.cuh"

namespace {
using namespace blsq_tall;

int host_chol_upper(double* M, const double* diag0, int n) {
    const double tol = 8.0 * n * 2.220446049250313e-16;
    for (int k = 0; k < n; k++) {
        double piv = M[k * n + k];
        if (!(piv > tol * diag0[k])) return k + 1;
        double r = sqrt(piv);
        for (int j = k; j < n; j++) M[k * n + j] = (j == k) ? r : M[k * n + j] / r;
        for (int i = k + 1; i < n; i++)
            for (int j = i; j < n; j++)
                M[i * n + j] = fma(-M[k * n + i], M[k * n + j], M[i * n + j]);
    }
    return 0;
}

void host_inv_upper(const double* R, double* X, int n) {
    for (int j = 0; j < n; j++) {
        for (int i = j + 1; i < n; i++) X[i * n + j] = 0.0;
        X[j * n + j] = 1.0 / R[j * n + j];
        for (int i = j - 1; i >= 0; i--) {
            double s = 0.0;
            for (int k = i + 1; k <= j; k++) s = fma(R[i * n + k], X[k * n + j], s);
            X[i * n + j] = -s / R[i * n + i];
        }
    }
}
}  // namespace

extern "C" {

int64_t blsq_tall_gram_work_size(int n) { return (n < 2 || n > 256) ? -2 : 8; }
int64_t blsq_tall_fac_size(int n) { return (n < 2 || n > 256) ? -2 : FacLayout(n).SIZE; }

int64_t blsq_tall_record_size(int n) {
    return (n < 2 || n > 256) ? -2 : (((int64_t)n * n + 2 * n + 1 + 1) & ~(int64_t)1);
}
int blsq_tall_sample_stride(int64_t m, int n) {
    if (n < 2 || n > 256 || m < 0) return 1;
    int64_t s = m / (64 * (int64_t)n * n);
    return (int)(s < 1 ? 1 : (s > 8 ? 8 : s));
}

int blsq_tall_gram(int pass, int64_t m, int n, const double* J, const double* f,
                   const double* rinvp, int sstride, double*, double* out, void*) {
    if (n < 2 || n > 256) return BLSQ_E_UNSUPPORTED;
    std::vector<double> X;
    if (pass == 2) {
        // unpack the fragment-ordered R1^-1
        X.assign((size_t)n * n, 0.0);
        const int nb = nb_for(n);
        for (int kb = 0; kb < nb; kb++)
            for (int jb = kb; jb < nb; jb++) {
                const int q = tri_block(nb, kb, jb);
                for (int half = 0; half < 2; half++)
                    for (int lane = 0; lane < 32; lane++) {
                        int r = 8 * kb + 4 * half + (lane & 3), c = 8 * jb + (lane >> 2);
                        if (r < n && c < n) X[(size_t)r * n + c] = rinvp[(q * 2 + half) * 32 + lane];
                    }
            }
    }
    for (int e = 0; e < n * n + 2 * n + 1; e++) out[e] = 0.0;
    std::vector<double> y(n);
    // pass 1 may sample one 64-row tile out of sstride (same map as the kernel)
    const int T = 64;
    if (pass == 2 || sstride < 1) sstride = 1;
    const int64_t ntiles = (m + T - 1) / T;
    const int64_t njobs = (pass == 1 && sstride > 1) ? ntiles / sstride : ntiles;
    for (int64_t job = 0; job < njobs; job++) {
      int64_t t = job;
      if (pass == 1 && sstride > 1) {
          uint32_t h = (uint32_t)job * 2654435761u;
          h ^= h >> 15;
          t = job * sstride + (int64_t)(h % (uint32_t)sstride);
      }
      for (int64_t r = t * T; r < (t + 1) * T && r < m; r++) {
        const double* row = J + r * n;
        if (pass == 2) {
            for (int c = 0; c < n; c++) {
                double s = 0.0;
                for (int k = 0; k <= c; k++) s = fma(row[k], X[(size_t)k * n + c], s);
                y[c] = s;
            }
        } else {
            for (int c = 0; c < n; c++) y[c] = row[c];
        }
        for (int i = 0; i < n; i++) {
            for (int j = i; j < n; j++) out[i * n + j] = fma(y[i], y[j], out[i * n + j]);
            if (pass == 2) {
                out[n * n + i] = fma(y[i], f[r], out[n * n + i]);
                out[n * n + n + 1 + i] = fma(row[i], f[r], out[n * n + n + 1 + i]);
            }
        }
        if (pass == 2) out[n * n + n] = fma(f[r], f[r], out[n * n + n]);
      }
    }
    return 0;
}

int blsq_tall_factor(int pass, int n, int nranks, int64_t gstride, const double* grams,
                     double* fac, void*) {
    const FacLayout FL(n);
    const int n2 = n * n;
    std::vector<double> M(n2), diag0(n);
    auto pack = [&](const std::vector<double>& X) {
        const int nb = FL.nb;
        for (int kb = 0; kb < nb; kb++)
            for (int jb = kb; jb < nb; jb++) {
                const int q = tri_block(nb, kb, jb);
                for (int half = 0; half < 2; half++)
                    for (int lane = 0; lane < 32; lane++) {
                        int r = 8 * kb + 4 * half + (lane & 3), c = 8 * jb + (lane >> 2);
                        fac[FL.RINVP + (q * 2 + half) * 32 + lane] =
                            (r < n && c < n && c >= r) ? X[(size_t)r * n + c] : 0.0;
                    }
            }
    };
    if (pass == 3) {
        for (int e = 0; e < n2; e++) { fac[FL.R1 + e] = fac[FL.R + e]; M[e] = fac[FL.R + e]; }
        std::vector<double> X(n2, 0.0);
        host_inv_upper(M.data(), X.data(), n);
        pack(X);
        fac[FL.REFINE] = 0.0;
        fac[FL.INFO] = 0.0;
        return 0;
    }
    double shift = 0.0, dmax = 0.0;
    int bad = 0;
    for (int attempt = 0; attempt < 12; attempt++) {
        const int nvec = (pass == 2 && attempt == 0) ? 2 * n + 1 : 0;
        for (int e = 0; e < n2 + nvec; e++) {
            double s = 0.0;
            for (int r = 0; r < nranks; r++) s += grams[(size_t)r * gstride + e];
            if (e < n2) {
                if (e / n == e % n) { if (attempt == 0) diag0[e / n] = s; s += shift; }
                M[e] = s;
            } else if (e < n2 + n) fac[FL.QTF + e - n2] = s;
            else if (e == n2 + n) fac[FL.OBJ] = s;
            else fac[FL.G + e - n2 - n - 1] = s;
        }
        if (attempt == 0) for (int i = 0; i < n; i++) dmax = diag0[i] > dmax ? diag0[i] : dmax;
        if (attempt == 0 && pass == 2) {
            bool far = false;
            const double tiny = 1e-8 * dmax;
            double dmin = dmax;
            for (int i = 0; i < n; i++) {
                if (!(diag0[i] > tiny)) continue;
                double rs = 0.0;
                for (int j = 0; j < n; j++) {
                    if (j == i || !(diag0[j] > tiny)) continue;
                    double gij = (j > i) ? M[i * n + j] : M[j * n + i];
                    rs += fabs(gij) / sqrt(diag0[i] * diag0[j]);
                }
                if (!(rs <= 0.5)) far = true;
                dmin = diag0[i] < dmin ? diag0[i] : dmin;
            }
            (void)dmin;
            fac[FL.REFINE] = far ? 1.0 : 0.0;
        }
        bad = host_chol_upper(M.data(), diag0.data(), n);
        if (!bad) break;
        if (!(dmax > 0.0)) break;
        shift = (shift == 0.0) ? 16.0 * n * 2.220446049250313e-16 * dmax : shift * 10.0;
    }
    fac[FL.SHIFT + pass - 1] = shift;
    if (pass == 2 && shift != 0.0) fac[FL.REFINE] = 0.0;
    if (bad) { fac[FL.INFO] = 1000.0 * pass + bad; return 0; }
    if (pass == 1) {
        for (int e = 0; e < n2; e++) fac[FL.R1 + e] = (e % n >= e / n) ? M[e] : 0.0;
        std::vector<double> X(n2, 0.0);
        host_inv_upper(M.data(), X.data(), n);
        pack(X);
        fac[FL.INFO] = 0.0;
        return 0;
    }
    const double* R1 = fac + FL.R1;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            double s = 0.0;
            if (j >= i) for (int k = i; k <= j; k++) s = fma(M[i * n + k], R1[k * n + j], s);
            fac[FL.R + i * n + j] = s;
        }
    double* qtf = fac + FL.QTF;
    for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int k = 0; k < i; k++) s = fma(M[k * n + i], qtf[k], s);
        qtf[i] = (qtf[i] - s) / M[i * n + i];
    }
    fac[FL.INFO] = 0.0;
    return 0;
}

int blsq_tall_sumsq(int64_t m, const double* f, double*, double* out, void*) {
    double s = 0.0;
    for (int64_t i = 0; i < m; i++) s = fma(f[i], f[i], s);
    out[0] = s;
    return 0;
}

int blsq_tall_layout(int n, int64_t* out) {
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
                    double* state, int32_t* istate, double*, void*) {
    if (n < 2 || n > BLSQ_MAX_TALL_N) return BLSQ_E_UNSUPPORTED;
    TallParams P;
    P.ftol = ftol; P.xtol = xtol; P.gtol = gtol;
    P.max_nfev = max_nfev; P.m = (double)m_total;
    P.jac_scaling = scaling ? 0 : 1; P.n = n; P.method = method;
    std::vector<double> red(128);
    std::vector<int> ints(5 * n + 1 + 64);
    Blk B;
    B.tid = 0; B.nt = 1; B.lane = 0; B.warp = 0; B.nwarps = 1; B.lanes = 1;
    B.red = red.data(); B.ired = ints.data() + 5 * n + 1;
    if (phase == 0) { tall_init(B, method, n, x0, lb, ub, state, istate); return 0; }
    if (phase == 1) {
        double obj_new = 0.0;
        for (int r = 0; r < nranks; r++) obj_new += ssq_parts[r];
        tall_judge(B, P, obj_new, first, lb, ub, state, istate);
        return 0;
    }
    std::vector<double> vec(TallWork::doubles(n)), rowbuf(n), A((size_t)n * n);
    TallWork W;
    W.carve(vec.data(), n);
    W.rowbuf = rowbuf.data();
    W.hits = ints.data(); W.flags = ints.data() + n; W.fr = ints.data() + 2 * n;
    W.marks = ints.data() + 3 * n; W.prog = ints.data() + 4 * n;
    W.A = A.data();
    for (int i = 0; i < n; i++) { W.lb[i] = lb[i]; W.ub[i] = ub[i]; }
    if (method == BLSQ_METHOD_TRF)
        tall_trf_propose(B, P, W, fac, x0, scaling, first, new_lin, A.data(), state, istate);
    else
        tall_dogbox_propose(B, P, W, fac, x0, scaling, first, new_lin, A.data(), state, istate);
    return 0;
}

}  // extern "C"
