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
            else a[r * C + j] = (Fp[j][slot * m + r] - f0) / dx[slot * N + j];
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
            state[b * SS + n + i] = x;
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
                       double* Xjac, void*) {
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
            if (method == BLSQ_METHOD_TRF) {
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
                    double xn = st[N_ + i];
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

}  // extern "C"
