// Batched mode kernels + C ABI (include/blsq.h) for sm_100a.
//
// Two kernels per round of the lock-step solve:
//
//   lin_kernel    HBM-bound.  A group of 8/16/32 lanes owns one problem; the
//                 lanes stream the rows of [J | f] with coalesced vector loads
//                 straight into registers and run a modified Gram-Schmidt QR
//                 with f as the last column, dot products reduced by warp
//                 shuffles inside the group.  Output per problem: packed R,
//                 Q^T f, g = J^T f, f.f  (a few hundred bytes).
//   round_kernel  FP64-bound.  One thread owns one problem and runs the whole
//                 n x n tail in registers (blsq_core.cuh): ratio test, accept,
//                 Coleman-Li scaling, SVD, LM parameter, candidates, select.
//
// Reference lines each stage stands in for are cited in blsq_core.cuh and
// include/blsq.h.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/blsq.h"
#include "blsq_core.cuh"
#include "blsq_lin.cuh"

using namespace blsq;
using namespace blsq_lin;

#define BLSQ_LAUNCH_CHECK()                                  \
    do {                                                     \
        cudaError_t e_ = cudaGetLastError();                 \
        if (e_ != cudaSuccess) return (int)e_;               \
    } while (0)

namespace {

#ifndef BLSQ_ROUND_THREADS
#define BLSQ_ROUND_THREADS 128
#endif
#ifndef BLSQ_ROUND_MINB
#define BLSQ_ROUND_MINB 4
#endif

// MODE 0: the whole round in one kernel.  TRF with a worklist (work != null):
// MODE 1 runs the round with the Gauss-Newton shortcut and appends the slots
// that need the SVD route to work[1..] (work[0] = their number); MODE 2 then
// finishes exactly those slots with dense warps (blsq_core.cuh trf_round_impl).
// 1: L2 prefetch of the TRF records at the top of the round (C2 full-batch
// round 0.248 -> 0.215 ms per 1e6 problems, profiles/r2_kbench_prefetch.jsonl).
// Measured against it and dropped: the same prefetch into L1 (no different), a
// one-wave grid striding over the slots and prefetching one slot ahead
// (0.313 ms), and staging both records through shared memory with 16-byte
// cp.async copies (0.266 ms).
#ifndef BLSQ_ROUND_PREFETCH
#define BLSQ_ROUND_PREFETCH 1
#endif
#if BLSQ_ROUND_PREFETCH == 2
#define BLSQ_PREFETCH(p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p))
#else
#define BLSQ_PREFETCH(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#endif

template <int N, int METHOD, int MODE>
__device__ __forceinline__ bool
round_slot(int64_t slot, int64_t A, const int32_t* __restrict__ idx,
           const double* __restrict__ lin, const double* __restrict__ x0,
           const double* __restrict__ lb, const double* __restrict__ ub,
           int bstride, const double* __restrict__ scaling, const SolveParams& P,
           int first, double* __restrict__ state,
           int32_t* __restrict__ istate, double* __restrict__ Xnew,
           double* __restrict__ Xjac, int32_t* __restrict__ work) {
    typedef LinRec<N> L;
    constexpr int SS = (METHOD == BLSQ_METHOD_TRF) ? TrfState<N>::SIZE
                                                   : DogState<N>::SIZE;
    if (slot >= A) return false;
    const int64_t pid = idx ? idx[slot] : slot;
    int32_t* ip = istate + pid * IS_SIZE;
    double sc[N];
#pragma unroll
    for (int i = 0; i < N; i++) sc[i] = scaling ? scaling[i] : 1.0;
    if (METHOD == BLSQ_METHOD_TRF) {
#if BLSQ_ROUND_PREFETCH
        // every line of the two records is requested NOW: the round reads them
        // block by block behind data-dependent branches, and at 16 warps per
        // SM one DRAM round trip per block is what the kernel waits for
        {
            const char* s0 = reinterpret_cast<const char*>(state + pid * (int64_t)SS);
            const char* l0 = reinterpret_cast<const char*>(lin + slot * (int64_t)L::SIZE);
#pragma unroll
            for (int o = 0; o < SS * 8; o += 128) BLSQ_PREFETCH(s0 + o);
#pragma unroll
            for (int o = 0; o < L::SIZE * 8; o += 128) BLSQ_PREFETCH(l0 + o);
            if (bstride) {
                BLSQ_PREFETCH(lb + pid * bstride);
                BLSQ_PREFETCH(ub + pid * bstride);
            }
        }
#endif
        // the records stay in memory: trf_round_impl fetches the blocks it
        // needs when it needs them (256-bit loads) and stores what changed
        int ist[4];
        {
            int4 t0 = *reinterpret_cast<const int4*>(ip);
            ist[0] = t0.x; ist[1] = t0.y; ist[2] = t0.z; ist[3] = t0.w;
        }
        if (ist[IS_STATUS] != ST_RUNNING) return false;
        const int rc = trf_round_impl<N, MODE>(
            state + pid * (int64_t)SS, ist, lin + slot * (int64_t)L::SIZE, x0 + pid * N,
            lb + pid * bstride, ub + pid * bstride, sc, P, first, Xnew + slot * N);
        *reinterpret_cast<int4*>(ip) = make_int4(ist[0], ist[1], ist[2], ist[3]);
        if (MODE == 1 && rc == TRF_DEFER)
            work[1 + atomicAdd(work, 1)] = (int32_t)slot;
        return ist[IS_STATUS] == ST_RUNNING;
    }
    constexpr int XNEW = N;
    int ist[IS_SIZE];
    {
        int4 t0 = *reinterpret_cast<const int4*>(ip);
        int4 t1 = *reinterpret_cast<const int4*>(ip + 4);
        ist[0] = t0.x; ist[1] = t0.y; ist[2] = t0.z; ist[3] = t0.w;
        ist[4] = t1.x; ist[5] = t1.y; ist[6] = t1.z; ist[7] = t1.w;
    }
    if (ist[IS_STATUS] != ST_RUNNING) return false;
    double st[SS];
    double* sp = state + pid * (int64_t)SS;
    if (SS % 2 == 0) {
#pragma unroll
        for (int i = 0; i < SS; i += 2) {
            double2 t = *reinterpret_cast<const double2*>(sp + i);
            st[i] = t.x;
            st[i + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < SS; i++) st[i] = sp[i];
    }
    double ln[L::SIZE];
    {
        const double* lp = lin + slot * (int64_t)L::SIZE;
#pragma unroll
        for (int i = 0; i < L::SIZE; i += 2) {
            double2 t = *reinterpret_cast<const double2*>(lp + i);
            ln[i] = t.x;
            ln[i + 1] = t.y;
        }
    }
    const bool go = dogbox_round<N>(st, ist, ln, x0 + pid * N, lb + pid * bstride,
                                    ub + pid * bstride, sc, P, first);
    if (SS % 2 == 0) {
#pragma unroll
        for (int i = 0; i < SS; i += 2)
            *reinterpret_cast<double2*>(sp + i) = make_double2(st[i], st[i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < SS; i++) sp[i] = st[i];
    }
    *reinterpret_cast<int4*>(ip) = make_int4(ist[0], ist[1], ist[2], ist[3]);
    *reinterpret_cast<int4*>(ip + 4) = make_int4(ist[4], ist[5], ist[6], ist[7]);
    if (go) {
#pragma unroll
        for (int i = 0; i < N; i++) Xnew[slot * N + i] = st[XNEW + i];
        if (Xjac) {
#pragma unroll
            for (int i = 0; i < N; i++) {
                double xj = st[XNEW + i];
                // the point dogbox.py:256-261 would evaluate J at
                int ob = ((ist[IS_FREE] >> i) & 1) ? get2(ist[IS_MARKS], i)
                                                   : get2(ist[IS_ONB], i);
                if (ob == -1) xj = lb[pid * bstride + i];
                if (ob == 1) xj = ub[pid * bstride + i];
                Xjac[slot * N + i] = xj;
            }
        }
    }
    return ist[IS_STATUS] == ST_RUNNING;
}

// CTAs per SM of the round kernel: TRF N <= 4 runs best at 4 x 128 threads
// (128 registers; 3 x 164 without spills measured 5 % slower), dogbox and the
// larger N at 2 (255 registers: C3, N = 6 dogbox, 0.60 ms per 10^6 problems at
// 4 CTAs, 0.50 at 3, 0.44 at 2 where it no longer spills).
template <int N, int METHOD> struct RoundCfg {
    static constexpr int MINB =
        (METHOD == BLSQ_METHOD_DOGBOX || N > 4) ? (BLSQ_ROUND_MINB > 2 ? 2 : BLSQ_ROUND_MINB)
                                                : BLSQ_ROUND_MINB;
};

template <int N, int METHOD, int MODE>
__global__ void __launch_bounds__(BLSQ_ROUND_THREADS, (RoundCfg<N, METHOD>::MINB))
round_kernel(int64_t A, const int32_t* __restrict__ idx,
             const double* __restrict__ lin, const double* __restrict__ x0,
             const double* __restrict__ lb, const double* __restrict__ ub,
             int bstride, const double* __restrict__ scaling, SolveParams P,
             int first, double* __restrict__ state,
             int32_t* __restrict__ istate, double* __restrict__ Xnew,
             double* __restrict__ Xjac, int32_t* __restrict__ work,
             int32_t* __restrict__ count) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (MODE == 2) {
        // the worklist length is only known on the device: a small grid
        // strides over it
        const int64_t cnt = work[0];
        for (int64_t w = tid; w < cnt; w += (int64_t)gridDim.x * blockDim.x)
            round_slot<N, METHOD, MODE>(work[1 + w], A, idx, lin, x0, lb, ub, bstride, scaling,
                                        P, first, state, istate, Xnew, Xjac, work);
        return;
    }
    const bool run = round_slot<N, METHOD, MODE>(tid, A, idx, lin, x0, lb, ub, bstride, scaling,
                                                 P, first, state, istate, Xnew, Xjac, work);
    if (!count) return;
    // running problems after this round: one atomic per CTA; the last CTA to
    // arrive publishes the total in count[2] and re-arms the two counters, so
    // the host needs no memset and no separate counting launch per round
    __shared__ int block_run;
    if (threadIdx.x == 0) block_run = 0;
    __syncthreads();
    const unsigned bal = __ballot_sync(0xffffffffu, run);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&block_run, __popc(bal));
    __syncthreads();
    if (threadIdx.x == 0) {
        if (block_run) atomicAdd(count, block_run);
        __threadfence();
        const int ticket = atomicAdd(count + 1, 1);
        if (ticket == (int)gridDim.x - 1) {
            __threadfence();
            count[2] = atomicExch(count, 0);
            count[1] = 0;
        }
    }
}

__global__ void init_kernel(int method, int64_t B, int n, int SS, int XNEW,
                            const double* __restrict__ x0,
                            const double* __restrict__ lb,
                            const double* __restrict__ ub, int bstride,
                            double* __restrict__ state,
                            int32_t* __restrict__ istate,
                            double* __restrict__ Xnew) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * n) return;
    int64_t b = t / n;
    int i = (int)(t % n);
    double x = x0[t];
    if (method == BLSQ_METHOD_TRF)
        x = strictly_feasible(x, lb[b * bstride + i], ub[b * bstride + i], 1e-10);
    state[b * SS + XNEW + i] = x;
    state[b * SS + i] = x;              // X
    Xnew[t] = x;
    if (i == 0) {
        int32_t* ip = istate + b * IS_SIZE;
        ip[IS_STATUS] = ST_RUNNING;
#pragma unroll
        for (int k = 1; k < IS_SIZE; k++) ip[k] = 0;
    }
}

__global__ void on_bound_kernel(int64_t B, int n,
                                const int32_t* __restrict__ istate,
                                int64_t* __restrict__ mask) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * n) return;
    int64_t b = t / n;
    int i = (int)(t % n);
    mask[t] = get2(istate[b * IS_SIZE + IS_ONB], i);
}

__global__ void count_running_kernel(int64_t B,
                                     const int32_t* __restrict__ idx,
                                     const int32_t* __restrict__ istate,
                                     int32_t* __restrict__ count) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool run = false;
    if (t < B) {
        int64_t pid = idx ? idx[t] : t;
        run = istate[pid * IS_SIZE + IS_STATUS] == ST_RUNNING;
    }
    unsigned bal = __ballot_sync(0xffffffffu, run);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(count, __popc(bal));
}


// ---- ordered compaction of the active set (three small launches) ------------
// Slot s (problem idx[s], or s) survives when its problem is still running;
// survivors keep their order.  Each block owns CMP_SLOTS consecutive slots.
constexpr int CMP_THREADS = 256;
constexpr int CMP_PER = 4;
constexpr int CMP_SLOTS = CMP_THREADS * CMP_PER;

__device__ __forceinline__ int compact_local(int64_t A, const int32_t* __restrict__ idx,
                                             const int32_t* __restrict__ istate,
                                             int64_t base, bool (&keep)[CMP_PER],
                                             int64_t (&pid)[CMP_PER], int& block_total) {
    __shared__ int wsum[CMP_THREADS / 32];
    int c = 0;
#pragma unroll
    for (int j = 0; j < CMP_PER; j++) {
        const int64_t s = base + (int64_t)threadIdx.x * CMP_PER + j;
        keep[j] = false;
        pid[j] = 0;
        if (s < A) {
            pid[j] = idx ? idx[s] : s;
            keep[j] = istate[pid[j] * IS_SIZE + IS_STATUS] == ST_RUNNING;
        }
        c += keep[j];
    }
    // exclusive scan of c over the block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < CMP_THREADS / 32; w++) {
        const int v = wsum[w];
        if (w < warp) before += v;
        total += v;
    }
    block_total = total;
    return before + inc - c;
}

__global__ void __launch_bounds__(CMP_THREADS)
compact_count_kernel(int64_t A, const int32_t* __restrict__ idx,
                     const int32_t* __restrict__ istate, int32_t* __restrict__ bsum) {
    bool keep[CMP_PER];
    int64_t pid[CMP_PER];
    int total;
    compact_local(A, idx, istate, (int64_t)blockIdx.x * CMP_SLOTS, keep, pid, total);
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

// exclusive scan of the block totals by one block (serial chunks per thread)
__global__ void __launch_bounds__(1024)
compact_scan_kernel(int nblocks, int32_t* __restrict__ bsum) {
    __shared__ int part[1024];
    const int per = (nblocks + 1023) / 1024;
    const int b0 = threadIdx.x * per;
    int s = 0;
    for (int k = 0; k < per; k++)
        if (b0 + k < nblocks) s += bsum[b0 + k];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int t = 0; t < 1024; t++) { const int v = part[t]; part[t] = run; run += v; }
        bsum[nblocks] = run;            // the number of survivors, for the host
    }
    __syncthreads();
    int run = part[threadIdx.x];
    for (int k = 0; k < per; k++)
        if (b0 + k < nblocks) { const int v = bsum[b0 + k]; bsum[b0 + k] = run; run += v; }
}

__global__ void __launch_bounds__(CMP_THREADS)
compact_scatter_kernel(int64_t A, const int32_t* __restrict__ idx,
                       const int32_t* __restrict__ istate, int n,
                       const int32_t* __restrict__ boff,
                       const double* __restrict__ Xnew, const double* __restrict__ Xjac,
                       int32_t* __restrict__ idx_out, int64_t* __restrict__ idx64_out,
                       double* __restrict__ Xnew_out, double* __restrict__ Xjac_out) {
    bool keep[CMP_PER];
    int64_t pid[CMP_PER];
    int total;
    const int64_t base = (int64_t)blockIdx.x * CMP_SLOTS;
    int pos = boff[blockIdx.x] + compact_local(A, idx, istate, base, keep, pid, total);
#pragma unroll
    for (int j = 0; j < CMP_PER; j++) {
        if (!keep[j]) continue;
        const int64_t s = base + (int64_t)threadIdx.x * CMP_PER + j;
        idx_out[pos] = (int32_t)pid[j];
        if (idx64_out) idx64_out[pos] = pid[j];
        for (int i = 0; i < n; i++) {
            Xnew_out[(int64_t)pos * n + i] = Xnew[s * n + i];
            if (Xjac) Xjac_out[(int64_t)pos * n + i] = Xjac[s * n + i];
        }
        pos++;
    }
}

template <int N, int G, bool MULTI>
int launch_lin_g(int64_t A, const int32_t* idx, int m, const double* F,
                 const double* J, const PtrList<N>& pl, const double* dx,
                 int jac_mode, const int32_t* istate, double* lin,
                 cudaStream_t s) {
    int64_t threads = A * G;
    int64_t blocks = (threads + BLSQ_LIN_THREADS - 1) / BLSQ_LIN_THREADS;
    if (blocks > 0x7fffffff) return BLSQ_E_UNSUPPORTED;
    if (jac_mode == 0)
        lin_kernel<N, G, 0, MULTI><<<(unsigned)blocks, BLSQ_LIN_THREADS, 0, s>>>(
            A, idx, m, F, J, pl, dx, istate, lin);
    else if (jac_mode == 1)
        lin_kernel<N, G, 1, MULTI><<<(unsigned)blocks, BLSQ_LIN_THREADS, 0, s>>>(
            A, idx, m, F, J, pl, dx, istate, lin);
    else
        lin_kernel<N, G, 2, MULTI><<<(unsigned)blocks, BLSQ_LIN_THREADS, 0, s>>>(
            A, idx, m, F, J, pl, dx, istate, lin);
    BLSQ_LAUNCH_CHECK();
    return 0;
}

template <int N>
int launch_lin(int64_t A, const int32_t* idx, int m, const double* F,
               const double* J, const double* const* Fp_host, const double* dx,
               int jac_mode, const int32_t* istate, double* lin,
               cudaStream_t s) {
    constexpr int RPL = LinCfg<N>::RPL;
    PtrList<N> pl;
    const int np = jac_mode == 2 ? 2 * N : (jac_mode == 1 ? N : 0);
    for (int j = 0; j < 2 * N; j++) pl.p[j] = (j < np) ? Fp_host[j] : nullptr;
    // smallest lane group whose registers hold all m rows (else 32 + chunks)
    if (m <= 8 * RPL)
        return launch_lin_g<N, 8, false>(A, idx, m, F, J, pl, dx, jac_mode, istate, lin, s);
    if (m <= 16 * RPL)
        return launch_lin_g<N, 16, false>(A, idx, m, F, J, pl, dx, jac_mode, istate, lin, s);
    if (m <= 32 * RPL)
        return launch_lin_g<N, 32, false>(A, idx, m, F, J, pl, dx, jac_mode, istate, lin, s);
    return launch_lin_g<N, 32, true>(A, idx, m, F, J, pl, dx, jac_mode, istate, lin, s);
}

template <int N>
int launch_round(int method, int64_t A, const int32_t* idx, const double* lin,
                 const double* x0, const double* lb, const double* ub,
                 int bstride, const double* scaling, SolveParams P, int first,
                 double* state, int32_t* istate, double* Xnew, double* Xjac,
                 int32_t* work, int32_t* count, cudaStream_t s) {
    int64_t blocks = (A + BLSQ_ROUND_THREADS - 1) / BLSQ_ROUND_THREADS;
    if (blocks > 0x7fffffff) return BLSQ_E_UNSUPPORTED;
    const unsigned gb = (unsigned)blocks;
    if (method == BLSQ_METHOD_TRF && work) {
        cudaError_t e = cudaMemsetAsync(work, 0, sizeof(int32_t), s);
        if (e != cudaSuccess) return (int)e;
        round_kernel<N, BLSQ_METHOD_TRF, 1><<<gb, BLSQ_ROUND_THREADS, 0, s>>>(
            A, idx, lin, x0, lb, ub, bstride, scaling, P, first, state, istate, Xnew, Xjac, work,
            count);
        BLSQ_LAUNCH_CHECK();
        // ~5 % of the problems land on the worklist: an eighth of the grid
        // (at least 4 CTAs per SM) strides over it
        unsigned g2 = gb / 8 > 592u ? gb / 8 : (gb < 592u ? gb : 592u);
        round_kernel<N, BLSQ_METHOD_TRF, 2><<<g2, BLSQ_ROUND_THREADS, 0, s>>>(
            A, idx, lin, x0, lb, ub, bstride, scaling, P, first, state, istate, Xnew, Xjac, work,
            nullptr);
    } else if (method == BLSQ_METHOD_TRF) {
        round_kernel<N, BLSQ_METHOD_TRF, 0><<<gb, BLSQ_ROUND_THREADS, 0, s>>>(
            A, idx, lin, x0, lb, ub, bstride, scaling, P, first, state, istate, Xnew, Xjac, work,
            count);
    } else {
        round_kernel<N, BLSQ_METHOD_DOGBOX, 0><<<gb, BLSQ_ROUND_THREADS, 0, s>>>(
            A, idx, lin, x0, lb, ub, bstride, scaling, P, first, state, istate, Xnew, Xjac, work,
            count);
    }
    BLSQ_LAUNCH_CHECK();
    return 0;
}

}  // namespace

#ifdef BLSQ_ONLY_N46      /* quick builds for tools/build_variants.sh */
#define BLSQ_DISPATCH_N(n, CALL)          \
    switch (n) {                          \
        case 4: { constexpr int N_ = 4; CALL; } break; \
        case 6: { constexpr int N_ = 6; CALL; } break; \
        default: return BLSQ_E_UNSUPPORTED; \
    }
#else
#define BLSQ_DISPATCH_N(n, CALL)          \
    switch (n) {                          \
        case 1: { constexpr int N_ = 1; CALL; } break; \
        case 2: { constexpr int N_ = 2; CALL; } break; \
        case 3: { constexpr int N_ = 3; CALL; } break; \
        case 4: { constexpr int N_ = 4; CALL; } break; \
        case 5: { constexpr int N_ = 5; CALL; } break; \
        case 6: { constexpr int N_ = 6; CALL; } break; \
        case 7: { constexpr int N_ = 7; CALL; } break; \
        case 8: { constexpr int N_ = 8; CALL; } break; \
        default: return BLSQ_E_UNSUPPORTED; \
    }
#endif

extern "C" {

int blsq_state_layout(int method, int n, int* out) {
    if (!out) return BLSQ_E_BADARG;
    if (method != BLSQ_METHOD_TRF && method != BLSQ_METHOD_DOGBOX) return BLSQ_E_BADARG;
    BLSQ_DISPATCH_N(n, {
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
    BLSQ_DISPATCH_N(n, { return LinRec<N_>::SIZE; });
    return BLSQ_E_UNSUPPORTED;
}

int blsq_init_batched(int method, int64_t B, int n, const double* x0,
                      const double* lb, const double* ub, int bstride,
                      double* state, int32_t* istate, double* Xnew,
                      void* stream) {
    if (B < 0 || !x0 || !lb || !ub || !state || !istate || !Xnew) return BLSQ_E_BADARG;
    if (bstride != 0 && bstride != n) return BLSQ_E_BADARG;
    int lay[9];
    int rc = blsq_state_layout(method, n, lay);
    if (rc) return rc;
    if (B == 0) return 0;
    int64_t blocks = (B * n + 255) / 256;
    init_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        method, B, n, lay[0], lay[2], x0, lb, ub, bstride, state, istate, Xnew);
    BLSQ_LAUNCH_CHECK();
    return 0;
}

int blsq_linearise_batched(int64_t A, const int32_t* idx, int m, int n,
                           const double* F, const double* J,
                           const double* const* Fp_host, const double* dx,
                           int jac_mode, const int32_t* istate, double* lin,
                           void* stream) {
    if (A < 0 || m < 1 || !F || !istate || !lin) return BLSQ_E_BADARG;
    if (jac_mode < 0 || jac_mode > 2) return BLSQ_E_BADARG;
    if (jac_mode == 0 && !J) return BLSQ_E_BADARG;
    if (jac_mode != 0 && (!dx || !Fp_host)) return BLSQ_E_BADARG;
    if (jac_mode != 0 && n >= 1 && n <= BLSQ_MAX_BATCHED_N)
        for (int j = 0; j < jac_mode * n; j++)
            if (!Fp_host[j]) return BLSQ_E_BADARG;
    if (A == 0) return 0;
    BLSQ_DISPATCH_N(n, {
        return launch_lin<N_>(A, idx, m, F, J, Fp_host, dx, jac_mode, istate,
                              lin, (cudaStream_t)stream);
    });
    return 0;
}

int blsq_round_batched(int method, int64_t A, const int32_t* idx, int m, int n,
                       const double* lin, const double* x0, const double* lb,
                       const double* ub, int bstride, const double* scaling,
                       double ftol, double xtol, double gtol, int max_nfev,
                       int first, double* state, int32_t* istate, double* Xnew,
                       double* Xjac, int32_t* work, int32_t* count, void* stream) {
    if (A < 0 || !lin || !x0 || !lb || !ub || !state || !istate || !Xnew) return BLSQ_E_BADARG;
    if (method != BLSQ_METHOD_TRF && method != BLSQ_METHOD_DOGBOX) return BLSQ_E_BADARG;
    if (bstride != 0 && bstride != n) return BLSQ_E_BADARG;
    if (A == 0) {
        if (count) {
            cudaError_t e = cudaMemsetAsync(count, 0, 3 * sizeof(int32_t), (cudaStream_t)stream);
            if (e != cudaSuccess) return (int)e;
        }
        return 0;
    }
    SolveParams P;
    P.ftol = ftol; P.xtol = xtol; P.gtol = gtol;
    P.max_nfev = max_nfev; P.m = m; P.jac_scaling = scaling ? 0 : 1;
    BLSQ_DISPATCH_N(n, {
        return launch_round<N_>(method, A, idx, lin, x0, lb, ub, bstride,
                                scaling, P, first, state, istate, Xnew, Xjac,
                                work, count, (cudaStream_t)stream);
    });
    return 0;
}

int blsq_dogbox_on_bound(int64_t B, int n, const int32_t* istate,
                         int64_t* mask, void* stream) {
    if (B < 0 || n < 1 || n > BLSQ_MAX_BATCHED_N || !istate || !mask) return BLSQ_E_BADARG;
    if (B == 0) return 0;
    int64_t blocks = (B * n + 255) / 256;
    on_bound_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(B, n, istate, mask);
    BLSQ_LAUNCH_CHECK();
    return 0;
}

int blsq_count_running(int64_t B, const int32_t* idx, const int32_t* istate,
                       int32_t* count, void* stream) {
    if (B < 0 || !istate || !count) return BLSQ_E_BADARG;
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int32_t), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    if (B == 0) return 0;
    int64_t blocks = (B + 255) / 256;
    count_running_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(B, idx, istate, count);
    BLSQ_LAUNCH_CHECK();
    return 0;
}

int64_t blsq_compact_work_size(int64_t A) {
    if (A < 0) return BLSQ_E_BADARG;
    return (A + CMP_SLOTS - 1) / CMP_SLOTS + 1;
}

int blsq_compact_batched(int64_t A, const int32_t* idx, const int32_t* istate, int n,
                         const double* Xnew, const double* Xjac, int32_t* idx_out,
                         int64_t* idx64_out, double* Xnew_out, double* Xjac_out,
                         int32_t* work, void* stream) {
    if (A < 0 || n < 1 || !istate || !Xnew || !idx_out || !Xnew_out || !work) return BLSQ_E_BADARG;
    if (Xjac && !Xjac_out) return BLSQ_E_BADARG;
    if (A >= ((int64_t)1 << 31)) return BLSQ_E_UNSUPPORTED;      /* problem ids are int32 */
    if (A == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const int nblocks = (int)((A + CMP_SLOTS - 1) / CMP_SLOTS);
    compact_count_kernel<<<nblocks, CMP_THREADS, 0, s>>>(A, idx, istate, work);
    BLSQ_LAUNCH_CHECK();
    compact_scan_kernel<<<1, 1024, 0, s>>>(nblocks, work);
    BLSQ_LAUNCH_CHECK();
    compact_scatter_kernel<<<nblocks, CMP_THREADS, 0, s>>>(A, idx, istate, n, work, Xnew, Xjac,
                                                          idx_out, idx64_out, Xnew_out, Xjac_out);
    BLSQ_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

#ifdef BLSQ_PHASE_CLOCKS
// tools/phase_probe.py only: cycles [0..15] and visits [16..31] per phase of
// trf_round_impl since the last reset
extern "C" int blsq_debug_phase_read(unsigned long long* out, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out, blsq_phase_acc, sizeof(blsq_phase_acc));
    if (e == cudaSuccess && reset) {
        unsigned long long z[32] = {0};
        e = cudaMemcpyToSymbol(blsq_phase_acc, z, sizeof(z));
    }
    return (int)e;
}
#endif
