// Tall mode, the O(m n^2) part: Gram passes of CholeskyQR2 on [J | f] with
// FP64 tensor-core tiles (DMMA, mma.sync.m8n8k4.f64) staged through TMA bulk
// copies, plus the ||f||^2 reduction of a trial evaluation.
//
// Stands in for the dense factorisations of the reference on one tall problem:
//   trf.py:244      g = J^T f
//   trf.py:264-274  svd([J_h; diag]) -- via J = QR, see DESIGN.md section 2.2
//   dogbox.py:170,197-199  J^T f, lstsq(J_free, -f), J_free g_free
//
// gram_kernel<NB, PASS>: persistent CTAs (one per SM).  One producer warp
// streams 64-row tiles of J (row-major, n = up to 8*NB columns) and f into a
// ring of shared-memory stages, one cp.async.bulk per row so that the rows can
// be padded (LD = 8*NB + 4 doubles) and every fragment load below is bank
// conflict free; mbarriers carry the full/empty hand-shake.  Consumer warps:
//   PASS 2 only:  Y = tile * Rinv (Rinv = R1^-1 upper triangular) by DMMA into
//                 a second shared buffer,
//   both passes:  the upper 8x8 blocks of G += Y^T Y (or J^T J) by DMMA, the
//                 blocks split over ROLES warp roles, the rows of the tile over
//                 the KS warps of a role; role 0 also accumulates Y^T f, f.f.
// Partials are written per CTA and summed in CTA order by gram_reduce_kernel,
// so the result is deterministic for a given device.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <utility>

#include "../../include/blsq.h"
#include "blsq_tall_common.cuh"

#define BLSQ_LAUNCH_CHECK()                                  \
    do {                                                     \
        cudaError_t e_ = cudaGetLastError();                 \
        if (e_ != cudaSuccess) return (int)e_;               \
    } while (0)

namespace {

// ---- PTX helpers ----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// D(8x8) += A(8x4, row) * B(4x8, col); lane holds a = A[lane/4][lane%4],
// b = B[lane%4][lane/4], c = C[lane/4][2*(lane%4) + {0,1}]
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c[0]), "+d"(c[1])
        : "d"(a), "d"(b));
}

// ---- block bookkeeping ------------------------------------------------------
// upper-triangle blocks (i <= j < NB) enumerated row by row
template <int NB>
struct Tri {
    static constexpr int COUNT = NB * (NB + 1) / 2;
    __host__ __device__ static constexpr int row(int q) {
        int i = 0;
        while (q >= NB - i) { q -= NB - i; i++; }
        return i;
    }
    __host__ __device__ static constexpr int col(int q) {
        int i = 0;
        while (q >= NB - i) { q -= NB - i; i++; }
        return i + q;
    }
    __host__ __device__ static constexpr int index(int i, int j) {
        return i * NB - (i * (i - 1)) / 2 + (j - i);
    }
};

template <int NB> struct GramCfg;
// T rows per tile, S1/S2 ring stages in pass 1/2, CW consumer warps, R1/R2
// warp roles over the Gram blocks in pass 1/2, CG column blocks per phase-A
// item, YB2: pass 2 double-buffers Y so that one block barrier per tile suffices
#ifndef BLSQ_G8_CW          /* tuning knobs of the n = 64 kernels (tools/build_gram_variants.sh) */
#define BLSQ_G8_CW 8
#define BLSQ_G8_R1 2
#define BLSQ_G8_R2 4
#define BLSQ_G8_CG 8
#define BLSQ_G8_S1 4
#define BLSQ_G8_S2 3
#define BLSQ_G8_YB2 true
#endif
template <> struct GramCfg<2>  { static constexpr int T = 64, S1 = 4, S2 = 4, CW = 4, R1 = 1, R2 = 1, CG = 2; static constexpr bool RINV_SMEM = true,  YB2 = true; };
template <> struct GramCfg<4>  { static constexpr int T = 64, S1 = 4, S2 = 4, CW = 8, R1 = 1, R2 = 1, CG = 4; static constexpr bool RINV_SMEM = true,  YB2 = true; };
template <> struct GramCfg<8>  { static constexpr int T = 64, S1 = BLSQ_G8_S1, S2 = BLSQ_G8_S2, CW = BLSQ_G8_CW, R1 = BLSQ_G8_R1, R2 = BLSQ_G8_R2, CG = BLSQ_G8_CG; static constexpr bool RINV_SMEM = true,  YB2 = BLSQ_G8_YB2; };
template <> struct GramCfg<16> { static constexpr int T = 64, S1 = 2, S2 = 2, CW = 8, R1 = 8, R2 = 8, CG = 8; static constexpr bool RINV_SMEM = false, YB2 = false; };
template <> struct GramCfg<32> { static constexpr int T = 32, S1 = 2, S2 = 2, CW = 8, R1 = 8, R2 = 8, CG = 8; static constexpr bool RINV_SMEM = false, YB2 = false; };

// Shared-memory layout of one tile of T rows.  The tile is kept as four
// QUARTERS of QR = T/4 consecutive rows, each quarter dense (row stride RS = n)
// so that ONE bulk copy moves it, and quarter h starts 4 doubles (8 banks)
// after a multiple of 32 banks.  A DMMA k-chunk takes its four k rows from the
// four quarters (row kc of each), which makes every fragment load conflict
// free: lanes (lr, lc) read double lr of row kc in quarter lc.
template <int NB, int PASS>
struct GramLayout {
    typedef GramCfg<NB> C;
    static constexpr int W = 8 * NB;            // padded column count
    static constexpr int QR = C::T / 4;         // rows per quarter
    static constexpr int QS = QR * W + 4;       // quarter stride (doubles)
    static constexpr int NBLK = Tri<NB>::COUNT;
    // Blocks per role.  In pass 2 role 0 also accumulates Y^T f and J^T f
    // (2 NB DFMAs per chunk = NB/4 DMMAs of pipe time), so it gets fewer blocks.
    // SPLIT: the Gram blocks are divided over SPLIT CTAs (consecutive
    // blockIdx.x) that stream the same tiles.  At n = 256 the upper half of
    // the Gram matrix is 263 KB of accumulators -- more than the register file
    // of one SM -- so one CTA cannot hold it without spilling; two CTAs hold
    // 33-34 blocks per warp each and J is read twice (it is compute bound by
    // 5x).  Pass 1 only: measured 8.7 -> 2.2 ms per 2^20 rows (85 % of the
    // FP64 peak on the executed blocks); in pass 2 both CTAs would have to
    // repeat the whole Y = J R1^-1 product, which dominates it (43.7 -> 59.6
    // ms), so pass 2 stays one CTA until Y can be shared through a cluster.
    // ROLES_CTA = roles inside one CTA, ROLES = roles over the block list.
    static constexpr int SPLIT = (NB >= 32 && PASS == 1) ? 2 : 1;
    static constexpr int ROLES_CTA = (PASS == 2) ? C::R2 : C::R1;
    static constexpr int ROLES = ROLES_CTA * SPLIT;
    static constexpr int FW = (PASS == 2) ? (NB + 3) / 4 : 0;
    static constexpr int B0_ = (NBLK + FW) / ROLES - FW;
    static constexpr int B0 = (ROLES == 1) ? NBLK : (B0_ < 1 ? 1 : B0_);         // role 0
    static constexpr int BPR = (ROLES == 1) ? NBLK : (NBLK - B0 + ROLES - 2) / (ROLES - 1);
    __host__ __device__ static constexpr int role_begin(int r) {
        return r == 0 ? 0 : B0 + (r - 1) * BPR;
    }
    __host__ __device__ static constexpr int role_count(int r) {
        return r == 0 ? B0
                      : (role_begin(r) >= NBLK ? 0
                                               : (role_begin(r) + BPR <= NBLK ? BPR
                                                                              : NBLK - role_begin(r)));
    }
    static constexpr int MAXQ = B0 > BPR ? B0 : BPR;
    static constexpr int KS = C::CW / ROLES_CTA;                   // row split
    static constexpr int THREADS = (C::CW + 1) * 32;
    static constexpr int S = (PASS == 2) ? C::S2 : C::S1;          // ring stages
    static constexpr int STAGE_DOUBLES = 4 * QS + C::T;            // tile + f (pass 2)
    static constexpr int YBUF_DOUBLES = (PASS == 2) ? (C::YB2 ? 8 : 4) * QS : 0;
    static constexpr int RINV_DOUBLES = (PASS == 2 && C::RINV_SMEM) ? NBLK * 64 : 0;
    static constexpr int PER = NBLK * 64 + 2 * W + 2;              // one warp copy (even: double2 stores)
    static constexpr int RED_DOUBLES = (KS > 1) ? KS * PER : 0;
    static constexpr int RING_DOUBLES = S * STAGE_DOUBLES;
    // the cross-warp reduction reuses the ring once the pipeline has drained
    static constexpr int BASE_DOUBLES = RING_DOUBLES > RED_DOUBLES ? RING_DOUBLES : RED_DOUBLES;
    static constexpr int MAIN_DOUBLES = BASE_DOUBLES + YBUF_DOUBLES + RINV_DOUBLES;
    static constexpr size_t SMEM_BYTES = (size_t)MAIN_DOUBLES * 8 + 2 * S * 8 + 16;
    // per-CTA partial record: G (W x W, upper blocks) | Y^T f (W) | f.f, pad |
    // J^T f (W); the vectors and f.f are produced by pass 2 only
    static constexpr int OFF_V1 = W * W, OFF_FF = W * W + W, OFF_V2 = W * W + W + 2;
    static constexpr int REC = W * W + 2 * W + 2;
};

// Pass 1 may run on a SAMPLE of the tiles (it only has to deliver a
// preconditioner R1 with cond(J R1^-1) = O(1); pass 2 sees every row and
// blsq_tall_factor verifies the result): one tile out of every `sstride`
// consecutive ones, at a hashed position inside the group.
__device__ __forceinline__ int64_t sample_tile(int64_t j, int sstride) {
    if (sstride <= 1) return j;
    uint32_t h = (uint32_t)j * 2654435761u;
    h ^= h >> 15;
    return j * sstride + (int64_t)(h % (uint32_t)sstride);
}

template <int NB, int Q0, int... Qs>
__device__ __forceinline__ void mma_role_blocks(double (&acc)[sizeof...(Qs)][2],
                                                const double (&frag)[NB],
                                                std::integer_sequence<int, Qs...>) {
    (dmma(acc[Qs], frag[Tri<NB>::row(Q0 + Qs)], frag[Tri<NB>::col(Q0 + Qs)]), ...);
}

// EXACT: n == 8*NB (no column padding, row stride is a compile-time constant)
template <int NB, int PASS, bool EXACT, int ROLE>
__device__ __forceinline__ void consumer_loop(int64_t njobs, const double* __restrict__ rinvp_g,
                                              int n, double* smem, uint64_t* full,
                                              uint64_t* empty, double* __restrict__ rec) {
    typedef GramCfg<NB> C;
    typedef GramLayout<NB, PASS> L;
    constexpr int T = C::T, S = L::S, CW = C::CW, W = L::W, QR = L::QR, QS = L::QS;
    constexpr int Q0 = L::role_begin(ROLE);
    constexpr int NQ = L::role_count(ROLE);
    static_assert(NQ > 0, "every role must own at least one Gram block");
    constexpr int KS = L::KS;
    constexpr int KU = (NB >= 32) ? 1 : 2;      // two k-chunks of fragments do not fit at NB = 32
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ks = warp / L::ROLES_CTA;
    const int lr = lane >> 2, lc = lane & 3;
    const int RS = EXACT ? W : n;               // row stride inside a quarter

    double* ybuf0 = smem + L::BASE_DOUBLES;
    const double* rinvp = C::RINV_SMEM ? (ybuf0 + L::YBUF_DOUBLES) : rinvp_g;
    int ypar = 0;

    double acc[NQ][2];
#pragma unroll
    for (int q = 0; q < NQ; q++) { acc[q][0] = 0.0; acc[q][1] = 0.0; }
    // pass 2, role 0: Y^T f, J^T f (trf.py:244 / dogbox.py:170) and f.f
    double gf[NB], gj[NB], ff = 0.0;
#pragma unroll
    for (int b = 0; b < NB; b++) { gf[b] = 0.0; gj[b] = 0.0; }

    int stage = 0;
    uint32_t phase = 0;
    for (int64_t t = blockIdx.x / L::SPLIT; t < njobs; t += gridDim.x / L::SPLIT) {
        mbar_wait(&full[stage], phase);
        const double* tj = smem + (size_t)stage * L::STAGE_DOUBLES;
        const double* tf = tj + 4 * QS;
        double* ybuf = ybuf0 + (C::YB2 ? ypar * 4 * QS : 0);
        if (PASS == 2) {
            // ---- phase A: Y = tile * Rinv; a strip is 8 rows (two from each
            //      quarter), an item is a strip x CG column blocks ----
            constexpr int NSTRIP = T / 8;
            constexpr int NCG = NB / C::CG;
            constexpr int ITEMS = NSTRIP * NCG;
            // Everyone must be done reading this Y buffer.  Double buffered
            // (YB2): its last readers were phase B two tiles ago, and every
            // warp has passed the barrier below once since -> no barrier here.
            if (!C::YB2) named_bar_sync(1, CW * 32);
            for (int it = warp; it < ITEMS; it += CW) {
                const int strip = it % NSTRIP;
                const int cg = (NCG == 1) ? 0 : it / NSTRIP;
                double y[C::CG][2];
#pragma unroll
                for (int jj = 0; jj < C::CG; jj++) { y[jj][0] = 0.0; y[jj][1] = 0.0; }
                const int qrow = 2 * strip + (lr >> 2);
                const double* arow = tj + (lr & 3) * QS + qrow * RS + lc;
                if (NCG == 1) {
#pragma unroll
                    for (int kc = 0; kc < 2 * NB; kc++) {
                        const double a = (EXACT || kc * 4 + lc < n) ? arow[kc * 4] : 0.0;
#pragma unroll
                        for (int jj = 0; jj < C::CG; jj++) {
                            if (kc <= 2 * jj + 1) {
                                const int q = Tri<NB>::index(kc >> 1, jj);
                                const double b = C::RINV_SMEM
                                    ? rinvp[(q * 2 + (kc & 1)) * 32 + lane]
                                    : __ldg(rinvp + (q * 2 + (kc & 1)) * 32 + lane);
                                dmma(y[jj], a, b);
                            }
                        }
                    }
                } else {
                    const int kmax = 2 * (cg * C::CG + C::CG);     // chunks of 4 columns
                    for (int kc = 0; kc < kmax; kc++) {
                        const double a = (EXACT || kc * 4 + lc < n) ? arow[kc * 4] : 0.0;
                        const int kb = kc >> 1;
                        const int qrow0 = kb * NB - (kb * (kb - 1)) / 2 - kb;   // + j
#pragma unroll
                        for (int jj = 0; jj < C::CG; jj++) {
                            const int j = cg * C::CG + jj;
                            if (kc <= 2 * j + 1) {
                                const int q = qrow0 + j;
                                const double b = C::RINV_SMEM
                                    ? rinvp[(q * 2 + (kc & 1)) * 32 + lane]
                                    : __ldg(rinvp + (q * 2 + (kc & 1)) * 32 + lane);
                                dmma(y[jj], a, b);
                            }
                        }
                    }
                }
                // two 8-byte stores per block, the element order swapped on odd
                // rows: the 16 lanes of a store phase then hit 32 distinct banks
                // (a 16-byte store of the C fragment is 2-way conflicted here)
                double* yrow = ybuf + (lr & 3) * QS + qrow * W + 2 * lc;
                const int odd = lr & 1;
#pragma unroll
                for (int jj = 0; jj < C::CG; jj++) {
                    const int j = cg * C::CG + jj;
                    yrow[8 * j + odd] = odd ? y[jj][1] : y[jj][0];
                    yrow[8 * j + 1 - odd] = odd ? y[jj][0] : y[jj][1];
                }
            }
            named_bar_sync(1, CW * 32);
        }
        // ---- phase B: G += src^T src; chunk kc = row kc of the four quarters ----
        const double* src = (PASS == 2) ? ybuf : tj;
        const int SRS = (PASS == 2) ? W : RS;
#pragma unroll KU
        for (int kc = ks; kc < QR; kc += KS) {
            const double* base = src + lc * QS + kc * SRS + lr;
            double frag[NB];
#pragma unroll
            for (int b = 0; b < NB; b++)
                frag[b] = (PASS == 2 || EXACT || 8 * b + lr < n) ? base[8 * b] : 0.0;
            mma_role_blocks<NB, Q0>(acc, frag, std::make_integer_sequence<int, NQ>{});
            if (PASS == 2 && ROLE == 0) {
                const double fk = tf[lc * QR + kc];
                const double* rawb = tj + lc * QS + kc * RS + lr;
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    gf[b] = fma(frag[b], fk, gf[b]);
                    const double raw = (EXACT || 8 * b + lr < n) ? rawb[8 * b] : 0.0;
                    gj[b] = fma(raw, fk, gj[b]);
                }
                if (lr == 0) ff = fma(fk, fk, ff);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == S) { stage = 0; phase ^= 1; }
        ypar ^= 1;
    }

    // ---- fold the KS row-split copies and write this CTA's record ----
    if (PASS == 2 && ROLE == 0) {
#pragma unroll
        for (int b = 0; b < NB; b++) {
            gf[b] += __shfl_xor_sync(0xffffffffu, gf[b], 1);
            gf[b] += __shfl_xor_sync(0xffffffffu, gf[b], 2);
            gj[b] += __shfl_xor_sync(0xffffffffu, gj[b], 1);
            gj[b] += __shfl_xor_sync(0xffffffffu, gj[b], 2);
        }
        ff += __shfl_xor_sync(0xffffffffu, ff, 1);
        ff += __shfl_xor_sync(0xffffffffu, ff, 2);
    }
    if (KS == 1) {
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const int bi = Tri<NB>::row(Q0 + q), bj = Tri<NB>::col(Q0 + q);
            double* dst = rec + (size_t)(8 * bi + lr) * W + 8 * bj + 2 * lc;
            *reinterpret_cast<double2*>(dst) = make_double2(acc[q][0], acc[q][1]);
        }
        if (PASS == 2 && ROLE == 0) {
            if (lc == 0) {
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    rec[L::OFF_V1 + 8 * b + lr] = gf[b];
                    rec[L::OFF_V2 + 8 * b + lr] = gj[b];
                }
            }
            if (lane == 0) rec[L::OFF_FF] = ff;
        }
        return;
    }
    // KS > 1: through shared memory.  Every bulk copy has landed (each was
    // waited for); wait until every consumer warp has left its last tile
    // before the ring is reused.
    named_bar_sync(2, CW * 32);
    double* red = smem + (size_t)ks * L::PER;
#pragma unroll
    for (int q = 0; q < NQ; q++)
        *reinterpret_cast<double2*>(red + (Q0 + q) * 64 + lane * 2) =
            make_double2(acc[q][0], acc[q][1]);
    if (PASS == 2 && ROLE == 0) {
        if (lc == 0) {
#pragma unroll
            for (int b = 0; b < NB; b++) {
                red[L::NBLK * 64 + 8 * b + lr] = gf[b];
                red[L::NBLK * 64 + W + 2 + 8 * b + lr] = gj[b];
            }
        }
        if (lane == 0) red[L::NBLK * 64 + W] = ff;
    }
}

template <int NB, int PASS, bool EXACT, int... ROLES_>
__device__ __forceinline__ void consumer_dispatch(int role, int64_t njobs, const double* rinvp_g,
                                                  int n, double* smem, uint64_t* full,
                                                  uint64_t* empty, double* rec,
                                                  std::integer_sequence<int, ROLES_...>) {
    ((role == ROLES_
          ? (consumer_loop<NB, PASS, EXACT, ROLES_>(njobs, rinvp_g, n, smem, full, empty, rec), 0)
          : 0),
     ...);
}

template <int NB, int PASS, bool EXACT>
__global__ void __launch_bounds__(GramLayout<NB, PASS>::THREADS, 1)
gram_kernel(int64_t m, int n, const double* __restrict__ J, const double* __restrict__ f,
            const double* __restrict__ rinvp_g, int sstride, double* __restrict__ partial) {
    typedef GramCfg<NB> C;
    typedef GramLayout<NB, PASS> L;
    constexpr int T = C::T, S = L::S, CW = C::CW, W = L::W, QR = L::QR, QS = L::QS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* smem = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::MAIN_DOUBLES);
    uint64_t* empty = full + S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* rec = partial + (size_t)blockIdx.x * L::REC;
    const int RS = EXACT ? W : n;

    for (int i = threadIdx.x; i < L::MAIN_DOUBLES; i += L::THREADS) smem[i] = 0.0;
    for (int i = threadIdx.x; i < L::REC; i += L::THREADS) rec[i] = 0.0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (PASS == 2 && C::RINV_SMEM) {
        double* rinv_s = smem + L::BASE_DOUBLES + L::YBUF_DOUBLES;
        for (int e = threadIdx.x; e < L::NBLK * 64; e += L::THREADS) rinv_s[e] = rinvp_g[e];
    }
    // make the generic-proxy zero fill visible before the async proxy writes
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    const int64_t ntiles = (m + T - 1) / T;
    // jobs = tiles (pass 2, or pass 1 on every row) or sampled tiles (pass 1)
    const int64_t njobs = (PASS == 1 && sstride > 1) ? ntiles / sstride : ntiles;
    if (warp == CW) {
        // ---------------- producer ----------------
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t q_bytes = (uint32_t)(QR * n) * 8u;
        for (int64_t job = blockIdx.x / L::SPLIT; job < njobs; job += gridDim.x / L::SPLIT) {
            mbar_wait(&empty[stage], phase ^ 1);
            const int64_t t = (PASS == 1) ? sample_tile(job, sstride) : job;
            const int64_t row0 = t * T;
            const int rows = (m - row0 < T) ? (int)(m - row0) : T;
            double* dj = smem + (size_t)stage * L::STAGE_DOUBLES;
            double* df = dj + 4 * QS;
            if (rows == T) {
                if (lane == 0)
                    mbar_expect_tx(&full[stage], 4u * q_bytes + (PASS == 2 ? T * 8u : 0u));
                __syncwarp();
                if (lane < 4)
                    bulk_g2s(dj + lane * QS, J + (row0 + lane * QR) * n, q_bytes, &full[stage]);
                else if (lane == 4 && PASS == 2)
                    bulk_g2s(df, f + row0, T * 8u, &full[stage]);
            } else {
                // ragged last tile: plain loads, rows past m are zero
                for (int e = lane; e < T * n; e += 32) {
                    const int r = e / n, c = e % n;
                    dj[(r / QR) * QS + (r % QR) * RS + c] = (r < rows) ? J[(row0 + r) * n + c] : 0.0;
                }
                if (PASS == 2)
                    for (int r = lane; r < T; r += 32) df[r] = (r < rows) ? f[row0 + r] : 0.0;
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[stage]);
            }
            if (++stage == S) { stage = 0; phase ^= 1; }
        }
    } else {
        consumer_dispatch<NB, PASS, EXACT>((blockIdx.x % L::SPLIT) * L::ROLES_CTA + warp % L::ROLES_CTA,
                                           njobs, rinvp_g, n, smem, full, empty, rec,
                                           std::make_integer_sequence<int, L::ROLES>{});
    }
    if (L::KS > 1) {
        // fixed-order sum over the row-split copies
        __syncthreads();
        constexpr int PER = L::PER;
        for (int e = threadIdx.x; e < PER; e += L::THREADS) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < L::KS; k++) s += smem[(size_t)k * PER + e];
            if (e < L::NBLK * 64) {
                const int q = e >> 6, ln = (e >> 1) & 31, c = e & 1;
                int bi = 0, qq = q;
                while (qq >= NB - bi) { qq -= NB - bi; bi++; }
                const int bj = bi + qq;
                rec[(size_t)(8 * bi + (ln >> 2)) * W + 8 * bj + 2 * (ln & 3) + c] = s;
            } else {
                rec[W * W + (e - L::NBLK * 64)] = s;       // Y^T f | f.f, pad | J^T f
            }
        }
    }
}

// ===========================================================================
// Pass 2 for 128 < n <= 256 (NB = 32): a CLUSTER of two CTAs per tile stream.
//
// The upper half of a 256 x 256 Gram matrix is 528 blocks of 8 x 8 = 263 KB of
// FP64 accumulators, more than the register file of one SM: the one-CTA kernel
// above spills by construction (12 KB of spill stores; 221 ms per 12.5 M rows,
// 17 % of the FP64 peak).  Here the two CTAs of a cluster stream the same
// 32-row tiles; each owns half of the Gram blocks (33 per warp, in registers)
// and computes half of Y = J R1^-1, which it writes into BOTH CTAs' shared
// memory (st.shared::cluster through DSMEM), so the triangular product that
// dominates pass 2 is done once per tile, not once per CTA.
//
// Phase A, warp w of CTA c: column blocks j1 = 8 c + w and j2 = 31 - j1 of Y
// for all four 8-row strips of the tile (2 j1 + 2 j2 + 4 = 66 k-chunks for
// every warp: balanced), each R1^-1 fragment loaded ONCE per k-chunk from L2
// (prefetched PF chunks ahead) and used for the four strips.
// Hand-shake per tile, mbarriers in each CTA's shared memory:
//   yfree[0] (CW arrivals)  my warps have finished phase B of the last tile
//   yfree[1] (CW arrivals)  the PEER's warps have (remote arrive)
//   yready   (2 CW)         all 16 warps of the cluster have stored their part
//                           of Y into MY buffer
// ===========================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f64(uint32_t addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITC_%=:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONEC_%=;\n"
        "bra WAITC_%=;\n"
        "DONEC_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// Gram blocks of one role with the fragments taken straight from shared memory
// (base[8 b] = element of column block b for this lane): holding all 32
// fragments of a k-chunk in registers next to 33 accumulator pairs spills.
template <int NB, int Q0, int... Qs>
__device__ __forceinline__ void mma_role_blocks_lazy(double (&acc)[sizeof...(Qs)][2],
                                                     const double* __restrict__ base,
                                                     std::integer_sequence<int, Qs...>) {
    (dmma(acc[Qs], base[8 * Tri<NB>::row(Q0 + Qs)], base[8 * Tri<NB>::col(Q0 + Qs)]), ...);
}

struct Gram2C {
    static constexpr int NB = 32, T = 32, S = 1, CW = 8, W = 256, QR = T / 4;
    static constexpr int QS = QR * W + 4;
    static constexpr int NBLK = Tri<NB>::COUNT;                 // 528
    static constexpr int ROLES = 16;                            // 8 per CTA
    // 33 Gram blocks per warp.  The vector work is spread too: warp w of CTA 0
    // accumulates Y^T f, warp w of CTA 1 J^T f, each for the column blocks
    // b = w, w + 8, w + 16, w + 24 (one warp holding all 2 x 32 partial sums
    // spilled them); f.f by warp 0 of CTA 0.
    static constexpr int BPR = NBLK / ROLES;                    // 33
    static_assert(BPR * ROLES == NBLK, "528 = 16 x 33");
    __host__ __device__ static constexpr int role_begin(int r) { return r * BPR; }
    __host__ __device__ static constexpr int role_count(int r) { return BPR; }
    // NO producer warp: with 9 warps one SM sub-partition hosts three and the
    // register limit is 168 per thread; 33 accumulator pairs are 132 of them
    // and phase A needs ~60 more.  With 8 warps (two per sub-partition) the
    // limit is 255 and nothing spills; warp 0 issues the bulk copies of the
    // NEXT tile at the top of every iteration instead.
    static constexpr int THREADS = CW * 32;
    static constexpr int STAGE_DOUBLES = 4 * QS + T;
    static constexpr int RING_DOUBLES = S * STAGE_DOUBLES;
    // ONE stage for the J tile (it is only read in phase A; the copy of the
    // next tile is issued as soon as all warps are past it and lands during
    // phase B) and TWO Y buffers: phase A of tile t + 1 does not wait for the
    // slowest reader of tile t.
    static constexpr int YBUF_DOUBLES = 2 * 4 * QS;
    // per-warp ring of R1^-1 half fragments (2 per k-chunk: column blocks j2, j1),
    // filled by cp.async RB - 1 k-chunks ahead of their use
    static constexpr int RB = 6;
    static constexpr int BRING_DOUBLES = CW * RB * 2 * 32;
    static constexpr int MAIN_DOUBLES = RING_DOUBLES + YBUF_DOUBLES + BRING_DOUBLES;
    static constexpr int NBAR = 1 + 4 + 1;                      // full | yfree[2][2] | yready
    static constexpr size_t SMEM_BYTES = (size_t)MAIN_DOUBLES * 8 + NBAR * 8 + 16;
    static constexpr int OFF_V1 = W * W, OFF_FF = W * W + W, OFF_V2 = W * W + W + 2;
    static constexpr int REC = W * W + 2 * W + 2;
};
static_assert(Gram2C::SMEM_BYTES <= 232448, "gram2c shared memory");

// warp 0 of each CTA: start the copies of tile `job` into ring stage `stage`
template <bool EXACT>
__device__ __forceinline__ void gram2c_issue(int64_t job, int stage, int64_t m, int n,
                                             const double* __restrict__ J,
                                             const double* __restrict__ f, double* smem,
                                             uint64_t* full) {
    typedef Gram2C L;
    constexpr int T = L::T, QR = L::QR, QS = L::QS, W = L::W;
    const int lane = threadIdx.x & 31;
    const int RS = EXACT ? W : n;
    const int64_t row0 = job * T;
    const int rows = (m - row0 < T) ? (int)(m - row0) : T;
    double* dj = smem + (size_t)stage * L::STAGE_DOUBLES;
    double* df = dj + 4 * QS;
    if (rows == T) {
        const uint32_t q_bytes = (uint32_t)(QR * n) * 8u;
        if (lane == 0) mbar_expect_tx(&full[stage], 4u * q_bytes + T * 8u);
        __syncwarp();
        if (lane < 4)
            bulk_g2s(dj + lane * QS, J + (row0 + lane * QR) * n, q_bytes, &full[stage]);
        else if (lane == 4)
            bulk_g2s(df, f + row0, T * 8u, &full[stage]);
    } else {
        // ragged last tile: plain loads, rows past m are zero
        for (int e = lane; e < T * n; e += 32) {
            const int r = e / n, c = e % n;
            dj[(r / QR) * QS + (r % QR) * RS + c] = (r < rows) ? J[(row0 + r) * n + c] : 0.0;
        }
        for (int r = lane; r < T; r += 32) df[r] = (r < rows) ? f[row0 + r] : 0.0;
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[stage]);
    }
    __syncwarp();
}

// phase B of one role: its 33 Gram blocks (static block list, accumulators in
// the shared `acc` array) from the Y tile in shared memory
template <int ROLE>
__device__ __forceinline__ void gram2c_blocks(double (&acc)[Gram2C::BPR][2],
                                              const double* __restrict__ base) {
    mma_role_blocks_lazy<Gram2C::NB, ROLE * Gram2C::BPR>(
        acc, base, std::make_integer_sequence<int, Gram2C::BPR>{});
}
template <int... ROLES_>
__device__ __forceinline__ void gram2c_blocks_dispatch(int role, double (&acc)[Gram2C::BPR][2],
                                                       const double* __restrict__ base,
                                                       std::integer_sequence<int, ROLES_...>) {
    ((role == ROLES_ ? (gram2c_blocks<ROLES_>(acc, base), 0) : 0), ...);
}
template <int ROLE>
__device__ __forceinline__ void gram2c_store(const double (&acc)[Gram2C::BPR][2],
                                             double* __restrict__ rec, int lr, int lc) {
    constexpr int NB = Gram2C::NB, W = Gram2C::W, Q0 = ROLE * Gram2C::BPR;
#pragma unroll
    for (int q = 0; q < Gram2C::BPR; q++) {
        const int bi = Tri<NB>::row(Q0 + q), bj = Tri<NB>::col(Q0 + q);
        double* dst = rec + (size_t)(8 * bi + lr) * W + 8 * bj + 2 * lc;
        *reinterpret_cast<double2*>(dst) = make_double2(acc[q][0], acc[q][1]);
    }
}
template <int... ROLES_>
__device__ __forceinline__ void gram2c_store_dispatch(int role,
                                                      const double (&acc)[Gram2C::BPR][2],
                                                      double* __restrict__ rec, int lr, int lc,
                                                      std::integer_sequence<int, ROLES_...>) {
    ((role == ROLES_ ? (gram2c_store<ROLES_>(acc, rec, lr, lc), 0) : 0), ...);
}

// The tile loop of one consumer warp.  Everything except the Gram-block list of
// phase B is the SAME code for all sixteen roles (one copy in the instruction
// cache: with a body per role the warps stalled on instruction fetch, 4.2 of
// 15 stall cycles per issue in the first version).
template <bool EXACT>
__device__ __forceinline__ void gram2c_consumer(int role, int64_t njobs, int64_t job0,
                                                int64_t jstride, int64_t m,
                                                const double* __restrict__ J,
                                                const double* __restrict__ f,
                                                const double* __restrict__ rinvp, int n,
                                                double* smem, uint64_t* full, uint64_t* yfree,
                                                uint64_t* yready, double* __restrict__ rec,
                                                uint32_t crank) {
    typedef Gram2C L;
    constexpr int NB = L::NB, W = L::W, QR = L::QR, QS = L::QS, RB = L::RB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lr = lane >> 2, lc = lane & 3;
    const int RS = EXACT ? W : n;
    double* ybuf0 = smem + L::RING_DOUBLES;
    const double* bring = smem + L::RING_DOUBLES + L::YBUF_DOUBLES + warp * (RB * 2 * 32) + lane;
    const uint32_t bring_s = smem_u32(bring);
    const uint32_t peer = crank ^ 1u;
    const uint32_t ybuf0_peer = map_to_cta(smem_u32(ybuf0), peer);
    const uint32_t yready_peer = map_to_cta(smem_u32(yready), peer);
    const uint32_t yready_self = map_to_cta(smem_u32(yready), crank);
    // yfree[2 b]: my readers are done with my Y buffer b; yfree[2 b + 1]: the
    // peer's readers with theirs (they arrive here remotely)
    const uint32_t yfree_self = map_to_cta(smem_u32(yfree), crank);
    const uint32_t yfree_peer = map_to_cta(smem_u32(yfree), peer);
    // this warp's two column blocks of Y
    const int j1 = (int)crank * 8 + warp, j2 = NB - 1 - j1;
    const int kmax = 2 * j2 + 2, k1max = 2 * j1 + 2;
    const bool cta0 = crank == 0;
    const double* tj = smem;
    const double* tf = tj + 4 * QS;

    double acc[L::BPR][2];
#pragma unroll
    for (int q = 0; q < L::BPR; q++) { acc[q][0] = 0.0; acc[q][1] = 0.0; }
    // Y^T f of my two column blocks (columns 8 j + 2 lc + {0, 1}, partial over
    // the rows this lane holds), J^T f of column blocks warp + 8 i (CTA 1), f.f
    double gy1[2] = {0.0, 0.0}, gy2[2] = {0.0, 0.0}, gj[4] = {0.0, 0.0, 0.0, 0.0}, ff = 0.0;
    // R1^-1 half fragment of block (kc >> 1, j), k-half kc & 1, for this lane
    auto rin = [&](int kc, int j) -> const double* {
        const int kb = kc >> 1;
        const int q = kb * NB - (kb * (kb - 1)) / 2 + (j - kb);
        return rinvp + (size_t)(q * 2 + (kc & 1)) * 32 + lane;
    };

    uint32_t phase = 0;                    // of `full` and `yready`: one completion per tile
    uint32_t fphase[2] = {0u, 0u};         // of the yfree pair of each Y buffer
    int yb = 0;
    if (warp == 0 && job0 < njobs) gram2c_issue<EXACT>(job0, 0, m, n, J, f, smem, full);
    for (int64_t t = job0; t < njobs; t += jstride) {
        mbar_wait(full, phase);
        // ---- phase A: my two column blocks of Y for the four strips ----
        double y1[4][2], y2[4][2];
#pragma unroll
        for (int s = 0; s < 4; s++) { y1[s][0] = y1[s][1] = y2[s][0] = y2[s][1] = 0.0; }
        const double* arow = tj + (lr & 3) * QS + (lr >> 2) * RS + lc;   // strip s: + 2 s RS
        // R1^-1 fragments: every lane copies its own element of the two half
        // fragments of k-chunk kn into ring slot kn % RB (and reads only that
        // element back: no cross-lane hand-over), one cp.async group per k-chunk
        auto fetch = [&](int kn) {
            const int slot = kn % RB;
            if (kn < kmax)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(
                                 bring_s + (uint32_t)(slot * 2 * 32) * 8u),
                             "l"(rin(kn, j2))
                             : "memory");
            if (kn < k1max)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(
                                 bring_s + (uint32_t)((slot * 2 + 1) * 32) * 8u),
                             "l"(rin(kn, j1))
                             : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
#pragma unroll
        for (int u = 0; u < RB - 1; u++) fetch(u);
#pragma unroll 2
        for (int kc = 0; kc < kmax; kc++) {
            // groups 0 .. kc + RB - 2 are committed: chunk kc has landed when at
            // most RB - 2 of them are pending
            asm volatile("cp.async.wait_group %0;" ::"n"(RB - 2) : "memory");
            const int slot = kc % RB;
            const double bb2 = bring[slot * 2 * 32];
            const double bb1 = (kc < k1max) ? bring[(slot * 2 + 1) * 32] : 0.0;
            const bool in = EXACT || kc * 4 + lc < n;
            double a[4];
#pragma unroll
            for (int s = 0; s < 4; s++) a[s] = in ? arow[2 * s * RS + kc * 4] : 0.0;
            // refill the slot of chunk kc - 1 (its values were consumed by the
            // DMMAs of the last iteration)
            fetch(kc + RB - 1);
#pragma unroll
            for (int s = 0; s < 4; s++) dmma(y2[s], a[s], bb2);
            if (kc < k1max) {
#pragma unroll
                for (int s = 0; s < 4; s++) dmma(y1[s], a[s], bb1);
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        // Y^T f from the fragments in registers: y[s] holds row (quarter lr & 3,
        // 2 s + (lr >> 2)) of the tile
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const double fr = tf[(lr & 3) * QR + 2 * s + (lr >> 2)];
            gy1[0] = fma(y1[s][0], fr, gy1[0]);
            gy1[1] = fma(y1[s][1], fr, gy1[1]);
            gy2[0] = fma(y2[s][0], fr, gy2[0]);
            gy2[1] = fma(y2[s][1], fr, gy2[1]);
        }
        if (!cta0) {
            // J^T f of my column blocks straight from the raw tile (trf.py:244)
#pragma unroll 1
            for (int kc = 0; kc < QR; kc++) {
                const double fk = tf[lc * QR + kc];
                const double* rawb = tj + lc * QS + kc * RS + lr;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int b = warp + 8 * i;
                    const double raw = (EXACT || 8 * b + lr < n) ? rawb[8 * b] : 0.0;
                    gj[i] = fma(raw, fk, gj[i]);
                }
            }
        } else if (warp == 0) {
            const double fk = tf[lane];                  // T = 32 rows
            ff = fma(fk, fk, ff);
        }
        // the readers (both CTAs) of the tile that used this Y buffer last are done
        double* ybuf = ybuf0 + yb * 4 * QS;
        const uint32_t ybuf_peer = ybuf0_peer + (uint32_t)(yb * 4 * QS) * 8u;
        mbar_wait_cluster(&yfree[2 * yb], fphase[yb] ^ 1);
        mbar_wait_cluster(&yfree[2 * yb + 1], fphase[yb] ^ 1);
        fphase[yb] ^= 1u;
        {
            const int odd = lr & 1;
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const int off = (lr & 3) * QS + (2 * s + (lr >> 2)) * W + 2 * lc;
                double* yl = ybuf + off;
                const uint32_t yp = ybuf_peer + (uint32_t)off * 8u;
                // element order swapped on odd rows (bank conflicts, see above)
                const double v1a = odd ? y1[s][1] : y1[s][0], v1b = odd ? y1[s][0] : y1[s][1];
                const double v2a = odd ? y2[s][1] : y2[s][0], v2b = odd ? y2[s][0] : y2[s][1];
                yl[8 * j1 + odd] = v1a;
                yl[8 * j1 + 1 - odd] = v1b;
                yl[8 * j2 + odd] = v2a;
                yl[8 * j2 + 1 - odd] = v2b;
                st_cluster_f64(yp + (uint32_t)(8 * j1 + odd) * 8u, v1a);
                st_cluster_f64(yp + (uint32_t)(8 * j1 + 1 - odd) * 8u, v1b);
                st_cluster_f64(yp + (uint32_t)(8 * j2 + odd) * 8u, v2a);
                st_cluster_f64(yp + (uint32_t)(8 * j2 + 1 - odd) * 8u, v2b);
            }
        }
        // the lanes' stores are ordered before lane 0's release-arrive by the
        // warp barrier (cumulativity); no separate cluster fence
        __syncwarp();
        if (lane == 0) {
            mbar_arrive_cluster(yready_self);
            mbar_arrive_cluster(yready_peer);
        }
        mbar_wait_cluster(yready, phase);
        // all sixteen warps are past phase A: the J stage is free
        if (warp == 0 && t + jstride < njobs)
            gram2c_issue<EXACT>(t + jstride, 0, m, n, J, f, smem, full);
        // ---- phase B: my Gram blocks += Y^T Y (block list by role) ----
#pragma unroll 1
        for (int kc = 0; kc < QR; kc++) {
            const double* base = ybuf + lc * QS + kc * W + lr;
            gram2c_blocks_dispatch(role, acc, base, std::make_integer_sequence<int, L::ROLES>{});
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive_cluster(yfree_self + (uint32_t)(2 * yb) * 8u);        // my yfree[2 yb]
            mbar_arrive_cluster(yfree_peer + (uint32_t)(2 * yb + 1) * 8u);    // peer's yfree[2 yb + 1]
        }
        phase ^= 1u;
        yb ^= 1;
    }
    // ---- this CTA's record: its half of the Gram blocks, its vector parts ----
    gram2c_store_dispatch(role, acc, rec, lr, lc, std::make_integer_sequence<int, L::ROLES>{});
    // Y^T f: sum over the rows (lr), lanes lr == 0 hold columns 8 j + 2 lc + {0, 1}
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) {
        gy1[0] += __shfl_xor_sync(0xffffffffu, gy1[0], off);
        gy1[1] += __shfl_xor_sync(0xffffffffu, gy1[1], off);
        gy2[0] += __shfl_xor_sync(0xffffffffu, gy2[0], off);
        gy2[1] += __shfl_xor_sync(0xffffffffu, gy2[1], off);
    }
    if (lr == 0) {
        rec[L::OFF_V1 + 8 * j1 + 2 * lc] = gy1[0];
        rec[L::OFF_V1 + 8 * j1 + 2 * lc + 1] = gy1[1];
        rec[L::OFF_V1 + 8 * j2 + 2 * lc] = gy2[0];
        rec[L::OFF_V1 + 8 * j2 + 2 * lc + 1] = gy2[1];
    }
    if (!cta0) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            gj[i] += __shfl_xor_sync(0xffffffffu, gj[i], 1);
            gj[i] += __shfl_xor_sync(0xffffffffu, gj[i], 2);
        }
        if (lc == 0) {
#pragma unroll
            for (int i = 0; i < 4; i++) rec[L::OFF_V2 + 8 * (warp + 8 * i) + lr] = gj[i];
        }
    } else if (warp == 0) {
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) ff += __shfl_xor_sync(0xffffffffu, ff, off);
        if (lane == 0) rec[L::OFF_FF] = ff;
    }
}

template <bool EXACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Gram2C::THREADS, 1)
gram2c_kernel(int64_t m, int n, const double* __restrict__ J, const double* __restrict__ f,
              const double* __restrict__ rinvp, double* __restrict__ partial) {
    typedef Gram2C L;
    constexpr int T = L::T, CW = L::CW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* smem = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::MAIN_DOUBLES);
    uint64_t* yfree = full + 1;           // [2 b] mine, [2 b + 1] the peer's readers, b = 0, 1
    uint64_t* yready = yfree + 4;
    const int warp = threadIdx.x >> 5;
    const uint32_t crank = cluster_ctarank();
    double* rec = partial + (size_t)blockIdx.x * L::REC;

    for (int i = threadIdx.x; i < L::MAIN_DOUBLES; i += L::THREADS) smem[i] = 0.0;
    for (int i = threadIdx.x; i < L::REC; i += L::THREADS) rec[i] = 0.0;
    if (threadIdx.x == 0) {
        mbar_init(full, 1);
        for (int b = 0; b < 4; b++) mbar_init(&yfree[b], CW);
        mbar_init(yready, 2 * CW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                   // the peer's barriers exist before any remote arrive

    const int64_t njobs = (m + T - 1) / T;
    const int64_t job0 = blockIdx.x >> 1, jstride = gridDim.x >> 1;
    gram2c_consumer<EXACT>((int)crank * CW + warp, njobs, job0, jstride, m, J, f, rinvp, n, smem,
                           full, yfree, yready, rec, crank);
    // neither CTA may exit while the other can still write into its shared memory
    __syncthreads();
    cluster_sync_all();
}

// out (blsq_tall_record_size(n) doubles): G (n x n row-major, upper triangle
// valid) | Y^T f (n) | f.f | J^T f (n)  = sum over the P per-CTA records, in
// CTA order.
__global__ void gram_reduce_kernel(int P, int W, int REC, int n,
                                   const double* __restrict__ partial,
                                   double* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = n * n + 2 * n + 1;
    if (e >= total) return;
    int src;
    if (e < n * n) {
        const int r = e / n, c = e % n;
        if (c < r) { out[e] = 0.0; return; }
        src = r * W + c;
    } else if (e < n * n + n) {
        src = W * W + (e - n * n);
    } else if (e == n * n + n) {
        src = W * W + W;
    } else {
        src = W * W + W + 2 + (e - n * n - n - 1);
    }
    double s = 0.0;
    for (int p = 0; p < P; p++) s += partial[(size_t)p * REC + src];
    out[e] = s;
}

// ---- ||f||^2 of a trial evaluation (trf.py:311, dogbox.py:224) -------------
__global__ void __launch_bounds__(256) sumsq_partial_kernel(int64_t m, const double* __restrict__ f,
                                                            double* __restrict__ part) {
    __shared__ double sh[8];
    double s = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const double v = f[i];
        s = fma(v, v, s);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += sh[w];
        part[blockIdx.x] = t;
    }
}
__global__ void sumsq_final_kernel(int P, const double* __restrict__ part,
                                   double* __restrict__ out) {
    // one warp, fixed order
    double s = 0.0;
    for (int p = threadIdx.x; p < P; p += 32) s += part[p];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (threadIdx.x == 0) out[0] = s;
}

int sm_count() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 148;
    return sms > 0 ? sms : 148;
}

using blsq_tall::nb_for;

template <int NB, int PASS, bool EXACT>
int launch_gram_e(int64_t m, int n, const double* J, const double* f, const double* rinvp,
                  int sstride, double* work, double* out, cudaStream_t s) {
    typedef GramLayout<NB, PASS> L;
    typedef GramCfg<NB> C;
    const int sms = sm_count();
    int64_t ntiles = (m + C::T - 1) / C::T;
    if (PASS == 1 && sstride > 1) ntiles /= sstride;
    int64_t want = (ntiles > 0 ? ntiles : 1) * L::SPLIT;       // SPLIT CTAs per tile stream
    int grid = (int)(want < sms ? want : sms);
    grid -= grid % L::SPLIT;
    if (grid < L::SPLIT) grid = L::SPLIT;
    cudaError_t e = cudaFuncSetAttribute(gram_kernel<NB, PASS, EXACT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)L::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    gram_kernel<NB, PASS, EXACT><<<grid, L::THREADS, L::SMEM_BYTES, s>>>(m, n, J, f, rinvp, sstride,
                                                                         work);
    BLSQ_LAUNCH_CHECK();
    const int total = n * n + 2 * n + 1;
    gram_reduce_kernel<<<(total + 255) / 256, 256, 0, s>>>(grid, L::W, L::REC, n, work, out);
    BLSQ_LAUNCH_CHECK();
    return 0;
}

template <bool EXACT>
int launch_gram2c(int64_t m, int n, const double* J, const double* f, const double* rinvp,
                  double* work, double* out, cudaStream_t s) {
    typedef Gram2C L;
    cudaError_t e = cudaFuncSetAttribute(gram2c_kernel<EXACT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)L::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    // clusters that can be resident at once (a GPC with an odd number of free
    // SMs leaves one without a partner): persistent CTAs, so no more than that
    static int max_clusters = 0;
    if (max_clusters == 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * 74, 1, 1);
        cfg.blockDim = dim3(L::THREADS, 1, 1);
        cfg.dynamicSmemBytes = L::SMEM_BYTES;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, gram2c_kernel<EXACT>, &cfg) != cudaSuccess || nc < 1)
            nc = sm_count() / 2 - 4;
        (void)cudaGetLastError();
        max_clusters = nc;
        if (getenv("BLSQ_DEBUG")) fprintf(stderr, "# gram2c: max active clusters %d\n", nc);
    }
    const int64_t ntiles = (m + L::T - 1) / L::T;
    int clusters = (int)(ntiles < max_clusters ? (ntiles > 0 ? ntiles : 1) : max_clusters);
    const int grid = 2 * clusters;
    gram2c_kernel<EXACT><<<grid, L::THREADS, L::SMEM_BYTES, s>>>(m, n, J, f, rinvp, work);
    BLSQ_LAUNCH_CHECK();
    const int total = n * n + 2 * n + 1;
    gram_reduce_kernel<<<(total + 255) / 256, 256, 0, s>>>(grid, L::W, L::REC, n, work, out);
    BLSQ_LAUNCH_CHECK();
    return 0;
}

template <int NB, int PASS>
int launch_gram(int64_t m, int n, const double* J, const double* f, const double* rinvp,
                int sstride, double* work, double* out, cudaStream_t s) {
#ifndef BLSQ_NO_GRAM2C
    if (NB == 32 && PASS == 2) {
        if (n == 256) return launch_gram2c<true>(m, n, J, f, rinvp, work, out, s);
        return launch_gram2c<false>(m, n, J, f, rinvp, work, out, s);
    }
#endif
    if (n == 8 * NB)
        return launch_gram_e<NB, PASS, true>(m, n, J, f, rinvp, sstride, work, out, s);
    return launch_gram_e<NB, PASS, false>(m, n, J, f, rinvp, sstride, work, out, s);
}

}  // namespace

extern "C" {

int64_t blsq_tall_gram_work_size(int n) {
    if (n < 2 || n > 256) return BLSQ_E_UNSUPPORTED;
    const int W = 8 * nb_for(n);
    return (int64_t)sm_count() * (W * W + 2 * W + 2);
}

int64_t blsq_tall_record_size(int n) {
    if (n < 2 || n > 256) return BLSQ_E_UNSUPPORTED;
    return ((int64_t)n * n + 2 * n + 1 + 1) & ~(int64_t)1;
}

/* pass 1 may run on one tile out of `stride`: keep >= 64 n^2 sampled rows so
 * that cond(J R1^-1) stays close to 1 (entries of the sampled Gram deviate by
 * ~1/sqrt(rows)); never sample small problems */
int blsq_tall_sample_stride(int64_t m, int n) {
    if (n < 2 || n > 256 || m < 0) return 1;
    int64_t s = m / (64 * (int64_t)n * n);
    return (int)(s < 1 ? 1 : (s > 8 ? 8 : s));
}

int blsq_tall_gram(int pass, int64_t m, int n, const double* J, const double* f,
                   const double* Rinv, int sstride, double* work, double* out, void* stream) {
    if (m < 0 || !J || !work || !out) return BLSQ_E_BADARG;
    if (pass != 1 && pass != 2) return BLSQ_E_BADARG;
    if (pass == 2 && (!Rinv || !f)) return BLSQ_E_BADARG;
    if (pass == 2 || sstride < 1) sstride = 1;
    if (!f) f = J;                                  /* pass 1 never reads f */
    // rows are moved by 16-byte-granular bulk copies: a quarter tile is 16
    // rows = 128 n bytes, so any n keeps source and size 16-byte aligned
    if (n < 2 || n > 256) return BLSQ_E_UNSUPPORTED;
    if (((uintptr_t)J & 15) || ((uintptr_t)f & 15)) return BLSQ_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
#define BLSQ_GRAM_CASE(NB_)                                                         \
    case NB_:                                                                       \
        return pass == 1 ? launch_gram<NB_, 1>(m, n, J, f, Rinv, sstride, work, out, s) \
                         : launch_gram<NB_, 2>(m, n, J, f, Rinv, sstride, work, out, s);
    switch (nb_for(n)) {
        BLSQ_GRAM_CASE(2)
        BLSQ_GRAM_CASE(4)
        BLSQ_GRAM_CASE(8)
        BLSQ_GRAM_CASE(16)
        BLSQ_GRAM_CASE(32)
    }
#undef BLSQ_GRAM_CASE
    return BLSQ_E_UNSUPPORTED;
}

int blsq_tall_sumsq(int64_t m, const double* f, double* work, double* out, void* stream) {
    if (m < 0 || !f || !work || !out) return BLSQ_E_BADARG;
    const int sms = sm_count();
    int64_t want = (m + 256 * 8 - 1) / (256 * 8);
    int grid = (int)(want < 1 ? 1 : (want > 4 * sms ? 4 * sms : want));
    sumsq_partial_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(m, f, work);
    BLSQ_LAUNCH_CHECK();
    sumsq_final_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(grid, work, out);
    BLSQ_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
