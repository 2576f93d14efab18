// Shared bookkeeping of the tall-mode kernels (blsq_tall_gram.cu,
// blsq_tall_factor.cu, blsq_tall_round.cu): block counts and the layout of the
// factor record `fac` (see include/blsq.h, "tall mode").
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BLSQ_TALL_HD __host__ __device__
#else
#define BLSQ_TALL_HD
#endif

namespace blsq_tall {

// number of 8-column blocks the Gram kernels are instantiated for
inline BLSQ_TALL_HD int nb_for(int n) {
    if (n <= 16) return 2;
    if (n <= 32) return 4;
    if (n <= 64) return 8;
    if (n <= 128) return 16;
    return 32;
}

// fac (doubles):  R1 | R | scratch (n*n each) | Q^T f (n) | g (n) | f.f | info |
//                 shift1, shift2 | refine |
//                 R1^-1 in DMMA-fragment order (64 doubles per upper 8x8 block)
struct FacLayout {
    int n, nb;
    int64_t n2, R1, R, SCR, QTF, G, OBJ, INFO, SHIFT, REFINE, RINVP, SIZE;
    BLSQ_TALL_HD explicit FacLayout(int n_) : n(n_), nb(nb_for(n_)) {
        n2 = (int64_t)n * n;
        R1 = 0;
        R = n2;
        SCR = 2 * n2;
        QTF = 3 * n2;
        G = QTF + n;
        OBJ = G + n;
        INFO = OBJ + 1;
        SHIFT = INFO + 1;                 // diagonal shifts used by pass 1, 2
        REFINE = SHIFT + 2;               // 1: cond(J R1^-1) too large, run another pass
        RINVP = (REFINE + 1 + 1) & ~(int64_t)1;
        SIZE = RINVP + (int64_t)nb * (nb + 1) / 2 * 64;
    }
};

// index of block (i, j), i <= j < NB, in the row-by-row upper-triangle order
inline BLSQ_TALL_HD int tri_block(int nb, int i, int j) {
    return i * nb - (i * (i - 1)) / 2 + (j - i);
}

}  // namespace blsq_tall
