// score_kernels.cuh -- sm_100a kernel for block scoring: the reference's mafScoreRange (mz_scores.c:124-152),
// the sum-of-pairs score of an alignment block over a column range.  SURVEY 8(f) rank 1.
//
// What is computed is the reference's
//     sum_{i in range} sum_{rows r1 < r2}  SS(text1[i], text2[i]) - (i > 0 ? GAP2(text1[i-1], text2[i-1], text1[i], text2[i]) : 0)
// an O(rows^2 * columns) loop of table lookups.  HOW is new: ss[][] only distinguishes six character classes
// (mz_scores.c:39-54) and gop[] only dash / non-dash (mz_scores.c:57-79), and the six charged gop patterns are
// symmetric under swapping the two rows, so a column's contribution is a quadratic form of ten counts:
//     n_k   rows of class k in the column (A, C, G, T, other, '-')
//     t_pq  rows whose previous column is dash (p) / non-dash and whose own column is dash (q) / non-dash
//     sum_{k} S6[k][k] * n_k (n_k - 1) / 2  +  sum_{k<l} S6[k][l] * n_k n_l  -  gap_open * (t00 t01 + t01 t10 + t10 t11)
// (0001/0010 -> {00,01}, 0110/1001 -> {01,10}, 1101/1110 -> {10,11}: the six entries init_scores sets).  The work
// is O(rows * columns) byte classification -- memory-shaped, four columns per 32-bit load, counted in byte lanes.
// All arithmetic is integer; the reference accumulates integer-valued doubles, which is exact below 2^53, so the
// int64 sum converts to the same double.
#pragma once
#include "yama_kernels.cuh"

namespace yb {

// One scored block range.  Text layout in the blob: row j at off + j*pitch; a row is 4 lead bytes (the last one
// holds the column before the range, when there is one), then `size` text bytes, zero-padded to a multiple of 4.
struct ScoreMeta {
    unsigned long long off;     // byte offset of row 0 (4-aligned)
    int nrows, size;            // components, scored columns
    int pitch;                  // bytes between rows (multiple of 4)
    int firstGap;               // 1: the range does not start at text column 0, its first column is charged GAP2 too
};
struct ScoreUnit { int block, col0; };      // one warp: columns col0 .. col0+127 of a block

constexpr int SCORE_THREADS = 128;
constexpr int SCORE_UNIT_COLS = 128;
constexpr int SCORE_ROWS = 8;              // rows in flight per lane

template <typename T>
__device__ __forceinline__ T column_score(const ScoreConst &c_sc, const int n[6], int t01, int t10, int t11, int nrows, bool gap) {
    T s = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        s += (T)c_sc.S6[k][k] * (((T)n[k] * (T)(n[k] - 1)) / 2);
#pragma unroll
        for (int l = k + 1; l < 6; ++l) s += (T)c_sc.S6[k][l] * ((T)n[k] * (T)n[l]);
    }
    if (gap) {
        const T a00 = (T)(nrows - t01 - t10 - t11), a01 = (T)t01, a10 = (T)t10, a11 = (T)t11;
        s -= (T)c_sc.gap_open * (a00 * a01 + a01 * a10 + a10 * a11);
    }
    return s;
}

__global__ void __launch_bounds__(SCORE_THREADS)
yb_score_kernel(const ScoreMeta *__restrict__ metas, const ScoreUnit *__restrict__ units, int nUnits,
                const unsigned char *__restrict__ blob, unsigned long long *__restrict__ sums, int rows32,
                const __grid_constant__ ScoreConst c_sc) {
    const int lane = threadIdx.x & 31;
    const int u = blockIdx.x * (SCORE_THREADS / 32) + (threadIdx.x >> 5);
    if (u >= nUnits) return;
    const ScoreUnit un = units[u];
    const ScoreMeta bm = metas[un.block];
    const int c0 = un.col0 + 4 * lane;                       // this lane's four columns c0..c0+3 of the range
    const bool live = c0 < bm.size;
    const unsigned char *p = blob + bm.off + 4 + (live ? c0 : 0);
    const int pitch = bm.pitch;

    int wide[7][4];                                          // nA nC nG nT t01 t11 t10 per column, beyond 255 rows
#pragma unroll
    for (int k = 0; k < 7; ++k)
#pragma unroll
        for (int b = 0; b < 4; ++b) wide[k][b] = 0;

    for (int j0 = 0; j0 < bm.nrows; j0 += 255) {
        const int j1 = min(bm.nrows, j0 + 255);
        unsigned cA = 0, cC = 0, cG = 0, cT = 0, c01 = 0, c11 = 0, c10 = 0;      // byte lane b counts column c0+b
        // SCORE_ROWS rows per trip: all their loads are issued before the first one is consumed (the kernel is
        // bound by bytes in flight, and loads do not move across the warp shuffle below on their own)
        for (int j = j0; j < j1; j += SCORE_ROWS) {
            unsigned wv[SCORE_ROWS], lead[SCORE_ROWS];
#pragma unroll
            for (int k = 0; k < SCORE_ROWS; ++k) {
                const bool in = j + k < j1;
                wv[k] = (live && in) ? __ldg(reinterpret_cast<const unsigned *>(p + (size_t)k * pitch)) : 0u;
                lead[k] = (lane == 0 && in) ? __ldg(reinterpret_cast<const unsigned *>(p + (size_t)k * pitch - 4)) : 0u;
            }
            p += (size_t)min(SCORE_ROWS, j1 - j) * pitch;
#pragma unroll
            for (int k = 0; k < SCORE_ROWS; ++k) {
                if (j + k >= j1) break;
                const unsigned w = wv[k];
                unsigned pw = __shfl_up_sync(0xffffffffu, w, 1);
                if (lane == 0) pw = lead[k];                                        // the lead word, or the unit before
                const unsigned lw = w | 0x20202020u;
                const unsigned mD = eq_bytes80(w, 0x2d2d2d2du);
                const unsigned pD = __funnelshift_l(eq_bytes80(pw, 0x2d2d2d2du), mD, 8);   // dash mask of the column before
                cA += eq_bytes80(lw, 0x61616161u) >> 7; cC += eq_bytes80(lw, 0x63636363u) >> 7;
                cG += eq_bytes80(lw, 0x67676767u) >> 7; cT += eq_bytes80(lw, 0x74747474u) >> 7;
                c01 += (mD & ~pD) >> 7; c11 += (mD & pD) >> 7; c10 += (pD & ~mD) >> 7;
            }
        }
        const unsigned acc[7] = {cA, cC, cG, cT, c01, c11, c10};
#pragma unroll
        for (int k = 0; k < 7; ++k)
#pragma unroll
            for (int b = 0; b < 4; ++b) wide[k][b] += (int)((acc[k] >> (8 * b)) & 0xffu);
    }

    long long total = 0;
    if (live) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (c0 + b >= bm.size) break;
            const int nD = wide[4][b] + wide[5][b];
            int n[6] = {wide[0][b], wide[1][b], wide[2][b], wide[3][b], 0, nD};
            n[4] = bm.nrows - n[0] - n[1] - n[2] - n[3] - nD;
            const bool gap = (c0 + b > 0) || bm.firstGap;                        // mz_scores.c:143 (i > 0)
            // 32-bit products while |column| <= (max|S6| + gap_open) * rows^2 / 2 fits (rows32, from the host)
            total += bm.nrows <= rows32 ? (long long)column_score<int>(c_sc, n, wide[4][b], wide[6][b], wide[5][b], bm.nrows, gap)
                                      : column_score<long long>(c_sc, n, wide[4][b], wide[6][b], wide[5][b], bm.nrows, gap);
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
    if (lane == 0) atomicAdd(sums + un.block, (unsigned long long)total);
}

}  // namespace yb
