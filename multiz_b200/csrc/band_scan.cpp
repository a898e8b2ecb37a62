// band_scan.cpp -- host-side scan of one pair's band, written branch-free so that the compiler vectorises it.
// Compiled twice by the Makefile: once with -mavx2 (-DYB_BAND_SCAN_NAME=yb_band_scan_avx2) and once for the
// x86-64 baseline (yb_band_scan_generic, which also holds the run-time dispatcher yb_band_scan).
//
// One pass computes what mz_yama.c:58-71 validates (LB[0]==0, RB[M]==N, width >= min(N,10), LB and RB
// non-decreasing), the cell count `tback_size` of mz_yama.c:60-66, the widest row, and the wavefront schedule
// of the fill kernel (see make_schedule in yama_b200.cu).  A violation only sets a flag: the caller re-runs the
// scalar check, which words the message exactly as the reference does.
#include <cstdint>
#include <cstdlib>

#ifndef YB_BAND_SCAN_NAME
#error "compile with -DYB_BAND_SCAN_NAME=..."
#endif

// lanesOf(wmax, M) -> wavefront width B (32, 128, 256 ...) the pair will run with; the schedule depends on it
extern "C" int64_t YB_BAND_SCAN_NAME(int M, int N, const int32_t *__restrict__ LB, const int32_t *__restrict__ RB,
                                     int32_t *wmax, int32_t *__restrict__ sched, int32_t *nSteps,
                                     int (*lanesOf)(int, int), int32_t *lanes, int32_t *connected) {
    const int need = N < 10 ? N : 10;
    int bad = (LB[0] != 0) | (RB[M] != N);
    int64_t cells = 0;
    int wm = 0;
    {
        int badw = 0, sum = 0, r = 0;
        // int32 partial sums cannot overflow within 4096 rows of width < 2^19; flush into the int64 total
        while (r <= M) {
            const int end = (r + 4096 <= M + 1) ? r + 4096 : M + 1;
            sum = 0;
            for (; r < end; ++r) {
                const int w = RB[r] - LB[r];
                badw |= (w < need);
                sum += w + 1;
                wm = w > wm ? w : wm;
            }
            cells += sum;
            if (wm >= (1 << 19)) bad |= 2;      // beyond any kernel limit; the caller rejects it anyway
        }
        bad |= badw;
    }
    {
        // (connected: every row can be reached from the row above -- LB[r] <= RB[r-1] + 1; the reference does not ask for
        //  it and pre_yama's bands always are, but the fill variant without existence multipliers relies on it)
        int badm = 0, gap = 0;
        for (int r = 1; r <= M; ++r) {
            badm |= (LB[r] < LB[r - 1]) | (RB[r] < RB[r - 1]);
            gap |= (LB[r] > RB[r - 1] + 1);
        }
        bad |= badm;
        if (connected) *connected = !gap;
    }
    if (bad) return -1;
    *wmax = wm + 1;
    // schedule: rows Bb+1..Bb+B run on lanes 0..B-1 with column = step - (OFF_b + lane)
    const int B = lanesOf(wm + 1, M);
    *lanes = B;
    if (B <= 0) { *nSteps = 0; return cells; }             // wider than any kernel: the caller reports the limit
    int off = 0, steps = 8;
    const int nblk = (M + B - 1) / B;
    for (int b = 0; b < nblk; ++b) {
        if (sched) sched[b] = off;
        if (b == nblk - 1) {
            const int last = off + ((M - 1) % B) + RB[M];
            steps = ((last + 2) + 7) & ~7;
            break;
        }
        int nd = B;
        const int r0 = B * b + 1;
        const int r1 = (M - B < B * b + B) ? M - B : B * b + B;
        for (int r = r0; r <= r1; ++r) {
            const int v = RB[r + 1] - LB[r + B] + 3;
            nd = v > nd ? v : nd;
        }
        off += nd;
    }
    *nSteps = steps;
    return cells;
}

// Delta-coded band of one pair for the host->device copy: byte r = X[r] - X[r-1] when that lies in 0..254 (bands are
// non-decreasing and move a few columns per row), 255 marks a row whose step the caller lists separately; byte 0 is 0.
// Written so that the compiler vectorises it (32-bit subtract, compare, narrowing store).  Returns the number of marked rows.
#define YB_CAT2(a, b) a##b
#define YB_CAT(a, b) YB_CAT2(a, b)
#define YB_BAND_PACK_NAME YB_CAT(YB_BAND_SCAN_NAME, _pack)
extern "C" int YB_BAND_PACK_NAME(int M, const int32_t *__restrict__ X, uint8_t *__restrict__ out) {
    int marked = 0;
    out[0] = 0;
    for (int r = 1; r <= M; ++r) {
        const uint32_t d = (uint32_t)X[r] - (uint32_t)X[r - 1];
        const bool ok = d < 255u;
        out[r] = ok ? (uint8_t)d : (uint8_t)255;
        marked += !ok;
    }
    return marked;
}

#ifdef YB_BAND_SCAN_DISPATCH
extern "C" int yb_band_scan_avx2_pack(int, const int32_t *, uint8_t *);
extern "C" int yb_band_pack(int M, const int32_t *X, uint8_t *out) {
    static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("YB_NO_AVX2");
    return avx2 ? yb_band_scan_avx2_pack(M, X, out) : YB_BAND_PACK_NAME(M, X, out);
}
extern "C" int64_t yb_band_scan_avx2(int, int, const int32_t *, const int32_t *, int32_t *, int32_t *, int32_t *,
                                     int (*)(int, int), int32_t *, int32_t *);
extern "C" int64_t yb_band_scan(int M, int N, const int32_t *LB, const int32_t *RB, int32_t *wmax, int32_t *sched,
                                int32_t *nSteps, int (*lanesOf)(int, int), int32_t *lanes, int32_t *connected) {
    // YB_NO_AVX2=1 forces the baseline build (the CPU suite runs both)
    static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("YB_NO_AVX2");
    return avx2 ? yb_band_scan_avx2(M, N, LB, RB, wmax, sched, nSteps, lanesOf, lanes, connected)
                : YB_BAND_SCAN_NAME(M, N, LB, RB, wmax, sched, nSteps, lanesOf, lanes, connected);
}
#endif
