/* synth.c -- synthetic block-pair generator (measurement / test infrastructure, not the product).
 *
 * Produces yama() problems the way multiz's pre_yama() would hand them over (mz_preyama.c:174-259):
 *   A  K rows x M columns; row 0 plays the fixed reference row (v=1), so it is '-' only in columns
 *      that other species inserted.
 *   B  L rows x N columns, all-dash columns already removed (rmColDash, mz_preyama.c:87-108).
 *   LB/RB  from the columns where both blocks carry the same reference base (mz_preyama.c:240-258),
 *      then widened by smooth() (mz_preyama.c:17-35) to radius R.
 * Every species row is the ancestor with substitutions, per-site dashes and indel runs U[1,8].
 * Deterministic: pair i of a batch depends only on (seed, i) and its shape.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int64_t n;
    int32_t *K, *M, *L, *N;
    int64_t *offA, *offB, *offBand;  /* byte offsets into A/B, int offsets into LB/RB */
    uint8_t *A, *B;
    int32_t *LB, *RB;
    int64_t bytesA, bytesB, nBand;
    int64_t cells;
} synth_batch;

static inline uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
typedef struct { uint64_t s; } rng_t;
static inline uint64_t rnext(rng_t *r) {
    uint64_t x = r->s;
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    r->s = x;
    return x * 0x2545F4914F6CDD1Dull;
}
static inline double runif(rng_t *r) { return (double)(rnext(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline int rint_(rng_t *r, int lo, int hi) { return lo + (int)(rnext(r) % (uint64_t)(hi - lo + 1)); }
static const char BASES[4] = { 'A', 'C', 'G', 'T' };
static inline uint8_t rbase(rng_t *r) { return (uint8_t)BASES[rnext(r) & 3]; }
static inline uint8_t mutate(rng_t *r, uint8_t anc, double sub, double lower) {
    uint8_t b = anc;
    if (runif(r) < sub) b = (uint8_t)BASES[((anc == 'A' ? 0 : anc == 'C' ? 1 : anc == 'G' ? 2 : 3) + 1 + (int)(rnext(r) % 3)) & 3];
    if (lower > 0 && runif(r) < lower) {
        if (runif(r) < 0.1) b = 'N';
        else b = (uint8_t)(b | 0x20);
    }
    return b;
}

static void smooth(int *LB, int *RB, int M, int N, int radius) {     /* mz_preyama.c:17-35 */
    int rad = M < radius ? M : radius, j, i;
    for (i = j = 0; i <= M; ++i) { if (LB[i] > j) j = LB[i]; LB[i] = j; }
    for (i = M, j = N; i >= 0; --i) { if (RB[i] < j) j = RB[i]; RB[i] = j; }
    for (i = M; i > rad; --i) {
        int a = LB[i] - rad; if (a < 0) a = 0;
        LB[i] = a < LB[i - rad] ? a : LB[i - rad];
    }
    for (; i >= 0; --i) LB[i] = 0;
    for (i = 0; i < M - rad; ++i) {
        int a = RB[i] + rad; if (a > N) a = N;
        RB[i] = a > RB[i + rad] ? a : RB[i + rad];
    }
    for (; i <= M; ++i) RB[i] = N;
}

/* one pair; A must hold K*M bytes, B up to L*(2*M+16), LB/RB M+1 ints.  Returns N. */
static int make_pair(uint64_t seed, int K, int L, int M, int R, double sub, double dash, double indel,
                     double lower, uint8_t *A, uint8_t *B, int *LB, int *RB) {
    rng_t rg = { mix64(seed) | 1 };
    int capN = 2 * M + 16, N = 0;
    int *guide = (int *)malloc(sizeof(int) * (size_t)(M + 1));   /* B column of A column i, or -1 */
    int del_left = 0, ains_left = 0;
    for (int i = 1; i <= M; ++i) {
        uint8_t *acol = A + (size_t)(i - 1) * K;
        uint8_t anc = rbase(&rg);
        guide[i] = -1;
        /* a run of columns that exist only in A's non-reference rows (reference row is '-') */
        if (ains_left == 0 && K > 1 && i > 1 && i < M && runif(&rg) < indel) ains_left = rint_(&rg, 1, 8);
        if (ains_left > 0) {
            --ains_left;
            acol[0] = '-';
            int any = 0;
            for (int k = 1; k < K; ++k) {
                acol[k] = runif(&rg) < 0.5 ? '-' : mutate(&rg, anc, sub, lower);
                any |= acol[k] != '-';
            }
            if (!any) acol[1] = anc;
            continue;
        }
        acol[0] = mutate(&rg, anc, 0.0, lower);
        for (int k = 1; k < K; ++k) acol[k] = runif(&rg) < dash ? '-' : mutate(&rg, anc, sub, lower);
        /* columns present only in B (insertion in the second block) */
        if (i > 1 && runif(&rg) < indel) {
            int len = rint_(&rg, 1, 8);
            for (int q = 0; q < len && N < capN; ++q) {
                uint8_t *bcol = B + (size_t)N * L;
                uint8_t b2 = rbase(&rg);
                int any = 0;
                for (int l = 0; l < L; ++l) { bcol[l] = runif(&rg) < 0.5 ? '-' : mutate(&rg, b2, sub, lower); any |= bcol[l] != '-'; }
                if (!any) bcol[0] = b2;
                ++N;
            }
        }
        /* reference bases the second block does not cover with any residue (all-dash column, removed) */
        if (del_left == 0 && i > 1 && i < M && runif(&rg) < indel) del_left = rint_(&rg, 1, 8);
        if (del_left > 0) { --del_left; continue; }
        if (N < capN) {
            uint8_t *bcol = B + (size_t)N * L;
            int any = 0;
            for (int l = 0; l < L; ++l) { bcol[l] = runif(&rg) < dash ? '-' : mutate(&rg, anc, sub, lower); any |= bcol[l] != '-'; }
            if (!any) continue;     /* would have been removed by rmColDash */
            ++N;
            guide[i] = N;
        }
    }
    if (N == 0) {                    /* pre_yama never calls yama with N < 1 (mz_preyama.c:183) */
        for (int l = 0; l < L; ++l) B[l] = rbase(&rg);
        N = 1;
        guide[M] = 1;
    }
    for (int i = 0; i <= M; ++i) { LB[i] = 0; RB[i] = N; }
    for (int i = 1; i <= M; ++i)
        if (guide[i] > 0) {          /* mz_preyama.c:252-255 */
            if (LB[i] == 0 || LB[i] > guide[i]) LB[i] = guide[i];
            if (RB[i] == N || RB[i] < guide[i]) RB[i] = guide[i];
        }
    smooth(LB, RB, M, N, R);
    free(guide);
    return N;
}

synth_batch *synth_make(uint64_t seed, int64_t n, const int32_t *Ks, const int32_t *Ls, const int32_t *Ms, int R,
                        double sub, double dash, double indel, double lower) {
    synth_batch *sb = (synth_batch *)calloc(1, sizeof *sb);
    sb->n = n;
    sb->K = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    sb->M = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    sb->L = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    sb->N = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    sb->offA = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    sb->offB = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    sb->offBand = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    int64_t capA = 0, capB = 0, capBand = 0;
    for (int64_t i = 0; i < n; ++i) {
        capA += (int64_t)Ks[i] * Ms[i];
        capB += (int64_t)Ls[i] * (2 * (int64_t)Ms[i] + 16);
        capBand += Ms[i] + 1;
    }
    sb->A = (uint8_t *)malloc((size_t)capA + 16);
    uint8_t *Btmp = (uint8_t *)malloc((size_t)capB + 16);
    sb->LB = (int32_t *)malloc(sizeof(int32_t) * (size_t)capBand);
    sb->RB = (int32_t *)malloc(sizeof(int32_t) * (size_t)capBand);
    int64_t oa = 0, ob = 0, obd = 0, cells = 0;
    for (int64_t i = 0; i < n; ++i) {
        int K = Ks[i], L = Ls[i], M = Ms[i];
        sb->K[i] = K; sb->L[i] = L; sb->M[i] = M;
        sb->offA[i] = oa; sb->offB[i] = ob; sb->offBand[i] = obd;
        int N = make_pair(seed * 0x100000001b3ull + (uint64_t)i, K, L, M, R, sub, dash, indel, lower, sb->A + oa,
                          Btmp + ob, sb->LB + obd, sb->RB + obd);
        sb->N[i] = N;
        for (int r = 0; r <= M; ++r) cells += sb->RB[obd + r] - sb->LB[obd + r] + 1;
        oa += (int64_t)K * M; ob += (int64_t)L * N; obd += M + 1;
    }
    sb->B = (uint8_t *)realloc(Btmp, (size_t)ob + 16);
    sb->bytesA = oa; sb->bytesB = ob; sb->nBand = obd; sb->cells = cells;
    return sb;
}

void synth_free(synth_batch *sb) {
    if (!sb) return;
    free(sb->K); free(sb->M); free(sb->L); free(sb->N);
    free(sb->offA); free(sb->offB); free(sb->offBand);
    free(sb->A); free(sb->B); free(sb->LB); free(sb->RB);
    free(sb);
}
