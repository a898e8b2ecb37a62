/* yama_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C CPU restatement of the reference's banded profile-profile alignment
 * (reference: /root/reference/mz_yama.c:50-320, score macros mz_scores.h:13-15, band smoothing
 * mz_preyama.c:17-35).  It exists to CHECK the CUDA path: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load it.  The product (multiz_b200/csrc) never links, imports or
 * calls anything in oracle/.
 *
 * Parity pin: this file is itself checked, byte for byte (traceback matrix, final C/D/I, edit
 * script, output columns), against the UNMODIFIED reference compiled into oracle/_ref/
 * (oracle/Makefile, oracle/ref_hook.c) and against the committed fixtures in tests/golden/ that
 * were generated from that build (tools/make_golden.py).
 *
 * It deliberately keeps the reference's O(K*L)-per-cell character loops and table lookups
 * (ss[128][128], gop[16]) instead of the count-vector closed form the CUDA kernels use, so that a
 * kernel-vs-oracle comparison is a comparison of two different derivations.
 *
 * Arithmetic is int32 with two's-complement wrap (compile with -fwrapv), like the GPU.
 */
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned char uchar;

#define NEG_HALF (INT_MIN / 2)          /* mz_yama.c:29  MININT */
enum { TB_C = 0, TB_I = 1, TB_D = 2 };  /* mz_yama.c:24-26 */

typedef struct {
    const int *ss;   /* 128*128 substitution scores, row-major      (mz_scores.h:8)  */
    const int *gop;  /* 16 quasi-natural gap-open penalties          (mz_scores.h:9)  */
    int gap_ext;     /* per-residue gap extension                    (mz_scores.h:11) */
} score_tab;

static inline int is_dash(uchar ch) { return ch == '-'; }

/* gop index, mz_scores.h:14:  GAP(s,t,u,v) = gop[(s<<3)+(t<<2)+(u<<1)+v]
 * (s,t) = dash state of the two rows in the previous column, (u,v) in the current one. */
static inline int gap_pen(const score_tab *T, int s, int t, int u, int v) {
    return T->gop[(s << 3) + (t << 2) + (u << 1) + v];
}

/* three-way choice with the reference's tie rule (mz_yama.c:138-154,189-198,226-235):
 * from-C wins ties; between from-D and from-I, from-D needs to be strictly larger. */
static inline int choose3(int fromC, int fromD, int fromI, int *best) {
    if (fromC >= fromD && fromC >= fromI) { *best = fromC; return TB_C; }
    if (fromD > fromI) { *best = fromD; return TB_D; }
    *best = fromI;
    return TB_I;
}

static int count_residues(const uchar *col, int n) {
    int k = 0;
    for (int i = 0; i < n; i++) k += !is_dash(col[i]);
    return k;
}

/* Validation, mz_yama.c:58-71.  Returns cell count (tback_size) or -1 with msg filled. */
long oracle_check_band(int M, int N, const int *LB, const int *RB, char *msg, int msglen) {
    long cells = 0;
    if (LB[0] != 0 || RB[M] != N) {
        snprintf(msg, msglen, "LB and RB not terminated properly: %d %d %d", LB[0], RB[M], N);
        return -1;
    }
    int need = N < 10 ? N : 10;
    for (int r = 0; r <= M; r++) {
        int w = RB[r] - LB[r];
        if (w < need) {
            snprintf(msg, msglen, "RB[%d] - LB[%d] < %d, %d %d %d", r, r, need, RB[r], LB[r], N);
            return -1;
        }
        cells += w + 1;
        if (r > 0 && LB[r] < LB[r - 1]) { snprintf(msg, msglen, "LB not monotonic"); return -1; }
        if (r > 0 && RB[r] < RB[r - 1]) { snprintf(msg, msglen, "RB not monotonic"); return -1; }
    }
    return cells;
}

/* Band smoothing, mz_preyama.c:17-35. */
void oracle_smooth(int *LB, int *RB, int M, int N, int radius) {
    int rad = M < radius ? M : radius, run;
    run = 0;
    for (int i = 0; i <= M; i++) { if (LB[i] > run) run = LB[i]; LB[i] = run; }
    run = N;
    for (int i = M; i >= 0; i--) { if (RB[i] < run) run = RB[i]; RB[i] = run; }
    for (int i = M; i >= 0; i--) {
        if (i > rad) {
            int a = LB[i] - rad; if (a < 0) a = 0;
            LB[i] = a < LB[i - rad] ? a : LB[i - rad];
        } else LB[i] = 0;
    }
    for (int i = 0; i <= M; i++) {
        if (i < M - rad) {
            int a = RB[i] + rad; if (a > N) a = N;
            RB[i] = a > RB[i + rad] ? a : RB[i + rad];
        } else RB[i] = N;
    }
}

/* Fill + traceback + column assembly.
 *   A: K*M bytes, column i (1-based) at A+(i-1)*K;  B likewise with L,N.
 *   out_al: (M+N)*(K+L) bytes; tback: cell-count bytes (row-major, band-compact, mz_yama.c:79,98);
 *   final_cdi: C,D,I at (M,N); script: M+N bytes, reversed order like mz_yama.c:278.
 * Any output pointer may be NULL.  Returns the merged width m_new, or -1 (msg filled). */
int oracle_yama(const uchar *A, int K, int M, const uchar *B, int L, int N, const int *LB,
                const int *RB, const int *ss, const int *gop, int gap_ext, uchar *out_al,
                uchar *tback_out, int *final_cdi, uchar *script_out, char *msg, int msglen) {
    score_tab T = { ss, gop, gap_ext };
    long cells = oracle_check_band(M, N, LB, RB, msg, msglen);
    if (cells < 0) return -1;

#define ACOL(i) (A + (size_t)((i) - 1) * K)
#define BCOL(j) (B + (size_t)((j) - 1) * L)

    uchar *tb = (uchar *)malloc((size_t)cells);
    size_t *rowbase = (size_t *)malloc(sizeof(size_t) * (size_t)(M + 1));
    /* one DP row, reused in place exactly like the reference (mz_yama.c:82): entries right of the
     * previous row's RB still hold their initial NEG_HALF when they are read. */
    int *vC = (int *)malloc(sizeof(int) * (size_t)(N + 1));
    int *vD = (int *)malloc(sizeof(int) * (size_t)(N + 1));
    int *vI = (int *)malloc(sizeof(int) * (size_t)(N + 1));

    /* row 0, mz_yama.c:83-94: leading insertions cost extension only (end gaps are free) */
    size_t w = 0;
    rowbase[0] = 0;
    vC[0] = vD[0] = vI[0] = 0;
    tb[w++] = 0;
    for (int c = 1; c <= N; c++) {
        vC[c] = vD[c] = NEG_HALF;
        if (c <= RB[0]) {
            vI[c] = vI[c - 1] - count_residues(BCOL(c), L) * K * T.gap_ext;
            tb[w++] = (uchar)(TB_I << 4);
        } else
            vI[c] = NEG_HALF;
    }

    int curC = NEG_HALF, curD = NEG_HALF, curI = NEG_HALF;
    for (int r = 1; r <= M; r++) {
        const uchar *a_now = ACOL(r);
        const uchar *a_up = r > 1 ? ACOL(r - 1) : NULL;
        int lo = LB[r], hi = RB[r];
        rowbase[r] = w - (size_t)lo;
        int dgC, dgD, dgI;              /* values of grid point (r-1, c-1) */
        if (LB[r - 1] <= lo - 1) { dgC = vC[lo - 1]; dgD = vD[lo - 1]; dgI = vI[lo - 1]; }
        else dgC = dgD = dgI = NEG_HALF;
        curC = curD = curI = NEG_HALF;  /* values of grid point (r, c-1) */
        int a_res = count_residues(a_now, K);

        for (int c = lo; c <= hi; c++) {
            const uchar *b_now = BCOL(c);
            const uchar *b_left = c > 1 ? BCOL(c - 1) : NULL;
            int x, y, z, fI = 0, fC = 0, fD, nI, nC, nD;

            /* ---- I node: horizontal edge from (r, c-1); mz_yama.c:114-166 */
            if (c > lo) {
                x = curC; y = curD; z = curI;
                if (r < M) {
                    int okx = c > LB[r - 1] + 1, okz = c > lo + 1;
                    for (int i = 0; i < K; i++) {
                        int s = is_dash(a_now[i]);
                        for (int j = 0; j < L; j++) {
                            int t = b_left ? is_dash(b_left[j]) : 0;
                            int v = is_dash(b_now[j]);
                            if (okx) x -= gap_pen(&T, s, t, 1, v);
                            y -= gap_pen(&T, s, 1, 1, v);
                            if (okz) z -= gap_pen(&T, 1, t, 1, v);
                        }
                    }
                }
                fI = choose3(x, y, z, &nI);
                nI -= count_residues(b_now, L) * K * T.gap_ext;
            } else
                nI = NEG_HALF;

            /* ---- C node: diagonal edge from (r-1, c-1); mz_yama.c:169-205 */
            if (c > LB[r - 1]) {
                x = dgC; y = dgD; z = dgI;
                if (c > 1) {
                    int okx = r > 1 && c > LB[r - 2] + 1, oky = r > 1, okz = c > LB[r - 1] + 1;
                    for (int i = 0; i < K; i++) {
                        int s = a_up ? is_dash(a_up[i]) : 0;
                        int u = is_dash(a_now[i]);
                        for (int j = 0; j < L; j++) {
                            int t = is_dash(b_left[j]);
                            int v = is_dash(b_now[j]);
                            if (okx) x -= gap_pen(&T, s, t, u, v);
                            if (oky) y -= gap_pen(&T, s, 1, u, v);
                            if (okz) z -= gap_pen(&T, 1, t, u, v);
                        }
                    }
                }
                fC = choose3(x, y, z, &nC);
                for (int i = 0; i < K; i++)
                    for (int j = 0; j < L; j++)
                        nC += T.ss[128 * a_now[i] + b_now[j]];
            } else
                nC = NEG_HALF;

            /* ---- D node: vertical edge from (r-1, c); mz_yama.c:208-242 */
            x = vC[c]; y = vD[c]; z = vI[c];
            if (c > 0 && c < N) {
                int okx = r > 1 && c > LB[r - 2], oky = r > 1, okz = c > LB[r - 1];
                for (int i = 0; i < K; i++) {
                    int s = a_up ? is_dash(a_up[i]) : 0;
                    int u = is_dash(a_now[i]);
                    for (int j = 0; j < L; j++) {
                        int t = is_dash(b_now[j]);
                        if (okx) x -= gap_pen(&T, s, t, u, 1);
                        if (oky) y -= gap_pen(&T, s, 1, u, 1);
                        if (okz) z -= gap_pen(&T, 1, t, u, 1);
                    }
                }
            }
            fD = choose3(x, y, z, &nD);
            nD -= a_res * L * T.gap_ext;

            dgC = vC[c]; dgD = vD[c]; dgI = vI[c];
            vC[c] = curC = nC; vD[c] = curD = nD; vI[c] = curI = nI;
            tb[w++] = (uchar)(fC | (fD << 2) | (fI << 4));
        }
    }
    if (final_cdi) { final_cdi[0] = curC; final_cdi[1] = curD; final_cdi[2] = curI; }
    if (tback_out) memcpy(tback_out, tb, (size_t)cells);

    /* traceback, mz_yama.c:257-291; note the different D/I tie rule at the final grid point */
    uchar *ops = (uchar *)malloc((size_t)(M + N) + 1);
    int n_ops = 0, r = M, c = N, node, rc = 0;
    if (curC >= curD && curC >= curI) node = TB_C;
    else if (curD >= curI) node = TB_D;
    else node = TB_I;
    while (r > 0 || c > 0) {
        if (r < 0 || c < 0 || n_ops >= M + N) {
            snprintf(msg, msglen, "Error generating edit script.");
            rc = -1;
            break;
        }
        uchar st = tb[rowbase[r] + (size_t)c];
        ops[n_ops++] = (uchar)node;
        if (node == TB_I) { c--; node = st >> 4; }
        else if (node == TB_D) { r--; node = (st >> 2) & 3; }
        else if (node == TB_C) { r--; c--; node = st & 3; }
        else { snprintf(msg, msglen, "illegal node type in traceback"); rc = -1; break; }
    }
    if (rc == 0 && script_out) memcpy(script_out, ops, (size_t)n_ops);

    /* column assembly, mz_yama.c:293-313 */
    if (rc == 0) {
        int i = 0, j = 0, m = 0, W = K + L;
        for (int e = n_ops - 1; e >= 0; e--) {
            uchar *dst = out_al ? out_al + (size_t)m * W : NULL;
            int op = ops[e];
            if (op != TB_C && op != TB_I && op != TB_D) {
                snprintf(msg, msglen, "Illegal edit op: %d", op); rc = -1; break;
            }
            if (op != TB_I) i++;
            if (op != TB_D) j++;
            if (dst) {
                if (op == TB_I) memset(dst, '-', (size_t)K); else memcpy(dst, ACOL(i), (size_t)K);
                if (op == TB_D) memset(dst + K, '-', (size_t)L); else memcpy(dst + K, BCOL(j), (size_t)L);
            }
            m++;
        }
        if (rc == 0 && (i != M || j != N)) {
            snprintf(msg, msglen, "new_align: i=%d, j=%d, m=%d, M=%d, N=%d, M_new=%d\n", i, j, j, M, N, n_ops);
            rc = -1;
        }
        if (rc == 0) rc = n_ops;
    }
    free(ops); free(vC); free(vD); free(vI); free(rowbase); free(tb);
    return rc;
#undef ACOL
#undef BCOL
}

/* Score tables, restating mz_scores.c:34-81 (HOXD70 :9-14,23-24; HOX85 :16-21,26-27). */
void oracle_scores(int which, int *ss /*128*128*/, int *gop /*16*/, int *gap_ext) {
    static const int M70[4][4] = { { 91, -114, -31, -123 }, { -114, 100, -125, -31 },
                                   { -31, -125, 100, -114 }, { -123, -31, -114, 91 } };
    static const int M85[4][4] = { { 86, -135, -68, -157 }, { -135, 100, -148, -68 },
                                   { -68, -148, 100, -135 }, { -157, -68, -135, 86 } };
    const int (*mat)[4] = which == 85 ? M85 : M70;
    int open = which == 85 ? 600 : 400, ext = which == 85 ? 50 : 30;
    const char up[] = "ACGT", lo[] = "acgt";
    for (int i = 0; i < 128 * 128; i++) ss[i] = -100;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            ss[128 * up[i] + up[j]] = ss[128 * lo[i] + up[j]] = mat[i][j];
            ss[128 * up[i] + lo[j]] = ss[128 * lo[i] + lo[j]] = mat[i][j];
        }
    for (int i = 0; i < 128; i++) ss[128 * '-' + i] = ss[128 * i + '-'] = -ext;
    ss[128 * '-' + '-'] = 0;
    for (int i = 0; i < 16; i++) gop[i] = 0;
    /* (s,t,u,v): a gap opens in row 2 / row 1 / ... -- the six patterns of mz_scores.c:61-79 */
    static const int opens[6] = { 0x1 /*0001*/, 0x2 /*0010*/, 0x6 /*0110*/, 0x9 /*1001*/, 0xD /*1101*/, 0xE /*1110*/ };
    for (int i = 0; i < 6; i++) gop[opens[i]] = open;
    *gap_ext = ext;
}
