/* ref_hook.c -- TEST INFRASTRUCTURE ONLY (part of oracle/; never linked into the product).
 *
 * Thin harness around the UNMODIFIED reference sources, compiled where they lie under
 * /root/reference by oracle/Makefile into oracle/_ref/libyama_ref.so.  mz_yama.c is compiled with
 * -Dfree=refhook_free so that the five free() calls at the end of the reference's yama()
 * (mz_yama.c:315-319: tback_row, tback, dp, dashes, script) pass through refhook_free(), which lets
 * us look at the traceback matrix, the last DP row and the edit script before they are released --
 * without touching a line of the reference.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#undef free

typedef unsigned char uchar;

/* reference symbols (mz_yama.h:22, mz_scores.h:17-18) */
void yama(uchar **A, int K, int M, uchar **B, int L, int N, int *LB, int *RB, uchar ***OAL, int *OM);
void init_scores70(void);
void init_scores85(void);
extern int **ss, *gop, gap_open, gap_extend;
extern char *argv0;

static struct {
    int armed, nfree;
    long tback_size;
    int N, MN;
    uchar *tback_out;      /* tback_size bytes or NULL */
    int *final_cdi;        /* 3 ints (C,D,I at grid point (M,N)) or NULL */
    uchar *script_out;     /* M+N bytes or NULL */
} H;

void refhook_free(void *p) {
    if (H.armed) {
        int k = H.nfree++;
        if (k == 1 && H.tback_out)              /* 2nd free: tback (mz_yama.c:316) */
            memcpy(H.tback_out, p, (size_t)H.tback_size);
        else if (k == 2 && H.final_cdi) {       /* 3rd free: dp (mz_yama.c:317); dp_node = {D,C,I,pad} */
            int *dpN = (int *)p + 4 * (size_t)H.N;
            H.final_cdi[0] = dpN[1];
            H.final_cdi[1] = dpN[0];
            H.final_cdi[2] = dpN[2];
        } else if (k == 4 && H.script_out)      /* 5th free: script (mz_yama.c:319) */
            memcpy(H.script_out, p, (size_t)H.MN);
    }
    free(p);
}

void ref_init_scores(int which) {
    static char name[] = "ref";
    argv0 = name;
    if (which == 85) init_scores85(); else init_scores70();
}

void ref_get_tables(int *ss_out /*128*128*/, int *gop_out /*16*/, int *gap_ext) {
    for (int i = 0; i < 128; i++) memcpy(ss_out + 128 * i, ss[i], 128 * sizeof(int));
    memcpy(gop_out, gop, 16 * sizeof(int));
    *gap_ext = gap_extend;
}

/* A, B: contiguous column-major buffers (column i at A + (i-1)*K).  out_al: caller buffer of
 * (M+N)*(K+L) bytes.  Returns m_new.  tback_out/final_cdi/script_out may be NULL. */
int ref_yama(const uchar *Abuf, int K, int M, const uchar *Bbuf, int L, int N, int *LB, int *RB,
             uchar *out_al, uchar *tback_out, int *final_cdi, uchar *script_out) {
    uchar **A = (uchar **)malloc(sizeof(uchar *) * (size_t)(M > 0 ? M : 1)) - 1;
    uchar **B = (uchar **)malloc(sizeof(uchar *) * (size_t)(N > 0 ? N : 1)) - 1;
    for (int i = 1; i <= M; i++) A[i] = (uchar *)Abuf + (size_t)(i - 1) * K;
    for (int j = 1; j <= N; j++) B[j] = (uchar *)Bbuf + (size_t)(j - 1) * L;
    long ts = 0;
    for (int r = 0; r <= M; r++) ts += RB[r] - LB[r] + 1;
    H.armed = 1; H.nfree = 0; H.tback_size = ts; H.N = N; H.MN = M + N;
    H.tback_out = tback_out; H.final_cdi = final_cdi; H.script_out = script_out;
    uchar **AL = NULL; int m_new = 0;
    yama(A, K, M, B, L, N, LB, RB, &AL, &m_new);
    H.armed = 0;
    if (out_al) memcpy(out_al, AL[1], (size_t)m_new * (size_t)(K + L));
    free(AL[1]); free(AL + 1);            /* mz_yama.h:17-18 */
    free(A + 1); free(B + 1);
    return m_new;
}

/* Time `reps` calls without copying anything out (CPU baseline, bench.py cpu_baseline leg). */
long ref_yama_cells(int M, int *LB, int *RB) {
    long ts = 0;
    for (int r = 0; r <= M; r++) ts += RB[r] - LB[r] + 1;
    return ts;
}
