/* score_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C CPU restatement of the reference's block scoring, mafScoreRange
 * (reference: /root/reference/mz_scores.c:124-152, macros mz_scores.h:13-15).  It exists to CHECK the
 * CUDA path (multiz_b200/csrc/score_kernels.cuh): only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load it; the product never links, imports or calls anything in oracle/.
 *
 * Parity pin: checked against the UNMODIFIED reference function compiled into oracle/_ref/
 * (oracle/Makefile, oracle/ref_score_hook.c) and against the fixture tests/golden/score_small.npz that
 * was generated from that build (tools/make_golden.py).
 *
 * It keeps the reference's O(rows^2 * columns) pair loop and table lookups (ss[128][128], gop[16]) on
 * purpose: the kernel uses a count-vector quadratic form, so kernel-vs-oracle compares two derivations.
 */
#include <stddef.h>

typedef unsigned char uchar;

/* returns 0 and *out = score; -1 for a bad range (mz_scores.c:130-132), -2 for missing tables (:133-134) */
int oracle_score_range(int nrows, const uchar *const *rows, int text_size, int start, int size,
                       const int *ss /*128*128*/, const int *gop /*16*/, double *out) {
    if (start < 0 || size <= 0 || (long)start + size > text_size) return -1;
    if (ss == NULL || gop == NULL) return -2;
    double score = 0.0;                                             /* mz_scores.c:136 */
    for (int i = start; i < start + size; ++i)                      /* :137 */
        for (int r1 = 0; r1 < nrows; ++r1) {                        /* :138, list order == row order */
            const uchar now1 = rows[r1][i];
            for (int r2 = r1 + 1; r2 < nrows; ++r2) {               /* :140 */
                const uchar now2 = rows[r2][i];
                score += ss[128 * now1 + now2];                     /* :142 SS(br,bi) */
                if (i > 0) {                                        /* :143 */
                    const int s = rows[r1][i - 1] == '-', t = rows[r2][i - 1] == '-';
                    const int u = now1 == '-', v = now2 == '-';
                    score -= gop[(s << 3) + (t << 2) + (u << 1) + v];   /* :146 GAP2, mz_scores.h:14-15 */
                }
            }
        }
    *out = score;
    return 0;
}
