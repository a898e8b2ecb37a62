/* score_dropin.c -- the reference-side binding of block scoring (and of block output): multiz's own `mafScoreRange` symbol
 * (mz_scores.h:19, defined in mz_scores.c:124-152), re-routed.  The reference's definition is kept, renamed at
 * compile time (mz_scores.c is compiled with -DmafScoreRange=ref_mafScoreRange, integration/Makefile); this file
 * supplies the symbol every caller links to (mz_preyama.c:79, multi_util.c:509,:581,:611,:794,:802) and decides:
 *
 *   speculative pass of the record/replay driver (yama_dropin.cpp): return 0.  The pass's output is discarded and
 *       scores never steer the host (they are only stored in mafAli::score and printed by mafWrite, maf.c:257-258),
 *       so the O(rows^2 * columns) loop is simply skipped there;
 *   real pass, default: the reference's own function (kept host code);
 *   real pass, YB_SCORE=gpu: yb_score_blocks() of libyama_b200.so, one synchronous call per block -- pays off for
 *       deep alignments (tens of rows), costs a launch latency per block for shallow ones.
 *
 * This file is C because it walks the reference's own structs (maf.h:28-58), included from the reference tree.
 */
#include <stdlib.h>
#include "maf.h"

void fatalf(const char *fmt, ...);                                   /* util.c:21 */
void fatal(const char *msg);                                          /* util.c:17 */
extern int **ss;                                                      /* mz_scores.h:8 */
double ref_mafScoreRange(struct mafAli *maf, int start, int size);    /* the reference's, renamed */
/* yama_dropin.cpp */
int yb_dropin_score_mode(void);                                       /* 0: host function, 1: skip (speculative pass), 2: GPU */
double yb_dropin_score(int nrows, const unsigned char *const *rows, int text_size, int start, int size);
double mafScoreRange(struct mafAli *maf, int start, int size);

/* mafWrite (maf.h, maf.c:251-284; maf.c is compiled with -DmafWrite=ref_mafWrite): formatting a block costs an
 * fprintf per row; a speculative pass's output goes nowhere, so it is not produced.  The real pass writes as ever. */
void ref_mafWrite(FILE *f, struct mafAli *maf);
void yb_maf_write(FILE *f, struct mafAli *maf);                       /* maf_dropin.c: the same bytes, one fwrite per block */

/* ---- deferred output (yama_dropin.cpp, the single-pass driver) -------------------------------------------------------
 * In the deferred pass yama() answers with a PLACEHOLDER alignment: every column of A and of B once and in order, its
 * residues replaced by the byte YB_MARK.  mafBuild (mz_preyama.c:38-81) turns it into a block whose rows, names, starts and
 * sizes are already the final ones -- they depend on which residues a row holds, not on where the gaps go -- and the
 * host hands that block to mafWrite(stdout, .).  Here it is recognised by its marker, copied (duplicate_ali, maf.c:463)
 * and kept with its position in the output stream; when the batch has been aligned, yb_defer_emit() gives the copy its
 * real text, scores it (mafScoreRange, mz_scores.c:124) and writes it with the reference's own mafWrite. */
#define YB_MARK 1
/* What mafScoreRange answers in the deferred pass for a real block: scores are only ever stored in mafAli::score and
 * printed (maf.c:257-258), and every computed score is the score of the whole block as it is later written
 * (multi_util.c:509, :611; mz_preyama.c:79) -- so the O(rows^2 * columns) loop is not run inside the host's pass; a block
 * that reaches mafWrite with this value is scored when the pass is over, all such blocks at once. */
#define YB_SCORE_LATER (-9.0e15)
int yb_dropin_defer_active(void);
void yb_dropin_defer_block(FILE *f, void *ali_copy, int placeholder);

static int is_placeholder(const struct mafAli *maf) {
    const struct mafComp *c = maf ? maf->components : NULL;
    const char *t;
    if (c == NULL || c->text == NULL) return 0;
    for (t = c->text; *t == '-'; ++t) {}
    return *t == YB_MARK;
}

void mafWrite(FILE *f, struct mafAli *maf) {
    static int keep = -1;
    if (keep < 0) keep = getenv("YB_SPEC_WRITE") != NULL;        /* measurement knob: format in speculative passes too */
    if (!keep && yb_dropin_score_mode() == 1) return;
    if (yb_dropin_defer_active()) {                               /* every block is written when the pass is over */
        yb_dropin_defer_block(f, duplicate_ali(maf), is_placeholder(maf));
        return;
    }
    yb_maf_write(f, maf);
}

/* Give a captured placeholder block its alignment (al: m_new columns of W bytes, as yama() returns them: column-major).
 * Rows of the alignment that hold no residue were dropped by mafBuild (mz_preyama.c:67-70): the block's components are
 * the remaining rows, in order.  Returns 0, or -1 if the rows do not match the block. */
int yb_defer_fill(void *ali_copy, const unsigned char *al, int m_new, int W) {
    struct mafAli *a = (struct mafAli *)ali_copy;
    struct mafComp *c = a->components;
    int r, j;
    for (r = 0; r < W; ++r) {
        int any = 0;
        for (j = 0; j < m_new && !any; ++j) any = al[(size_t)j * W + r] != '-';
        if (!any) continue;
        if (c == NULL) return -1;
        free(c->text);
        c->text = (char *)malloc((size_t)m_new + 1);
        if (c->text == NULL) fatal("yama_b200: out of memory");
        for (j = 0; j < m_new; ++j) c->text[j] = (char)al[(size_t)j * W + r];
        c->text[m_new] = 0;
        c = c->next;
    }
    if (c != NULL) return -1;
    a->textSize = m_new;
    a->score = YB_SCORE_LATER;                                        /* mz_preyama.c:79 */
    return 0;
}

int yb_defer_wants_score(void *ali_copy) { return ((struct mafAli *)ali_copy)->score == YB_SCORE_LATER; }
int yb_defer_text_size(void *ali_copy) { return ((struct mafAli *)ali_copy)->textSize; }
int yb_defer_rows(void *ali_copy, const unsigned char **rows, int cap) {
    int n = 0;
    struct mafComp *c;
    for (c = ((struct mafAli *)ali_copy)->components; c != NULL; c = c->next, ++n)
        if (n < cap) rows[n] = (const unsigned char *)c->text;
    return n;
}
double yb_defer_host_score(void *ali_copy) {
    struct mafAli *a = (struct mafAli *)ali_copy;
    return ref_mafScoreRange(a, 0, a->textSize);
}
void yb_defer_write(FILE *f, void *ali_copy, int have_score, double score) {
    struct mafAli *a = (struct mafAli *)ali_copy;
    if (have_score) a->score = score;
    yb_maf_write(f, a);
    mafAliFree(&a);
}

double mafScoreRange(struct mafAli *maf, int start, int size) {
    const int mode = yb_dropin_score_mode();
    if (yb_dropin_defer_active()) {
        if (is_placeholder(maf)) return 0.0;                     /* scored when it has its real text */
        if (start < 0 || size <= 0 || start + size > maf->textSize)
            fatalf("mafScoreRange: start = %d, size = %d, textSize = %d\n", start, size, maf->textSize);
        if (ss == NULL) fatal("mafScoreRange: scores not initialized");
        return YB_SCORE_LATER;
    }
    if (mode == 0) return ref_mafScoreRange(maf, start, size);
    /* the reference's checks, in its order and wording (mz_scores.c:130-134) */
    if (start < 0 || size <= 0 || start + size > maf->textSize)
        fatalf("mafScoreRange: start = %d, size = %d, textSize = %d\n", start, size, maf->textSize);
    if (ss == NULL)
        fatal("mafScoreRange: scores not initialized");
    if (mode == 1) return 0.0;
    {
        static const unsigned char **rows = NULL;
        static int cap = 0;
        int n = 0;
        struct mafComp *c;
        for (c = maf->components; c != NULL; c = c->next) {
            if (n == cap) {
                cap = cap ? 2 * cap : 64;
                rows = realloc(rows, (size_t)cap * sizeof *rows);
                if (rows == NULL) fatal("mafScoreRange: out of memory");
            }
            rows[n++] = (const unsigned char *)c->text;
        }
        return yb_dropin_score(n, rows, maf->textSize, start, size);
    }
}
