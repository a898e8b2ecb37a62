/* ref_score_hook.c -- TEST INFRASTRUCTURE ONLY (part of oracle/; never linked into the product).
 *
 * Calls the UNMODIFIED reference mafScoreRange (mz_scores.c:124-152, compiled where it lies into
 * oracle/_ref/libyama_ref.so) on a block given as plain row pointers: builds the struct mafAli /
 * struct mafComp list the reference expects (declared by the reference's own maf.h, included from
 * /root/reference at build time) around the caller's text, without copying it.
 */
#include <stdlib.h>
#include "maf.h"

double mafScoreRange(struct mafAli *maf, int start, int size);   /* mz_scores.h:19 */

double ref_score_range(int nrows, const unsigned char *const *rows, int text_size, int start, int size) {
    struct mafAli ali;
    struct mafComp *comps = nrows > 0 ? calloc((size_t)nrows, sizeof *comps) : NULL;
    for (int j = 0; j < nrows; ++j) {
        comps[j].text = (char *)rows[j];
        comps[j].next = j + 1 < nrows ? &comps[j + 1] : NULL;
    }
    ali.next = NULL; ali.score = 0.0; ali.components = comps; ali.textSize = text_size; ali.chain_len = 0;
    double sc = mafScoreRange(&ali, start, size);
    free(comps);
    return sc;
}
