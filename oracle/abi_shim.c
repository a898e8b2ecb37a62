/* abi_shim.c -- TEST INFRASTRUCTURE ONLY: the subset of the C ABI of include/yama_b200.h that
 * integration/yama_dropin.cpp uses, answered by the CPU oracle (yama_oracle.c).  It exists so that the CPU
 * suite can exercise the HOST logic of the drop-in (speculative record/replay around the unmodified
 * reference host) without a GPU: integration/_ref/bin/multiz_shim links this, the product binaries
 * (integration/_ref/bin/multiz, multic) link multiz_b200/libyama_b200.so and nothing from oracle/.
 * Built as oracle/libyama_shim.so by oracle/Makefile; never shipped, never measured. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/yama_b200.h"

typedef unsigned char uchar;
long oracle_check_band(int M, int N, const int *LB, const int *RB, char *msg, int msglen);
int oracle_yama(const uchar *A, int K, int M, const uchar *B, int L, int N, const int *LB, const int *RB,
                const int *ss, const int *gop, int gap_ext, uchar *out_al, uchar *tback_out, int *final_cdi,
                uchar *script_out, char *msg, int msglen);

struct yb_ctx {
    int32_t ss[128 * 128], gop[16], gap_ext;
    int have_scores;
    char err[512];
    uint8_t *scripts;
    size_t cap;
};

int yb_create(const int *devices, int ndev, yb_ctx **out) {
    (void)devices; (void)ndev;
    *out = (yb_ctx *)calloc(1, sizeof(yb_ctx));
    return *out ? YB_OK : YB_ERR_ARG;
}
void yb_destroy(yb_ctx *c) { if (c) { free(c->scripts); free(c); } }
const char *yb_last_error(const yb_ctx *c) { return c ? c->err : "no context"; }
int yb_device_count(const yb_ctx *c) { (void)c; return 0; }
int yb_set_scores(yb_ctx *c, const int32_t *ss, const int32_t *gop, int32_t ge) {
    memcpy(c->ss, ss, sizeof c->ss); memcpy(c->gop, gop, sizeof c->gop); c->gap_ext = ge; c->have_scores = 1;
    return YB_OK;
}
int64_t yb_check_band(int32_t M, int32_t N, const int32_t *LB, const int32_t *RB, char *msg, int msglen) {
    char tmp[256];
    long r = oracle_check_band(M, N, LB, RB, msg ? msg : tmp, msg ? msglen : (int)sizeof tmp);
    return r < 0 ? YB_ERR_BAND : r;
}
int yb_run_batch(yb_ctx *c, int64_t n, const yb_job *jobs, yb_result *res, yb_stats *st) {
    size_t need = 16;
    for (int64_t i = 0; i < n; i++) need += (size_t)jobs[i].M + jobs[i].N + 1;
    if (need > c->cap) { free(c->scripts); c->scripts = (uint8_t *)malloc(need); c->cap = need; }
    size_t off = 0;
    int rc = YB_OK;
    int64_t cells = 0;
    for (int64_t i = 0; i < n; i++) {
        const yb_job *j = &jobs[i];
        int cdi[3] = {0, 0, 0};
        memset(&res[i], 0, sizeof res[i]);
        long nc = oracle_check_band(j->M, j->N, j->LB, j->RB, c->err, sizeof c->err);
        if (nc < 0) { res[i].status = YB_ERR_BAND; rc = YB_ERR_BAND; continue; }
        uint8_t *ops = c->scripts + off;
        int m = oracle_yama(j->A, j->K, j->M, j->B, j->L, j->N, j->LB, j->RB, c->ss, c->gop, c->gap_ext, NULL, NULL,
                            cdi, ops, c->err, sizeof c->err);
        if (m < 0) { res[i].status = YB_ERR_TRACEBACK; rc = YB_ERR_TRACEBACK; continue; }
        for (int k = 0; k < m; k++) {             /* pack in place, 2 bits per op (yama_b200.h) */
            uint8_t op = ops[k];
            if ((k & 3) == 0) ops[k >> 2] = 0;
            ops[k >> 2] |= (uint8_t)(op << (2 * (k & 3)));
        }
        res[i].m_new = m; res[i].C = cdi[0]; res[i].D = cdi[1]; res[i].I = cdi[2];
        res[i].cells = nc; res[i].script = c->scripts + off;
        cells += nc;
        off += (size_t)j->M + j->N + 1;
    }
    if (st) { memset(st, 0, sizeof *st); st->cells = cells; st->pairs = n; }
    return rc;
}
int yb_assemble(const yb_job *job, const yb_result *res, uint8_t *out) {
    const int K = job->K, L = job->L, W = K + L;
    int i = 0, j = 0, m = 0;
    for (int e = res->m_new - 1; e >= 0; --e, ++m) {
        int op = (res->script[e >> 2] >> (2 * (e & 3))) & 3;
        uint8_t *dst = out + (size_t)m * W;
        if (op != 1) i++;
        if (op != 2) j++;
        if (op < 0 || op > 2 || i > job->M || j > job->N) return YB_ERR_TRACEBACK;
        if (op == 1) memset(dst, '-', (size_t)K); else memcpy(dst, job->A + (size_t)(i - 1) * K, (size_t)K);
        if (op == 2) memset(dst + K, '-', (size_t)L); else memcpy(dst + K, job->B + (size_t)(j - 1) * L, (size_t)L);
    }
    return (i == job->M && j == job->N) ? YB_OK : YB_ERR_TRACEBACK;
}
int oracle_score_range(int nrows, const uchar *const *rows, int text_size, int start, int size,
                       const int *ss, const int *gop, double *out);
int yb_score_blocks(yb_ctx *c, int64_t n, const yb_block *blocks, double *scores, yb_stats *st) {
    if (!c->have_scores) { snprintf(c->err, sizeof c->err, "mafScoreRange: scores not initialized"); return YB_ERR_SCORES; }
    for (int64_t i = 0; i < n; i++) {
        const yb_block *b = &blocks[i];
        int rc = oracle_score_range(b->nrows, b->rows, b->text_size, b->start, b->size, c->ss, c->gop, &scores[i]);
        if (rc) {
            snprintf(c->err, sizeof c->err, "mafScoreRange: start = %d, size = %d, textSize = %d\n", b->start, b->size, b->text_size);
            return YB_ERR_ARG;
        }
    }
    if (st) { memset(st, 0, sizeof *st); st->pairs = n; }
    return YB_OK;
}
