/* yama_b200.h -- C ABI of libyama_b200.so: the B200 (sm_100a) replacement for multiz's `yama`
 * hot path (banded affine-gap profile-profile DP + traceback).
 *
 * Reference interfaces this boundary replaces (paths relative to the multiz source tree):
 *   - void yama(uchar **A,int K,int M,uchar **B,int L,int N,int *LB,int *RB,uchar ***OAL,int *OM)
 *         mz_yama.h:22, defined mz_yama.c:50-320; callers mz_preyama.c:260 and :335.
 *   - the implicit score-table inputs  int **ss, *gop; int gap_open, gap_extend;
 *         mz_scores.h:8-11, filled by init_scores70/85 (mz_scores.c:94-122).
 *   - the error convention fatal()/fatalf() -> "<argv0>: msg\n" on stderr + exit(1), util.c:17-32.
 *
 * Plain C: pointers and sizes only.  No torch / C++ types cross this boundary.
 * There is NO CPU fallback behind any entry point: if no usable CUDA device is present every call
 * that needs one returns YB_ERR_CUDA (and the drop-in yama() dies through fatal()).
 */
#ifndef YAMA_B200_H
#define YAMA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct yb_ctx yb_ctx;

enum {
    YB_OK = 0,
    YB_ERR_CUDA = -1,       /* CUDA runtime failure or no device                          */
    YB_ERR_BAND = -2,       /* LB/RB violate the checks of mz_yama.c:58-71                */
    YB_ERR_SCORES = -3,     /* ss/gop do not have the structure init_scores() produces    */
    YB_ERR_LIMIT = -4,      /* profile deeper than 255 rows / band wider than the kernel supports */
    YB_ERR_TRACEBACK = -5,  /* mz_yama.c:275/:290/:308/:311 conditions                     */
    YB_ERR_ARG = -6
};

/* One block pair = one reference yama() call (mz_yama.h:4-19).
 * A: K*M bytes, column i (1-based) at A+(i-1)*K, row k of that column at [k]  (A[i][k] of the
 *    reference; mz_preyama.c:203-205 allocates exactly this contiguous buffer).  B likewise.
 * LB, RB: M+1 ints each, the band of row 0..M. */
typedef struct {
    int32_t K, M, L, N;
    const uint8_t *A;
    const uint8_t *B;
    const int32_t *LB;
    const int32_t *RB;
} yb_job;

typedef struct {
    int32_t status;      /* YB_OK or a YB_ERR_* for this pair                                   */
    int32_t m_new;       /* merged width == number of edit ops (*OM of the reference)           */
    int32_t C, D, I;     /* the three node scores at grid point (M,N) (mz_yama.c:262-267); a node that no
                            alignment path reaches holds a value near INT_MIN/2 whose low bits are unspecified */
    int32_t reserved;
    int64_t cells;       /* DP cells of this pair == tback_size of mz_yama.c:60-66              */
    const uint8_t *script; /* m_new ops in the reference's own (reversed) order, mz_yama.c:278, PACKED 2 bits
                              per op: op i sits in bits 2*(i&3) of byte i>>2; codes 0=C (both columns)
                              1=I (B column) 2=D (A column).  (m_new+3)/4 bytes.  Owned by ctx, valid until
                              the next yb_run_batch/yb_flush/yb_destroy.  yb_script_unpack() expands it.   */
} yb_result;

typedef struct {
    double kernel_ms;    /* device time of profile+fill+traceback kernels, max over devices     */
    double h2d_ms, d2h_ms, pack_ms, total_ms;
    int64_t cells, pairs;
    int64_t h2d_bytes, d2h_bytes;
    int32_t kernel_launches;
    int32_t n_devices;
    double fill_ms, profile_ms, traceback_ms;   /* device time split, device 0 */
    double plan_ms;      /* device time of the planning kernels (band checks, schedule, launch order), max over devices */
    int64_t staged_bytes; /* input bytes that were copied once on the host (callers outside yb_host_alloc memory)     */
} yb_stats;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* A context is used by one thread at a time (the reference is single-threaded); it runs its own
 * helper threads and streams inside a call.  Different contexts are independent, also across threads. */
/* devices==NULL or ndev<=0: use every visible CUDA device. */
int yb_create(const int *devices, int ndev, yb_ctx **out);
void yb_destroy(yb_ctx *ctx);
const char *yb_last_error(const yb_ctx *ctx);   /* message in the reference's own wording */
int yb_device_count(const yb_ctx *ctx);

/* Pinned host memory owned by the context.  A caller that builds its jobs' A, B, LB and RB inside such blocks (any
 * layout; one block or several) lets yb_run_batch copy them to the device as they are, with no pass over them on the
 * host; inputs anywhere else are staged through one memcpy.  Results are identical either way.  The reference has no
 * counterpart (its yama() reads malloc'd buffers in place, mz_preyama.c:174-205). */
void *yb_host_alloc(yb_ctx *ctx, size_t bytes);
void yb_host_free(yb_ctx *ctx, void *p);

/* ---- score tables (replaces reading the globals of mz_scores.h:8-11) ----------------------- */
/* ss: 128*128 ints row-major (ss[c][d] of the reference), gop: 16 ints, gap_extend.
 * Verifies the 6-class structure (ACGT/acgt, other, '-') of mz_scores.c:39-54 and the six-pattern
 * gop of :57-79, then keeps S6/gap_open/gap_extend with the context (they reach the kernels as
 * arguments, so contexts with different tables can coexist in one process). */
int yb_set_scores(yb_ctx *ctx, const int32_t *ss, const int32_t *gop, int32_t gap_extend);

/* ---- batched path (the product) ------------------------------------------------------------ */
/* Align n independent block pairs.  Inputs are host memory owned by the caller and only read
 * during the call.  Pairs are sharded over the context's devices in contiguous, cell-balanced
 * ranges; results come back in job order.  results: n entries. */
int yb_run_batch(yb_ctx *ctx, int64_t n, const yb_job *jobs, yb_result *results, yb_stats *stats);

/* Resident variant used for kernel-only measurement: load once, step many times, fetch once. */
int yb_resident_load(yb_ctx *ctx, int64_t n, const yb_job *jobs);
int yb_resident_step(yb_ctx *ctx, yb_stats *stats);         /* kernels only, inputs already in HBM */
int yb_resident_fetch(yb_ctx *ctx, yb_result *results);     /* D2H of scripts + scores */

/* ---- record / replay queue used by the drop-in yama() ---------------------------------------- */
/* Copies the job (inputs are freed by the caller right after yama returns, mz_preyama.c:350-357). */
int64_t yb_submit(yb_ctx *ctx, const yb_job *job);            /* returns job id >= 0 or YB_ERR_* */
int yb_flush(yb_ctx *ctx, yb_stats *stats);                   /* runs everything submitted      */
int yb_fetch(yb_ctx *ctx, int64_t id, yb_result *out);
void yb_clear(yb_ctx *ctx);

/* Expands res->script into one byte per op (the reference's `script[]`, mz_yama.c:257-291): ops[i], i < m_new. */
int yb_script_unpack(const yb_result *res, uint8_t *ops);

/* ---- block scoring: mafScoreRange (mz_scores.c:124-152), SURVEY 8(f) rank 1 ------------------- */
/* One alignment block as mafScoreRange sees it: struct mafAli (maf.h:28-36) with its components' text
 * (struct mafComp::text, maf.h:43-58), and the column range to score. */
typedef struct {
    int32_t nrows;               /* number of components                                           */
    int32_t text_size;           /* mafAli::textSize                                               */
    int32_t start, size;         /* mafScoreRange's arguments: columns start .. start+size-1       */
    const uint8_t *const *rows;  /* nrows pointers to mafComp::text (text_size bytes, each < 128)  */
} yb_block;
/* scores[i] = mafScoreRange(block i, start, size) for n independent blocks, bit-identical doubles (the sums are
 * integers).  A bad range fails the whole call with YB_ERR_ARG and the reference's message (mz_scores.c:130-132);
 * no score tables -> YB_ERR_SCORES ("mafScoreRange: scores not initialized").  stats->cells counts row pairs x
 * columns, the reference's unit of work.  Runs on the context's first device. */
int yb_score_blocks(yb_ctx *ctx, int64_t n, const yb_block *blocks, double *scores, yb_stats *stats);

/* ---- column assembly (mz_yama.c:293-313), host side ----------------------------------------- */
/* Writes m_new*(K+L) bytes to out (caller-allocated). */
int yb_assemble(const yb_job *job, const yb_result *res, uint8_t *out);

/* Validation of mz_yama.c:58-71; returns cell count or YB_ERR_BAND with msg in the reference's
 * wording. */
int64_t yb_check_band(int32_t M, int32_t N, const int32_t *LB, const int32_t *RB, char *msg, int msglen);

/* Host-only facts about one pair, as the library derives them before packing: DP cells (tback_size of
 * mz_yama.c:60-66), widest band row and the number of wavefront steps of the fill kernel (the pair's
 * traceback matrix takes 32 bytes per step).  Runs the vectorised band scan AND the scalar restatement of
 * mz_yama.c:58-71 and returns YB_ERR_LIMIT should they ever disagree; YB_ERR_BAND (msg in the reference's
 * wording) for an invalid band. */
int yb_pair_facts(const yb_job *job, int64_t *cells, int32_t *wmax, int32_t *nsteps, char *msg, int msglen);

/* ---- sharding plan (the reference has no counterpart: it is single-process) ------ */
/* Cuts jobs 0..n-1, kept in reference order, into nparts contiguous ranges of near-equal cost, where
 * cost(job) = cells[job] + a fixed per-pair overhead.  cuts receives nparts+1 boundaries
 * (cuts[0]=0, cuts[nparts]=n).  The library uses exactly this plan to spread one batch over its
 * devices; a multi-process host (one rank per GPU) calls it to pick its own range. Host-only. */
int yb_plan_split(int64_t n, const int64_t *cells, int nparts, int64_t *cuts);

#ifdef __cplusplus
}
#endif
#endif
