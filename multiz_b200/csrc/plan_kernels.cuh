// plan_kernels.cuh -- K0: everything the fill needs to know about a wave's pairs, derived ON THE DEVICE from the band
// rows exactly as the caller handed them over (two int arrays LB, RB per pair, mz_yama.h:14-16):
//
//   yb_plan_kernel     one warp per pair: the validation of mz_yama.c:58-71 (LB[0]==0, RB[M]==N, width >= min(N,10),
//                      LB and RB non-decreasing), the cell count tback_size (mz_yama.c:60-66), the widest row, whether the
//                      band is connected, the kernel bin and class, the wavefront schedule (one offset per block of rows)
//                      and the step count.  Counts the pairs of every launch bucket.
//   yb_plan_scan       one CTA: prefix sums -- traceback bytes per pair -> the pairs' offsets in the traceback pool (pairs that
//                      no longer fit the pool are DEFERRED: the host runs them again later), bucket counts -> bucket starts
//                      and the kernel bins' ranges of the launch order; the wave summary, which stays on the device for the
//                      kernels that follow and reaches the host with the results.
//   yb_plan_scatter    one thread per pair: its place in the launch order (bin, then big pairs first) and in the list of
//                      long traceback paths.
//
// The host only lays the wave out from the pairs' DIMENSIONS (offsets are prefix sums of K*M, L*N, M+1 ...) and copies the
// callers' bytes; it never reads a band row and never waits for the plan: K1, K2 and K3 take their ranges from the summary
// in device memory, so a whole wave is queued in one go.  (Round 1 did all of this in a host pass -- band_scan.cpp, still behind
// yb_pair_facts -- which capped end-to-end throughput at what the host cores could scan: 0.23 scaling efficiency at 8 GPUs.)
#pragma once
#include "yama_kernels.cuh"

namespace yb {

constexpr int PLAN_NBINS = 9;              // kernel bins (yama_b200.cu: kBin), 5..8 are the bulk kernels
constexpr int PLAN_NB = 160;               // launch-order buckets per bin: quarter-octaves of the cell count, descending
constexpr int PLAN_BULK_BIN0 = 5;

struct PlanParams {
    int ring[PLAN_NBINS], warps[PLAN_NBINS], minRows2;   // bin geometry: ring entries, warps per pair; bin 2 takes pairs of >= minRows2 rows
    int skew[PLAN_NBINS];                                // fill_body3 bins: extra steps between the warps of a pair (else 0)
    int maxDepth, maxCls, maxAbsS, gapOpen, gapExt;      // limits and score magnitudes that decide the kernel class
    int tbLong;                                          // paths of at least this many moves go to the long-path list
    int slackBulk;                                       // schedule slack of the shuffle kernels (see yb_plan_kernel)
    unsigned launchMask;                                 // kernel bins that get a fill launch for this wave
};

// what the host reads back per wave (one small copy)
struct PlanSummary {
    int binStart[PLAN_NBINS + 1];          // ranges of the launch order per kernel bin
    int nValid, nLong, nFailed, firstFailed;             // firstFailed: lowest failing pair index of the wave (or -1)
    int nDeferred, pad;                    // pairs whose traceback matrix did not fit the wave's pool any more
    unsigned long long tbBytes;            // traceback pool bytes of the wave (the pairs that were placed)
    unsigned long long tbNeed;             // ... and of all its valid pairs, placed or deferred
    long long cells;                       // DP cells of the valid pairs
};

__device__ __forceinline__ int plan_bin_of(const PlanParams &pp, int wmax, int M) {
    if (wmax + 32 <= pp.ring[0]) return 0;
    if (wmax + 32 <= pp.ring[1]) return M >= pp.minRows2 ? 2 : 1;
    if ((pp.skew[3] ? f3_ring_need(wmax, pp.warps[3]) : wmax + 32) <= pp.ring[3]) return 3;
    if ((pp.skew[4] ? f3_ring_need(wmax, pp.warps[4]) : wmax + 32) <= pp.ring[4]) return 4;
    return -1;
}

constexpr int PLAN_THREADS = 256;

// ---- delta-coded bands ---------------------------------------------------------------------------------------------------
// The band is most of what a shallow pair sends over PCIe: 8 bytes per row (two ints, mz_yama.h:14-16) next to K + ~1
// bytes of sequence.  When the host has threads to spare it ships a pair's band as one byte per row and array -- the step
// X[r] - X[r-1], which is 0..254 for any band pre_yama builds except across a long indel -- and this kernel restores the
// caller's int arrays in device memory before anything reads them (K0 validates the restored rows, so an invalid band is
// still reported in the reference's words).  Steps outside 0..254 travel as exceptions: byte 255 plus an (index, step) entry.
struct BandPackHdr { int LB0, RB0, nExc, excStart; };      // then dl[align4(M+1)], dr[align4(M+1)] (pad bytes are 0)
struct BandExc { int idx, delta; };                        // idx: row r of LB, or M+1+r of RB
constexpr unsigned long long BAND_RAW = ~0ull;

__global__ void __launch_bounds__(PLAN_THREADS)
yb_band_expand(const PairMeta *__restrict__ metas, int nPairs, unsigned char *blob, const unsigned long long *__restrict__ packOff,
               const BandExc *__restrict__ exc) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    for (int p = warp; p < nPairs; p += nWarps) {
        const PairMeta pm = metas[p];
        const unsigned long long po = __ldg(packOff + p);
        if (pm.M < 1 || po == BAND_RAW) continue;
        const int M = pm.M;
        const BandPackHdr hdr = *reinterpret_cast<const BandPackHdr *>(blob + po);
        const int words = (M + 4) >> 2;                                 // align4(M+1) / 4
        const BandExc *ex = exc + hdr.excStart;
        for (int which = 0; which < 2; ++which) {
            const unsigned *src = reinterpret_cast<const unsigned *>(blob + po + sizeof(BandPackHdr)) + which * words;
            int *out = reinterpret_cast<int *>(blob + (which ? pm.offBand2 : pm.offBand));
            unsigned carry = (unsigned)(which ? hdr.RB0 : hdr.LB0);
            const int idx0 = which * (M + 1);
            for (int r0 = 0; r0 <= M; r0 += 128) {
                const int r = r0 + 4 * lane;
                const unsigned w = r <= M ? __ldg(src + (r >> 2)) : 0u;
                unsigned v0 = w & 0xffu, v1 = v0 + ((w >> 8) & 0xffu), v2 = v1 + ((w >> 16) & 0xffu), v3 = v2 + (w >> 24);
                unsigned inc = v3;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned o = __shfl_up_sync(FULL, inc, d);
                    if (lane >= d) inc += o;
                }
                const unsigned before = carry + inc - v3;
                v0 += before; v1 += before; v2 += before; v3 += before;
                for (int e = 0; e < hdr.nExc; ++e) {                    // (rare: a band step outside 0..254)
                    const BandExc x = ex[e];
                    const int at = x.idx - idx0;
                    if (at < 0 || at > M) continue;
                    const unsigned adj = (unsigned)x.delta - 255u;
                    if (r >= at) v0 += adj;
                    if (r + 1 >= at) v1 += adj;
                    if (r + 2 >= at) v2 += adj;
                    if (r + 3 >= at) v3 += adj;
                }
                if (r + 3 <= M) *reinterpret_cast<int4 *>(out + r) = make_int4((int)v0, (int)v1, (int)v2, (int)v3);
                else {
                    if (r <= M) out[r] = (int)v0;
                    if (r + 1 <= M) out[r + 1] = (int)v1;
                    if (r + 2 <= M) out[r + 2] = (int)v2;
                }
                carry += __shfl_sync(FULL, inc, 31);
            }
        }
    }
}

__global__ void __launch_bounds__(PLAN_THREADS)
yb_plan_kernel(PairMeta *metas, int nPairs, unsigned char *blob, PairOut *__restrict__ outs,
               unsigned long long *__restrict__ tbBytes, int *__restrict__ bucketOf,
               const __grid_constant__ PlanParams pp) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    for (int p = warp; p < nPairs; p += nWarps) {
        PairMeta pm = metas[p];
        PairOut o;
        o.m_new = 0; o.C = o.D = o.I = 0; o.status = 0; o.pad = 0; o.cells = 0;
        int bucket = -1;
        unsigned long long tbb = 0;
        if (pm.M < 1) {                                   // the host found the dimensions or pointers unusable
            o.status = -6;                                // YB_ERR_ARG
        } else {
            const int M = pm.M, N = pm.N;
            const int *LB = reinterpret_cast<const int *>(blob + pm.offBand);
            const int *RB = reinterpret_cast<const int *>(blob + pm.offBand2);
            const int need = N < 10 ? N : 10;
            // ---- mz_yama.c:58-71 + cell count, widest row, connectedness ------------------------------------------
            int bad = 0, gap = 0, wm = 0;
            long long cells = 0;
            for (int r = lane; r <= M; r += 32) {
                const int lb = __ldg(LB + r), rb = __ldg(RB + r);
                const int w = rb - lb;
                bad |= (w < need);
                cells += w + 1;
                wm = max(wm, w);
                if (r >= 1) {
                    const int lbp = __ldg(LB + r - 1), rbp = __ldg(RB + r - 1);
                    bad |= (lb < lbp) | (rb < rbp);
                    gap |= (lb > rbp + 1);
                }
            }
            if (lane == 0) bad |= (__ldg(LB) != 0) | (__ldg(RB + M) != N);
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                bad |= __shfl_xor_sync(FULL, bad, d);
                gap |= __shfl_xor_sync(FULL, gap, d);
                wm = max(wm, __shfl_xor_sync(FULL, wm, d));
                cells += __shfl_xor_sync(FULL, cells, d);
            }
            const int wmax = wm + 1;
            int bin = bad ? -1 : plan_bin_of(pp, wmax, M);
            if (bad) o.status = -2;                                            // YB_ERR_BAND
            else if (pm.K > pp.maxDepth || pm.L > 255 || bin < 0) o.status = -4;   // YB_ERR_LIMIT
            o.cells = bad ? 0 : cells;
            if (o.status == 0) {
                // ---- kernel class (see fill_body2): bounded scores, connected band, 16-bit weights -----------------
                int cls = 0;
                if (pp.maxCls >= 1 && bin <= 1 && !gap) {
                    const double work = ((double)M + N) * pm.K * pm.L * (double)(pp.gapOpen + pp.gapExt + pp.maxAbsS);
                    const long long w16 = max((long long)pm.K * (pp.gapOpen + pp.gapExt), 2ll * pm.K * pp.maxAbsS);
                    if (pp.maxCls >= 2 && 4 * w16 <= 32767 && work < (double)(1 << 26)) cls = 2;
                    else if (w16 <= 32767 && work < (double)(1 << 28)) cls = 1;
                    if (cls) bin = PLAN_BULK_BIN0 + 2 * (cls - 1) + bin;
                }
                // The host queues a fill kernel only for the bins recent waves used.  A pair of another bin moves to a
                // queued bin that can run it as well -- a wider ring of the same kind, a weaker kernel class (every KEYED
                // pair is a class-1 pair, every class-1 pair runs with existence multipliers too), one warp instead of a
                // CTA -- before it is deferred to a later wave (yb_plan_scan).
                for (int hop = 0; hop < 4 && !((pp.launchMask >> bin) & 1u); ++hop) {
                    int alt = -1;
                    switch (bin) {
                        case 7: alt = 5; break;                 // KEYED -> class 1 (ring 128)
                        case 8: alt = 6; break;                 //                  (ring 512)
                        case 5: alt = ((pp.launchMask >> 6) & 1u) ? 6 : 0; break;   // class 1 -> its wider ring, or with multipliers
                        case 6: alt = 1; break;
                        case 0: alt = 1; break;                 // ring 128 -> ring 512
                        case 2: alt = gap ? 1 : 8; break;       // a CTA per pair -> one warp (bulk if the band is connected ...)
                        case 3: alt = 4; break;                 // ring 2048 -> ring 4096
                        default: break;
                    }
                    if (alt == 8) {                             // ... and the scores are bounded: redo the class decision
                        const double work = ((double)M + N) * pm.K * pm.L * (double)(pp.gapOpen + pp.gapExt + pp.maxAbsS);
                        const long long w16 = max((long long)pm.K * (pp.gapOpen + pp.gapExt), 2ll * pm.K * pp.maxAbsS);
                        if (pp.maxCls >= 2 && 4 * w16 <= 32767 && work < (double)(1 << 26)) { alt = 8; cls = 2; }
                        else if (pp.maxCls >= 1 && w16 <= 32767 && work < (double)(1 << 28)) { alt = 6; cls = 1; }
                        else alt = 1;
                    }
                    if (alt < 0) break;
                    bin = alt;
                    cls = bin >= PLAN_BULK_BIN0 + 2 ? 2 : (bin >= PLAN_BULK_BIN0 ? 1 : 0);
                }
                // ---- wavefront schedule: rows Bb+1..Bb+B run on lanes 0..B-1 with column = step - (OFF_b + lane).  OFF grows
                // per block by at least B (lane 0 stays behind the last lane of the block above) and by enough that a lane
                // starts its next row only after the row below its current one has stopped reading it, `slack` = 3 steps
                // after that row's last column: the two-slot mailboxes of fill_body need them, and so does lane 0 of the
                // shuffle kernels -- it reads the ring at the top of a step, before a row switch in that step, so it must
                // have switched two steps before its new row's first cell for that cell's diagonal neighbour to be read
                // at the new row's column (measured: slack 2 saves 1.3 % of the fill and breaks 0.2 % of the pairs).
                // fill_body3 (skew > 0): the warps of a pair are `skew` extra steps apart, and lane 0 of a block of rows reads
                // the ring at least 8 steps after the last lane of the block above wrote it (another warp: a finished group)
                const int B = 32 * pp.warps[bin];
                const int skew = pp.skew[bin];
                const int slack = cls ? pp.slackBulk : 3;
                int *sched = reinterpret_cast<int *>(blob + pm.offSched);
                const int nblk = (M + B - 1) / B;
                int off = 0;
                for (int b = 0; b + 1 < nblk; ++b) {
                    if (lane == 0) sched[b] = off;
                    int nd = skew ? f3_min_advance(pp.warps[bin]) : B;
                    const int r0 = B * b + 1, r1 = min(M - B, B * b + B);
                    for (int r = r0 + lane; r <= r1; r += 32) nd = max(nd, __ldg(RB + r + 1) - __ldg(LB + r + B) + slack);
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) nd = max(nd, __shfl_xor_sync(FULL, nd, d));
                    off += nd;
                }
                if (lane == 0) sched[nblk - 1] = off;
                const int last = off + ((M - 1) % B) + skew * (((M - 1) % B) >> 5) + N;     // step of the last cell (RB[M] == N)
                const int nSteps = ((last + 2) + 7) & ~7;              // +1 step to publish the final scores, whole 8-step groups
                pm.nSteps = nSteps;
                pm.skew = skew;
                pm.lgLanes = 31 - __clz(B);
                pm.cls = cls;
                tbb = (unsigned long long)nSteps * (unsigned)B;        // one byte per lane and step
                const unsigned long long cc = (unsigned long long)max(cells, 1ll);
                const int lg = 63 - __clzll(cc);
                const int frac = lg >= 2 ? (int)((cc >> (lg - 2)) & 3) : 0;
                bucket = bin * PLAN_NB + (PLAN_NB - 1 - min(PLAN_NB - 1, lg * 4 + frac));
            }
        }
        if (lane == 0) {
            if (o.status == 0) {
                metas[p].nSteps = pm.nSteps; metas[p].lgLanes = pm.lgLanes; metas[p].cls = pm.cls; metas[p].skew = pm.skew;
            } else {
                metas[p].M = 0;                                        // K1..K3 skip the pair
            }
            outs[p] = o;
            tbBytes[p] = tbb;
            bucketOf[p] = bucket;
        }
    }
}

constexpr int YB_DEFERRED = -100;          // internal status: the pair runs again in a later wave (never reaches the caller)

// One CTA.  tbBytes[p] -> exclusive prefix (the pairs' traceback offsets); pairs past the pool's capacity are deferred;
// bucket counts -> bucket starts; summary.
__global__ void __launch_bounds__(1024)
yb_plan_scan(int nPairs, unsigned long long *__restrict__ tbBytes, int *__restrict__ bucketOf, int *__restrict__ bucketCount,
             int *__restrict__ bucketFill, PairOut *__restrict__ outs, PairMeta *__restrict__ metas,
             PlanSummary *__restrict__ summary, int tbLong, unsigned long long tbCapacity, unsigned launchMask) {
    __shared__ unsigned long long part[1024];
    __shared__ long long cellPart[1024];
    __shared__ int intPart[4][1024];
    const int t = threadIdx.x, T = blockDim.x;
    const int per = (nPairs + T - 1) / T;
    const int lo = min(nPairs, t * per), hi = min(nPairs, lo + per);
    unsigned long long sum = 0;
    for (int p = lo; p < hi; ++p) sum += tbBytes[p];
    part[t] = sum;
    __syncthreads();
    if (t == 0) {                                   // (1024 partial sums: a serial scan by one thread is a microsecond)
        unsigned long long acc = 0;
        for (int k = 0; k < T; ++k) { const unsigned long long v = part[k]; part[k] = acc; acc += v; }
        summary->tbNeed = acc;
    }
    __syncthreads();
    unsigned long long acc = part[t], used = 0;
    long long cells = 0;
    int failed = 0, firstFailed = 0x7fffffff, nLong = 0, nDeferred = 0;
    for (int p = lo; p < hi; ++p) {
        const unsigned long long v = tbBytes[p];
        tbBytes[p] = acc;
        int b = bucketOf[p];
        // deferred: the pair does not fit the traceback pool of this wave any more, or the host did not queue a fill
        // kernel for its bin (it only queues the bins recent waves used: an empty launch is not free)
        if (b >= 0 && (acc + v > tbCapacity || !((launchMask >> (b / PLAN_NB)) & 1u))) {
            outs[p].status = YB_DEFERRED;
            outs[p].C = (int)(v >> 20) + 1;         // (the traceback bytes it needs, in MiB, for the host's second attempt)
            metas[p].M = 0;
            bucketOf[p] = b = -1;
            ++nDeferred;
        } else if (b >= 0) {
            atomicAdd(bucketCount + b, 1);
            cells += outs[p].cells;
            used = acc + v;
            if (metas[p].M + metas[p].N >= tbLong) ++nLong;
        } else {
            ++failed;
            firstFailed = min(firstFailed, p);
        }
        acc += v;
    }
    part[t] = used; cellPart[t] = cells;
    intPart[0][t] = failed; intPart[1][t] = firstFailed; intPart[2][t] = nLong; intPart[3][t] = nDeferred;
    __syncthreads();
    if (t == 0) {
        unsigned long long u = 0;
        long long c = 0;
        int f = 0, ff = 0x7fffffff, nl = 0, nd = 0;
        for (int k = 0; k < T; ++k) {
            u = max(u, part[k]); c += cellPart[k];
            f += intPart[0][k]; ff = min(ff, intPart[1][k]); nl += intPart[2][k]; nd += intPart[3][k];
        }
        summary->tbBytes = u; summary->cells = c;
        summary->nFailed = f; summary->firstFailed = ff == 0x7fffffff ? -1 : ff; summary->nLong = nl; summary->nDeferred = nd; summary->pad = 0;
        // bucket starts (PLAN_NBINS * PLAN_NB counts) and the bins' ranges
        int a = 0;
        for (int b = 0; b < PLAN_NBINS; ++b) {
            summary->binStart[b] = a;
            for (int q = 0; q < PLAN_NB; ++q) {
                const int cnt = bucketCount[b * PLAN_NB + q];
                bucketCount[b * PLAN_NB + q] = a;                      // now: the bucket's first position
                bucketFill[b * PLAN_NB + q] = 0;
                a += cnt;
            }
        }
        summary->binStart[PLAN_NBINS] = a;
        summary->nValid = a;
    }
}

__global__ void __launch_bounds__(256)
yb_plan_scatter(int nPairs, const int *__restrict__ bucketOf, const int *__restrict__ bucketStart, int *__restrict__ bucketFill,
                const PairMeta *__restrict__ metas, int *__restrict__ order, int *__restrict__ longList, int *__restrict__ longFill,
                int tbLong) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nPairs) return;
    const int b = bucketOf[p];
    if (b < 0) return;
    order[bucketStart[b] + atomicAdd(bucketFill + b, 1)] = p;
    if (metas[p].M + metas[p].N >= tbLong) longList[atomicAdd(longFill, 1)] = p;
}

}  // namespace yb
