// yama_kernels.cuh -- sm_100a kernels for the batched yama hot path.
//
// What is computed is defined by the reference (multiz mz_yama.c:50-320); HOW is new:
//
//  K1 yb_profile_kernel   one CTA per block pair.  Every alignment column is reduced to a count
//                         vector: 6 character classes (A,C,G,T,other,'-'; mz_scores.c:39-54 only ever
//                         distinguishes these) and the dash-transition counts between neighbouring
//                         columns that the quasi-natural gap costs need (mz_scores.c:57-79), four bytes
//                         per step with zero-byte masks and popc.  With those, each O(K*L) loop of
//                         mz_yama.c:124-137/:174-201/:212-225 collapses to a 4-term integer dot product of
//                         byte counts (dp4a) and the sum-of-pairs score (mz_yama.c:199-201) to a 6-term
//                         16x8-bit dot product (dp2a).
//  K2 yb_fill_kernel_w    a wavefront of B = 32*G lanes per block pair (G = 1: one warp per pair; G = 4/8:
//                         one CTA per pair, for wide bands), persistent, pairs pulled from a queue.  Lane
//                         l owns rows l+1, l+1+B, ... and walks its row left to right; lane l is always
//                         one column behind lane l-1, so at every step the wavefront advances one
//                         anti-diagonal of the band.  (C,D,I) of the row above arrive through a 2-slot
//                         shared-memory mailbox per lane; last lane -> lane 0 goes through a ring that
//                         holds one band row.  One traceback byte per cell (mz_yama.c:253) is packed 4 at
//                         a time into 32-bit stores.
//  K3 yb_traceback_*      follow the packed pointers (mz_yama.c:257-291) and emit the edit script as 2-bit
//                         codes: one thread per pair, or one warp per path for long paths.
//
// Exactness: every in-band cell gets exactly the reference's int32 values and traceback decisions, including
// the "predecessor does not exist -> no gap-open charge" guards (mz_yama.c:131-135,180-186,218-223) and
// the MININT-minus-penalty values on the right fringe (mz_yama.c:93-94).  The guards are carried as data:
// every mailbox record holds, next to (C,D,I), the multipliers (-gap_open or 0) that a consumer applies
// to its gap-open counts for the C and I node of that grid point.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace yb {

constexpr int MININT = -1073741824;  // INT_MIN/2, mz_yama.c:29
constexpr int FLAG_C = 0, FLAG_I = 1, FLAG_D = 2;  // mz_yama.c:24-26

struct ScoreConst {
    int S6[6][6];   // substitution score per class pair
    int gap_open;   // GO
    int gap_ext;    // GE
};
// (passed to the kernels by value: score tables belong to a context, not to the process)

// ---- per-pair descriptor (device copy) --------------------------------------------------------
struct PairMeta {
    int K, M, L, N;
    unsigned long long offA, offB;     // byte offsets into the input blob
    unsigned long long offBand;        // byte offset into the input blob: LB[0..M] as the caller's ints (mz_yama.h:14-16)
    unsigned long long offBand2;       // ... RB[0..M]
    unsigned long long offSched;       // byte offset into the input blob: wavefront schedule, ceil(M/32) ints
    unsigned long long rowBase;        // index of row 0's RowRec in the RowRec pool (M+1 records)
    unsigned long long colBase;        // index of column 0's ColRec in the ColRec pool (N+1 records)
    unsigned long long scriptBase;     // 32-bit word index of this pair's packed script (ceil((M+N)/16) words)
    // (the byte offset of the pair's traceback matrix depends on the band, not only on the dimensions: it lives in a
    //  separate per-wave array so that the host can fill the metas in parallel, before those offsets are known)
    int nSteps;                        // wavefront steps of this pair (K0, from the schedule)
    int skew;                          // extra steps between the last lane of a warp and the first lane of the next (fill_body3; else 0)
    int lgLanes;                       // log2 of the wavefront width: 32 lanes (one warp) ... 256 lanes (a CTA) per pair
    int cls;                           // kernel class: 0 fill_body / fill_body3 (RowRec / ColRec), 1 fill_body2, 2 fill_body2 KEYED (RowRec2 / bulk ColRec)
};

// LB[r] / RB[r] of a pair
struct BandView {
    const int *lbp, *rbp;
    __device__ __forceinline__ BandView(const unsigned char *blob, const PairMeta &pm)
        : lbp(reinterpret_cast<const int *>(blob + pm.offBand)), rbp(reinterpret_cast<const int *>(blob + pm.offBand2)) {}
    __device__ __forceinline__ int lb(int r) const { return __ldg(lbp + r); }
    __device__ __forceinline__ int rb(int r) const { return __ldg(rbp + r); }
};

// Row record, 64 B (4 x 16 B).  av* are byte-count vectors matched against the column words (see
// ColRec) with dp4a; everything a lane needs to run one row of the band.
struct __align__(16) RowRec {
    unsigned avXC, avYC, avZC, avXI;   // q0: C-node x,y,z and I-node x coefficient bytes (avYC/avYD: 16-bit pairs
                                       //     pre-multiplied by -gap_open when the wave runs in y16 mode)
    unsigned avXD, avYD, avZD;         // q1: D-node x,y,z coefficient bytes (bytes 2,3 only)
    int eD;                            //     ndA*L*gap_ext  (mz_yama.c:239-242)
    unsigned w01, w23, w45;            // q2: S6^T * classcount(A row) as int16 pairs (sum-of-pairs weights)
    int LB16;                          //     16*LB[r]
    int RB16, LBp16;                   // q3: 16*RB[r], 16*LB[r-1]
    int off;                           //     wavefront schedule: this row computes column (step - off)
    int RBn;                           //     warp-sized wavefronts: RB[r+1] (RB[r] on the last row), how far the row below
                                       //     reads us; CTA-sized wavefronts: 16*RB[r-1] + 16, where the row above ends;
                                       //     fill_body3 (PairMeta::skew > 0): 16*LB[r-2] (INT_MAX for row 1), and `off` is 16*off
};

// Traceback matrix layout: the wavefront (B = 32 << (lgLanes-5) lanes: one warp, or the warps of a CTA) advances one
// anti-diagonal per step; lane l computes one cell per step t (cell (r,c) with l = (r-1) mod B and t = c + off[r]).
// The byte of (l,t) lives at  (t>>3)*8B + 8*l + (t&7):  every 4 steps a lane stores one 32-bit word, and two
// consecutive stores of a warp fill 256 contiguous bytes.  A 32-B sector holds 4 lanes x 8 steps -- the shape a
// traceback path likes: a diagonal move is (l-1, t-2), an up move (l-1, t-1), a left move (l, t-1).
__device__ __forceinline__ unsigned long long tb_byte(unsigned lane, unsigned t, int lgLanes) {
    return ((unsigned long long)(t >> 3) << (lgLanes + 3)) + (lane << 3) + (t & 7u);
}

// Column record, 16 B.
//  w0 = (b01, b10, ndB, dB)           raw transition / dash counts of column c against column c-1
//  w1 = (n_A, n_C, n_G, n_T)          class counts
//  w2 = (n_X, n_dash, ndB', dB')      ndB',dB' zeroed for c==0 and c==N  (mz_yama.c:211, end gaps free)
//  w3 = w0 zeroed for c==1            (mz_yama.c:173, no gap-open at the start)
struct __align__(16) ColRec { unsigned w0, w1, w2, w3; };

struct PairOut { int m_new, C, D, I, status, pad; long long cells; };   // cells, status: K0; the rest: K2 / K3

__device__ __forceinline__ int classify(unsigned ch) {
    unsigned u = ch | 0x20u;
    if (ch == '-') return 5;
    if (u == 'a') return 0;
    if (u == 'c') return 1;
    if (u == 'g') return 2;
    if (u == 't') return 3;
    return 4;
}

__device__ __forceinline__ int dp4a_uu(unsigned a, unsigned b, int c) {
    int d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_lo_su(unsigned a16x2, unsigned b8x4, int c) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a16x2), "r"(b8x4), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(unsigned a16x2, unsigned b8x4, int c) {
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a16x2), "r"(b8x4), "r"(c));
    return d;
}
__device__ __forceinline__ unsigned pack4(unsigned b0, unsigned b1, unsigned b2, unsigned b3) {
    return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
}
__device__ __forceinline__ unsigned pack16(int lo, int hi) {
    return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16);
}

// ---- 4 bytes at a time: class counts and dash transitions of one alignment column / row --------------------
// 0x80 in every byte of x that is zero (exact, no carries across bytes)
__device__ __forceinline__ unsigned zero_bytes80(unsigned x) {
    return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
}
__device__ __forceinline__ unsigned eq_bytes80(unsigned w, unsigned pattern) { return zero_bytes80(w ^ pattern); }
// unaligned 32-bit read of bytes p..p+3 from two aligned words (blob sections are padded, reading past n is safe)
__device__ __forceinline__ unsigned load4(const unsigned char *p) {
    const unsigned *a = reinterpret_cast<const unsigned *>(reinterpret_cast<unsigned long long>(p) & ~3ull);
    const unsigned sh = ((unsigned)reinterpret_cast<unsigned long long>(p) & 3u) * 8u;
    return __funnelshift_r(__ldg(a), __ldg(a + 1), sh);
}
struct Census {
    unsigned n[6];          // A, C, G, T, other, dash  (mz_scores.c:39-54 distinguishes nothing else)
    unsigned t00, t01, t10, t11;   // (previous is dash, this is dash) transition counts; previous == none counts as non-dash
};
__device__ __forceinline__ Census census(const unsigned char *now, const unsigned char *prev, int n) {
    Census q;
#pragma unroll
    for (int k = 0; k < 6; ++k) q.n[k] = 0;
    q.t00 = q.t01 = q.t10 = q.t11 = 0;
    for (int j = 0; j < n; j += 4) {
        const unsigned valid = (n - j >= 4) ? 0x80808080u : (0x80808080u >> (8 * (4 - (n - j))));
        const unsigned w = load4(now + j);
        const unsigned lw = w | 0x20202020u;
        const unsigned mD = eq_bytes80(w, 0x2d2d2d2du) & valid;                 // '-'
        const unsigned mA = eq_bytes80(lw, 0x61616161u) & valid, mC = eq_bytes80(lw, 0x63636363u) & valid;
        const unsigned mG = eq_bytes80(lw, 0x67676767u) & valid, mT = eq_bytes80(lw, 0x74747474u) & valid;
        const unsigned pD = prev ? (eq_bytes80(load4(prev + j), 0x2d2d2d2du) & valid) : 0u;
        q.n[0] += __popc(mA); q.n[1] += __popc(mC); q.n[2] += __popc(mG); q.n[3] += __popc(mT); q.n[5] += __popc(mD);
        q.n[4] += __popc(valid & ~(mA | mC | mG | mT | mD));
        q.t11 += __popc(pD & mD); q.t10 += __popc(pD & ~mD); q.t01 += __popc(valid & ~pD & mD);
    }
    q.t00 = (unsigned)n - q.t01 - q.t10 - q.t11;
    return q;
}

// =================================================================================================
// K1: column / row profiles, traceback row offsets, wavefront schedule.  One CTA per pair.
// =================================================================================================
constexpr int K1_THREADS = 128;
struct RowRec2;
__device__ __forceinline__ void profile_bulk(const PairMeta &pm, const unsigned char *__restrict__ blob, RowRec *__restrict__ rowPool,
                                             ColRec *__restrict__ colPool, const ScoreConst &c_sc);

__global__ void __launch_bounds__(K1_THREADS, 10)     // 48 registers: 10 CTAs per SM hide more of the load latency (0.97 -> 0.92 ms on cfg2)
yb_profile_kernel(const PairMeta *__restrict__ metas, const unsigned char *__restrict__ blob,
                  RowRec *__restrict__ rowPool, ColRec *__restrict__ colPool, int y16,
                  const __grid_constant__ ScoreConst c_sc) {
    const PairMeta pm = metas[blockIdx.x];
    const int K = pm.K, M = pm.M, L = pm.L, N = pm.N;
    if (M < 1) return;                                  // invalid pair (rejected on the host)
    const unsigned char *A = blob + pm.offA;
    const unsigned char *B = blob + pm.offB;
    const BandView band(blob, pm);
    RowRec *rows = rowPool + pm.rowBase;
    ColRec *cols = colPool + pm.colBase;
    const int *sched = reinterpret_cast<const int *>(blob + pm.offSched);
    const int GE = c_sc.gap_ext;
    if (pm.cls != 0) { profile_bulk(pm, blob, rowPool, colPool, c_sc); return; }

    // ---- columns of B (c = 0..N) -------------------------------------------------------------
    for (int c = threadIdx.x; c <= N; c += K1_THREADS) {
        ColRec cr = {0u, 0u, 0u, 0u};
        if (c >= 1) {
            const unsigned char *now = B + (size_t)(c - 1) * L;
            const Census q = census(now, c > 1 ? now - L : nullptr, L);    // mz_yama.c:128 (t==0 when col==1)
            const unsigned *n = q.n;
            const unsigned b01 = q.t01, b10 = q.t10;
            unsigned dB = n[5], ndB = (unsigned)L - dB;
            cr.w0 = pack4(b01, b10, ndB, dB);
            cr.w1 = pack4(n[0], n[1], n[2], n[3]);
            bool inner = (c < N);                               // mz_yama.c:211
            cr.w2 = pack4(n[4], n[5], inner ? ndB : 0u, inner ? dB : 0u);
            cr.w3 = (c > 1) ? cr.w0 : 0u;                       // mz_yama.c:173
        }
        cols[c] = cr;
    }

    // ---- rows of A (r = 0..M) ------------------------------------------------------------------
    for (int r = threadIdx.x; r <= M; r += K1_THREADS) {
        RowRec rr;
        rr.avXC = rr.avYC = rr.avZC = rr.avXI = rr.avXD = rr.avYD = rr.avZD = 0u;
        rr.eD = 0; rr.w01 = rr.w23 = rr.w45 = 0u;
        const int lb = band.lb(r), lbp = r > 0 ? band.lb(r - 1) : 0;
        rr.LB16 = lb * 16; rr.RB16 = band.rb(r) * 16; rr.LBp16 = lbp * 16;
        rr.off = r >= 1 ? sched[(r - 1) >> pm.lgLanes] + ((r - 1) & ((1 << pm.lgLanes) - 1)) : 0;
        rr.RBn = r < M ? band.rb(r + 1) : band.rb(r);
        if (pm.lgLanes > 5 && r >= 1) rr.RBn = band.rb(r - 1) * 16 + 16;      // (row 0 keeps RB[1]: the ring initialisation)
        if (pm.skew > 0 && r >= 1) {                                           // fill_body3
            const int l = (r - 1) & ((1 << pm.lgLanes) - 1);
            rr.off = 16 * (rr.off + pm.skew * (l >> 5));
            rr.RBn = r >= 2 ? band.lb(r - 2) * 16 : 0x7fffffff;
        }
        if (r >= 1) {
            const unsigned char *now = A + (size_t)(r - 1) * K;
            const Census q = census(now, r > 1 ? now - K : nullptr, K);    // mz_yama.c:175,213 (s==0 when row==1)
            int n[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) n[k] = (int)q.n[k];
            const unsigned a00 = q.t00, a01 = q.t01, a10 = q.t10, a11 = q.t11;
            unsigned dA = (unsigned)n[5], ndA = (unsigned)K - dA;
            // gap-open counts as dot products with the column bytes (b01, b10, ndB, dB):
            //   C.x: a00*b01 + a11*b10 + a01*ndB + a10*dB   C.y: dA*ndB + a10*dB   C.z: ndA*dB + dA*b10
            //   I.x: ndA*ndB + dA*b10    I.y: K*ndB    I.z: K*b10
            //   D.x: ndA*ndB' + a10*dB'   D.y: a10*(ndB'+dB')   D.z: ndA*(ndB'+dB')   on bytes 2,3 of w2
            // The y candidates (from a D node) are never gated, so when K*gap_open fits 16 bits (y16) their weights are
            // stored already multiplied by -gap_open and K2 applies them with ONE dp2a on bytes 2,3.
            const int nGO = -c_sc.gap_open;
            if (r > 1) {                                        // mz_yama.c:180-184, :218-221 (row>1)
                rr.avXC = pack4(a00, a11, a01, a10);
                rr.avYC = y16 ? pack16((int)dA * nGO, (int)a10 * nGO) : pack4(0, 0, dA, a10);
                rr.avXD = pack4(0, 0, ndA, a10);
                rr.avYD = y16 ? pack16((int)a10 * nGO, (int)a10 * nGO) : pack4(0, 0, a10, a10);
            }
            rr.avZC = pack4(0, dA, 0, ndA);
            rr.avZD = pack4(0, 0, ndA, ndA);
            if (r < M) rr.avXI = pack4(0, dA, ndA, 0);         // mz_yama.c:123 (row<M)
            rr.eD = (int)ndA * L * GE;
            int w[6];
#pragma unroll
            for (int l = 0; l < 6; ++l) {
                int acc = 0;
#pragma unroll
                for (int k = 0; k < 6; ++k) acc += n[k] * c_sc.S6[k][l];
                w[l] = acc;
            }
            rr.w01 = pack16(w[0], w[1]); rr.w23 = pack16(w[2], w[3]); rr.w45 = pack16(w[4], w[5]);
        }
        rows[r] = rr;
    }
}

// =================================================================================================
// K2: banded three-state fill, one wavefront (warp or CTA) per pair.
// =================================================================================================
__device__ __forceinline__ unsigned smem_u32(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint4 lds128(unsigned addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(unsigned addr, int a, int b, int c, unsigned d) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// (a & b) ^ c in one LOP3: ring / mailbox slot addressing
__device__ __forceinline__ unsigned and_xor(unsigned a, unsigned b, unsigned c) {
    unsigned d;
    asm("lop3.b32 %0, %1, %2, %3, 0x6a;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// (a & b) | c in one LOP3
__device__ __forceinline__ unsigned and_or(unsigned a, unsigned b, unsigned c) {
    unsigned d;
    asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ unsigned launder(unsigned x) {
    asm volatile("" : "+r"(x));
    return x;
}

// Traceback byte of a cell: three 2-bit fields (C node at bit 0, D node at bit 2, I node at bit 4, like
// mz_yama.c:24-26,253), each holding  e = notC | gt<<1  with  notC = "the from-C candidate is not the maximum"
// and gt = "from-D is strictly greater than from-I" -- the two comparisons of the reference's tie rule
// (mz_yama.c:138-154): e even -> came from C, e == 1 -> from I, e == 3 -> from D.  A node that does not exist
// (mz_yama.c:163-165, :202-204) gets notC = 0, i.e. reads as "from C" = 0, the value the reference leaves there,
// so the traceback needs no band lookups.  Every comparison becomes ONE predicated add of an immediate into the
// running 32-bit word (4 cells per word), instead of a select chain.
template <int SH>
__device__ __forceinline__ int pick3(int x, int y, int z, bool exists, unsigned &acc) {
    const int m = __vimax3_s32(x, y, z);
    if ((x != m) && exists) acc += 1u << (24 + SH);
    if (y > z) acc += 2u << (24 + SH);
    return m;
}

// RING: ring entries, power of two, >= widest band row + 32.
// G:    warps per pair.  The 32*G lanes of a group form ONE wavefront (lane l owns rows l+1, l+1+32G, ...): G = 1 is a
//       warp per pair, synchronised with __syncwarp(); G > 1 is a CTA per pair, synchronised with __syncthreads() --
//       the form for wide bands, where one ring per PAIR (instead of per warp) keeps the shared-memory footprint small
//       enough for full occupancy, and long pairs get 32*G cells per step.
// P:    groups (= pairs in flight) per CTA; P > 1 only with G = 1.
// Y16:  every pair of the launch has K*gap_open <= 32767: the ungated y candidates and the I-node z candidate take
//       one dp2a with 16-bit weights instead of dp4a/extract + imad.
template <int RING, int G, int P, bool Y16, bool GATED = true>
__device__ __forceinline__ void
fill_body(const PairMeta *__restrict__ metas, const int *__restrict__ orderBase, const int *__restrict__ binRange,
          int *__restrict__ queue, const RowRec *__restrict__ rowPool,
          const ColRec *__restrict__ colPool, unsigned char *__restrict__ tbPool,
          const unsigned long long *__restrict__ tbBase, PairOut *__restrict__ outs, const int gapOpen, const int gapExt) {
    static_assert(G == 1 || P == 1, "several groups per CTA only for warp-sized groups");
    // the bin's slice of the launch order: [binRange[0], binRange[1]) as K0 left it in device memory (the host never sees it)
    const int *order = orderBase + __ldg(binRange);
    const int nPairs = __ldg(binRange + 1) - __ldg(binRange);
    constexpr int B = 32 * G;                                     // lanes of the wavefront
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [ P rings of RING records | P x B lanes x 2 mailbox records | queue slot ]; rings are RING*16-aligned
    const int grp = (G == 1) ? (int)(threadIdx.x >> 5) : 0;
    const int lane = (G == 1) ? (int)(threadIdx.x & 31) : (int)threadIdx.x;   // lane within the group's wavefront
    // dynamic smem base is at least 16-aligned; rings need RING*16 alignment for the (a&b)^c addressing
    const unsigned smem0 = (smem_u32(smem_raw) + RING * 16 - 1) & ~(unsigned)(RING * 16 - 1);
    const unsigned ringAddr = smem0 + grp * RING * 16;
    const unsigned boxAddr = smem0 + P * RING * 16 + grp * B * 32;
    const unsigned slotAddr = smem0 + P * RING * 16 + P * B * 32;
    // Warp-sized wavefronts: the steps of the mailbox protocol are separated by __syncwarp().  ptxas proves the warp
    // converged at every one of them (the only divergent region, the row switch, re-joins at its BSYNC before the barrier)
    // and emits no instruction for it -- not even for bar.warp.sync with a mask it cannot see through -- so
    // compute-sanitizer's racecheck sees no ordering.  -DYB_FORCE_WARPSYNC (the `make sanitize` build,
    // profiles/r2_sanitizer.md) separates the steps with a named CTA barrier of 32 threads per warp instead, which is
    // never elided: same protocol, visible to the tool.
#ifdef YB_FORCE_WARPSYNC
    auto group_sync = [&]() { if (G == 1) asm volatile("bar.sync %0, 32;" ::"r"(grp + 1) : "memory"); else __syncthreads(); };
#else
    auto group_sync = [&]() { if (G == 1) __syncwarp(); else __syncthreads(); };
#endif
    const int GO = gapOpen;
    const int nGO = -GO;
    const unsigned E_both = pack16(nGO, nGO);

    // Mailbox of lane l: 32 B at boxAddr + 32*l, two 16-B slots; slot of column c is (c&1) ^ ((l>>2)&1)
    // (the swizzle keeps 128-bit accesses of a quarter-warp on distinct banks).  Lane l reads what lane
    // l-1 wrote (lane 0 reads the ring), writes its own mailbox (the last lane writes the ring).
    auto boxOf = [&](int l) { return boxAddr + 32u * l + (((unsigned)l >> 2) & 1u) * 16u; };
    const unsigned rdBase = launder((lane == 0) ? ringAddr : boxOf(lane - 1));
    const unsigned rdMask = (lane == 0) ? (unsigned)(RING * 16 - 16) : 16u;
    const unsigned wrBase = launder((lane == B - 1) ? ringAddr : boxOf(lane));
    const unsigned wrMask = (lane == B - 1) ? (unsigned)(RING * 16 - 16) : 16u;

    for (;;) {
        int slot = 0;
        if (G == 1) {
            if (lane == 0) slot = atomicAdd(queue, 1);
            slot = __shfl_sync(0xffffffffu, slot, 0);
        } else {
            if (lane == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(slotAddr), "r"(atomicAdd(queue, 1)) : "memory");
            __syncthreads();
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(slot) : "r"(slotAddr) : "memory");
        }
        if (slot >= nPairs) break;
        const int p = order[slot];
        const PairMeta pm = metas[p];
        const int M = pm.M;
        const RowRec *rows = rowPool + pm.rowBase;
        const ColRec *cols = colPool + pm.colBase;
        unsigned char *tb = tbPool + __ldg(tbBase + p);
        const unsigned nKGE_lo = launder((unsigned)(-(pm.K * gapExt)) & 0xffffu);   // dp2a.hi weight of byte 2 (ndB)
        const int KGE = pm.K * gapExt;
        const int nSteps = pm.nSteps;
        const int N16 = pm.N * 16;
        const int KnGO = pm.K * nGO;                        // I-node y / z charge per residue / per closing gap of B
        const unsigned E_c0 = pack16(nGO, 0);

        // ---- row 0 (mz_yama.c:83-94) into the ring (first warp of the group) ----------------------
        if (G == 1 || lane < 32) {
            const int RB0 = rows[0].RB16 >> 4;
            const int RB1 = rows[0].RBn;
            int carry = 0;
            for (int base = 0; base <= RB1; base += 32) {
                int c = base + lane;
                int nd = 0;
                if (c >= 1 && c <= RB0) nd = (int)((__ldg(&cols[c].w0) >> 16) & 0xffu);
                int inc = nd;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int o = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= d) inc += o;
                }
                unsigned a = and_xor((unsigned)c << 4, (unsigned)(RING * 16 - 16), ringAddr);
                if (c <= RB0) {
                    int I0 = -(carry + inc) * KGE;
                    if (c == 0) sts128(a, 0, 0, 0, 0u);                    // (0,0): nothing is ever charged
                    else sts128(a, MININT, MININT, I0, pack16(0, nGO));    // only the I node exists in row 0
                } else if (c <= RB1) {
                    sts128(a, MININT, MININT, MININT, E_both);             // stale dp[] entry, mz_yama.c:93-94
                }
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
        }

        // ---- per-lane row state --------------------------------------------------------------------
        int r = lane + 1;
        unsigned avXC = 0, avYC = 0, avZC = 0, avXI = 0, avXD = 0, avYD = 0, avZD = 0;
        int gIrow = 0, gIz = 0;
        unsigned w01 = 0, w23 = 0, w45 = 0, Efirst = 0;
        int eD = 0, LB16 = 0x7fffffff, RB16 = 0x7fffffff, LBp16 = 0, c16 = 0;
        int rdMax16 = 0x7fffffff;                              // CTA-sized wavefronts: 16*(RB[r-1]+1), see the row-end code
        auto load_row = [&](int t) {
            const uint4 *rp = reinterpret_cast<const uint4 *>(rows + r);       // (recomputed: no pointer carried by the loop)
            uint4 q0 = __ldg(rp), q1 = __ldg(rp + 1), q2 = __ldg(rp + 2), q3 = __ldg(rp + 3);
            avXC = q0.x; avYC = q0.y; avZC = q0.z; avXI = q0.w;
            avXD = q1.x; avYD = q1.y; avZD = q1.z; eD = (int)q1.w;
            w01 = q2.x; w23 = q2.y; w45 = q2.z; LB16 = (int)q2.w;
            RB16 = (int)q3.x; LBp16 = (int)q3.y;
            c16 = (t - (int)q3.z) * 16;
            if (G > 1) rdMax16 = (int)q3.w;
            gIrow = r < M ? (Y16 ? (KnGO & 0xffff) : KnGO) : 0;     // mz_yama.c:123: no I-node gap-open on the last row
            gIz = Y16 ? gIrow << 16 : gIrow;                        // the z candidate's weight sits on byte 1 (b10)
            // first cell of the row: its I node never exists, its C node only if the band moved right
            Efirst = (LB16 > LBp16) ? E_c0 : 0u;
        };
        if (r <= M) load_row(0);
        unsigned acc = 0;
        const size_t tbWord = (size_t)lane * 2;               // this lane's word inside an 8-step group (see tb_byte)
        int Cl = MININT, Dl = MININT, Il = MININT, gCl = 0, gIl = 0;     // grid point (r, c-1)
        int Cd = MININT, Dd = MININT, Id = MININT, gCd = 0, gId = 0;     // grid point (r-1, c-1)
        group_sync();

        for (int t4 = 0; t4 < nSteps; t4 += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                // ---- grid point (r-1, c): read before anybody may overwrite it ------------------------
                const uint4 up = lds128(and_xor((unsigned)(G > 1 ? min(c16, rdMax16) : c16), rdMask, rdBase));
#ifdef YB_FORCE_WARPSYNC
                // sanitize build: the stores of this step (a finished row fills slots whose previous content its reader took just
                // above) come after every lane's read in program order; a barrier says so to the tool
                if (G == 1) group_sync();
#endif
                if (c16 > RB16) {
                    // ---- this lane finished its row -----------------------------------------------------
                    // (a) the row below keeps reading us up to its own right bound: stale dp[] entries (mz_yama.c:93-94).
                    //     A mailbox lane of a warp-sized wavefront fills both of its slots (the reader, in the same warp,
                    //     took our last column at the top of this step); its ring lane writes them column by column.
                    if (G > 1) {
                        // the reader may sit in another warp and may not have taken our last column yet: touch only
                        // the slot of column RB+1 (never the one of column RB); readers clamp their column to it
                        sts128(and_xor((unsigned)(RB16 + 16), wrMask, wrBase), MININT, MININT, MININT, E_both);
                    } else if (lane != B - 1) {
                        sts128(wrBase, MININT, MININT, MININT, E_both);
                        sts128(wrBase ^ 16u, MININT, MININT, MININT, E_both);
                    } else {
                        const int RBn = __ldg(reinterpret_cast<const int *>(rows + r) + 15);   // RowRec::RBn
#pragma unroll 1
                        for (int cc = (RB16 >> 4) + 1; cc <= RBn; ++cc)
                            sts128(and_xor((unsigned)cc << 4, wrMask, wrBase), MININT, MININT, MININT, E_both);
                    }
                    // (b) move one wavefront width down; a lane that runs out of rows idles, and the one that just
                    //     finished row M leaves the final scores (checked only on that rare path)
                    r += B;
                    if (r <= M) load_row(t4 + u);
                    else {
                        if (r - B == M) { outs[p].C = Cl; outs[p].D = Dl; outs[p].I = Il; }
                        LB16 = 0x7fffffff; RB16 = 0x7fffffff; rdMax16 = 0x7fffffff;
                    }
                }
                const int Cu = (int)up.x, Du = (int)up.y, Iu = (int)up.z;
                // GATED = false: every pair of the launch is small enough (host-side bound) that a candidate from a node
                // that does not exist -- exactly MININT -- can be charged like any other without ever winning a comparison
                // against a real one, so the existence multipliers need not travel with the records
                const int gCu = GATED ? (int)(short)(up.w & 0xffffu) : nGO, gIu = GATED ? ((int)up.w) >> 16 : nGO;
                const bool active = (c16 >= LB16);

                // column record of column c; lanes outside their row read a clamped (valid) column and discard the result
                const uint4 cw = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const unsigned char *>(cols) +
                                                                       (unsigned)__vimin_s32_relu(c16, N16)));

                // ---- I node (mz_yama.c:114-166) -----------------------------------------------------------
                int vI, vC, vD;
                acc >>= 8;                                   // make room for this cell's byte (bits 24..31)
                const bool hasI = c16 > LB16, hasC = c16 > LBp16;
                {
                    int x = Cl + dp4a_uu(cw.x, avXI, 0) * (GATED ? gCl : nGO);
                    int y, z;
                    if (Y16) {
                        y = dp2a_hi_su((unsigned)gIrow, cw.x, Dl);                    // K*ndB opens (mz_yama.c:131-134)
                        z = dp2a_lo_su((unsigned)(GATED ? gIl : gIz), cw.x, Il);       // K*b10, if I(r,c-1) exists
                    } else {
                        y = Dl + (int)__byte_perm(cw.x, 0, 0x4442) * gIrow;
                        z = Il + (int)__byte_perm(cw.x, 0, 0x4441) * (GATED ? gIl : gIz);
                    }
                    vI = pick3<4>(x, y, z, hasI, acc);
                    vI = dp2a_hi_su(nKGE_lo, cw.x, vI);            // - ndB*K*gap_ext (mz_yama.c:158-161)
                }
                vI = hasI ? vI : MININT;
                // ---- C node (mz_yama.c:169-205) -----------------------------------------------------------
                {
                    int x = Cd + dp4a_uu(cw.w, avXC, 0) * (GATED ? gCd : nGO);
                    int y = Y16 ? dp2a_hi_su(avYC, cw.w, Dd) : Dd + dp4a_uu(cw.w, avYC, 0) * nGO;
                    int z = Id + dp4a_uu(cw.w, avZC, 0) * (GATED ? gId : nGO);
                    vC = pick3<0>(x, y, z, hasC, acc);
                    vC = dp2a_lo_su(w01, cw.y, vC);
                    vC = dp2a_hi_su(w23, cw.y, vC);
                    vC = dp2a_lo_su(w45, cw.z, vC);
                }
                vC = hasC ? vC : MININT;
                // ---- D node (mz_yama.c:208-242) -----------------------------------------------------------
                {
                    int x = Cu + dp4a_uu(cw.z, avXD, 0) * gCu;
                    int y = Y16 ? dp2a_hi_su(avYD, cw.z, Du) : Du + dp4a_uu(cw.z, avYD, 0) * nGO;
                    int z = Iu + dp4a_uu(cw.z, avZD, 0) * gIu;
                    vD = pick3<2>(x, y, z, true, acc) - eD;
                }
                if (active) {
                    sts128(and_xor((unsigned)c16, wrMask, wrBase), vC, vD, vI, GATED ? (hasI ? E_both : Efirst) : 0u);
                }
                // four steps of this lane = one 32-bit word of its 8-step group (see tb_byte)
                if (u == 3) reinterpret_cast<unsigned *>(tb)[(size_t)(t4 >> 3) * (2 * B) + tbWord + ((t4 >> 2) & 1)] = acc;
                Cl = vC; Dl = vD; Il = vI;
                if (GATED) { gCl = hasC ? nGO : 0; gIl = hasI ? gIz : 0; gCd = gCu; gId = gIu; }
                Cd = Cu; Dd = Du; Id = Iu;
                c16 += 16;
                group_sync();
            }
        }
        group_sync();
    }
}

// =================================================================================================
// K2, bulk form (fill_body2): one warp per pair, for pairs of kernel class 1 and 2 (PairMeta::cls; the host admits a pair
// when its band is connected, its scores are bounded -- the GATED = false argument of fill_body, DESIGN section 2 -- and
// every pre-multiplied weight below fits 16 bits).  Same wavefront and schedule as fill_body<RING,1,P>, rebuilt around
// what ncu showed of that kernel (profiles/r1_fill_summary.md: 75 % issue slots, L1/shared data pipe 72 %, long-scoreboard
// stalls on the row records, 89 issue slots per cell):
//
//  * (C,D,I) of the row above arrive by three warp shuffles from lane l-1's registers instead of a shared-memory mailbox
//    (LDS.128 + STS.128 per lane and step = 8 crossbar cycles per warp step; three SHFL = 3).  Only lane 0 reads and lane 31
//    writes the ring that carries a band row from one 32-row block to the next.  A lane that is not inside its row hands
//    down MININT -- exactly the never-written dp[] entries of mz_yama.c:93-94 -- so a finished row needs no stale stores.
//  * a lane's NEXT row record is copied global -> shared (cp.async, a private 64-B slot per lane) while it walks its
//    current row; the row switch reads it with four LDS.128 instead of four dependent global loads.
//  * records in the bulk layout (RowRec2 / ColRec2): every two-term gap-open count is ONE dp2a with 16-bit weights already
//    multiplied by -gap_open, accumulated straight onto the predecessor's value; the I node's extension charge rides on its
//    three candidates (max(x,y,z) - e = max(x-e, y-e, z-e)); the class "other" is eliminated from the sum-of-pairs dot
//    product (n_X = ndB - n_A - n_C - n_G - n_T), which frees the column bytes the pairs need.
//  * KEYED (class 2): node values are carried as 4*value + p with p = 2 for a C node, 1 for an I node, 0 for a D node, and
//    every weight is a multiple of 4.  A candidate inherits the p of the node it comes from, so ONE 3-way maximum decides
//    value and tie-break together: 4x+2 > 4y and 4x+2 > 4z+1 iff x >= y and x >= z (from-C wins ties), 4y > 4z+1 iff y > z
//    (from-D beats from-I only strictly) -- the rule of mz_yama.c:138-154 -- and the low two bits of the maximum ARE the
//    traceback pointer (2: from C, 0: from D, 1: from I): one funnel shift per node into the packed word instead of two
//    compares and two predicated adds.  Needs real scores below 2^26 (host-checked per pair).
// =================================================================================================
// Column record of the bulk layout, 16 B.  "m": zeroed for c == 1 (mz_yama.c:173, no gap-open at the start);
// ndB', dB': zeroed for c == 0 and c == N (mz_yama.c:211, end gaps free).
//  w0 = (n_A, n_C, n_G, n_T)            class counts
//  w1 = (ndB, dB, ndB, b10)             lo: sum-of-pairs terms of "other" and dash; hi: the I node's pair
//  w2 = (b01, b10, ndB, dB) m           C node: x (all four, dp4a), y (hi)
//  w3 = (dB m, b10 m, ndB', dB')        lo: C node z; hi: D node x, y, z
// Row record of the bulk layout, 64 B.  p16(lo, hi) are 16-bit weights, g = -gap_open * scale, e = K * gap_extend * scale.
struct __align__(16) RowRec2 {
    unsigned avXC;      // bytes (a00, a11, a01, a10), row > 1
    unsigned wYC;       // p16(dA*g, a10*g) on (ndB, dB), row > 1
    unsigned wZC;       // p16(ndA*g, dA*g) on (dB, b10)
    unsigned wXI;       // p16(ndA*g - e, dA*g) on (ndB, b10); row M: p16(-e, 0)  (mz_yama.c:123)
    unsigned wXD;       // p16(ndA*g, a10*g) on (ndB', dB'), row > 1
    unsigned wYD;       // p16(a10*g, a10*g), row > 1
    unsigned wZD;       // p16(ndA*g, ndA*g)
    int eD;             // ndA * L * gap_extend * scale  (mz_yama.c:239-242)
    unsigned w01, w23;  // p16 of S6^T * classcount(A row) minus its "other" entry, classes A C | G T, times scale
    unsigned w45;       // p16(other, dash) entries, times scale, on (ndB, dB)
    int LB16;           // 16*LB[r]
    int RB16;           // 16*RB[r]
    int LBc16;          // the C node of (r,c) exists iff 16*c > LBc16 = max(16*LB[r] - 16, 16*LB[r-1])
    int off16;          // wavefront schedule: this row computes column (step - off), times 16
    int RBn;            // RB[r+1] (RB[r] on the last row): how far the row below reads us
};
static_assert(sizeof(RowRec2) == sizeof(RowRec), "both layouts share the row pool");

__device__ __forceinline__ void cp_async16(unsigned dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// A KEYED node value back in the reference's units.  Values of nodes no alignment path reaches sit around the keyed
// sentinel -2^30 = 4 * (-2^28): they are moved to where the reference has them, around MININT = -2^30 (the drift of such a
// value away from its sentinel is kept: it is the reference's "MININT - penalty", mz_yama.c:93-94).
template <bool KEYED>
__device__ __forceinline__ int unkey(int v) {
    if (!KEYED) return v;
    const int u = v >> 2;
    return v < -(1 << 29) ? u - 3 * (1 << 28) : u;
}

// K1 for a pair of class 1 / 2: the same census, written in the bulk layout (see RowRec2 above)
__device__ __forceinline__ void profile_bulk(const PairMeta &pm, const unsigned char *__restrict__ blob, RowRec *__restrict__ rowPool,
                                             ColRec *__restrict__ colPool, const ScoreConst &c_sc) {
    const int K = pm.K, M = pm.M, L = pm.L, N = pm.N;
    const int scale = pm.cls == 2 ? 4 : 1;
    const unsigned char *A = blob + pm.offA;
    const unsigned char *B = blob + pm.offB;
    const BandView band(blob, pm);
    RowRec2 *rows = reinterpret_cast<RowRec2 *>(rowPool + pm.rowBase);
    ColRec *cols = colPool + pm.colBase;
    const int *sched = reinterpret_cast<const int *>(blob + pm.offSched);
    const int g = -c_sc.gap_open * scale, e = K * c_sc.gap_ext * scale;
    for (int c = threadIdx.x; c <= N; c += K1_THREADS) {
        ColRec cr = {0u, 0u, 0u, 0u};
        if (c >= 1) {
            const unsigned char *now = B + (size_t)(c - 1) * L;
            const Census q = census(now, c > 1 ? now - L : nullptr, L);    // mz_yama.c:128 (t==0 when col==1)
            const unsigned *n = q.n;
            const unsigned b01 = q.t01, b10 = q.t10, dB = n[5], ndB = (unsigned)L - dB;
            const bool inner = (c < N), later = (c > 1);                   // mz_yama.c:211, :173
            cr.w0 = pack4(n[0], n[1], n[2], n[3]);
            cr.w1 = pack4(ndB, dB, ndB, b10);
            cr.w2 = later ? pack4(b01, b10, ndB, dB) : 0u;
            cr.w3 = pack4(later ? dB : 0u, later ? b10 : 0u, inner ? ndB : 0u, inner ? dB : 0u);
        }
        cols[c] = cr;
    }
    for (int r = threadIdx.x; r <= M; r += K1_THREADS) {
        RowRec2 rr;
        rr.avXC = rr.wYC = rr.wZC = rr.wXI = rr.wXD = rr.wYD = rr.wZD = 0u;
        rr.eD = 0; rr.w01 = rr.w23 = rr.w45 = 0u;
        const int lb = band.lb(r), lbp = r > 0 ? band.lb(r - 1) : 0;
        rr.LB16 = lb * 16; rr.RB16 = band.rb(r) * 16;
        rr.LBc16 = max(lb * 16 - 16, lbp * 16);
        rr.off16 = r >= 1 ? 16 * (sched[(r - 1) >> 5] + ((r - 1) & 31)) : 0;
        rr.RBn = r < M ? band.rb(r + 1) : band.rb(r);
        if (r >= 1) {
            const unsigned char *now = A + (size_t)(r - 1) * K;
            const Census q = census(now, r > 1 ? now - K : nullptr, K);    // mz_yama.c:175,213 (s==0 when row==1)
            int n[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) n[k] = (int)q.n[k];
            const int a00 = (int)q.t00, a01 = (int)q.t01, a10 = (int)q.t10, a11 = (int)q.t11;
            const int dA = n[5], ndA = K - dA;
            // gap-open counts:  C.x: a00*b01 + a11*b10 + a01*ndB + a10*dB   C.y: dA*ndB + a10*dB   C.z: ndA*dB + dA*b10
            //   I.x: ndA*ndB + dA*b10   I.y: K*ndB   I.z: K*b10   D.x: ndA*ndB' + a10*dB'   D.y: a10*(ndB'+dB')   D.z: ndA*(ndB'+dB')
            if (r > 1) {                                        // mz_yama.c:180-184, :218-221 (row>1)
                rr.avXC = pack4((unsigned)a00, (unsigned)a11, (unsigned)a01, (unsigned)a10);
                rr.wYC = pack16(dA * g, a10 * g);
                rr.wXD = pack16(ndA * g, a10 * g);
                rr.wYD = pack16(a10 * g, a10 * g);
            }
            rr.wZC = pack16(ndA * g, dA * g);
            rr.wZD = pack16(ndA * g, ndA * g);
            rr.wXI = r < M ? pack16(ndA * g - e, dA * g) : pack16(-e, 0);      // mz_yama.c:123 (row<M), :158-161
            rr.eD = ndA * L * c_sc.gap_ext * scale;
            int w[6];
#pragma unroll
            for (int l = 0; l < 6; ++l) {
                int acc = 0;
#pragma unroll
                for (int k = 0; k < 6; ++k) acc += n[k] * c_sc.S6[k][l];
                w[l] = acc * scale;
            }
            // n_X = ndB - n_A - n_C - n_G - n_T: the "other" weight moves onto ndB
            rr.w01 = pack16(w[0] - w[4], w[1] - w[4]); rr.w23 = pack16(w[2] - w[4], w[3] - w[4]); rr.w45 = pack16(w[4], w[5]);
        }
        rows[r] = rr;
    }
}

#ifndef YB_F2_WARPS
#define YB_F2_WARPS 8
#endif
// YB_F2_SW: a lane that has finished its row moves to its next row at the next step that is a multiple of YB_F2_SW (1, 2, 4
// or 8) instead of at once.  The row switch is the one divergent region of the kernel -- some forty instructions executed for
// a single lane, about every second step of a warp whose 32 rows end one step apart -- and with YB_F2_SW = 4 up to four lanes
// share one pass through it.  The price: a lane idles up to YB_F2_SW - 1 steps past its row (handing down MININT, as a lane
// outside its row always does), so the schedule leaves that many more steps between blocks of rows (PlanParams::slackBulk =
// 3 + YB_F2_SW - 1), one more compare per step (c <= RB[r]), and the final scores are picked up where cell (M,N) is computed
// -- in one of the last two 8-step groups, which run a copy of the step code with that check -- not at the switch.
// Measured on cfg2 (profiles/r2_ab_sw.txt): 1 -> 5.69 ms, 2 -> 5.98 ms, 4 -> 5.44 ms per fill.
#ifndef YB_F2_SW
#define YB_F2_SW 4
#endif
constexpr int F2_SW = YB_F2_SW;
constexpr int F2_WARPS = YB_F2_WARPS;                         // pairs in flight per CTA
constexpr size_t COL_PAD = 32768;                             // bytes of slack before and after a wave's column records: lanes outside
                                                              // their row read (and discard) up to about two band rows off either end

struct FalseTag { static constexpr bool value = false; };
struct TrueTag { static constexpr bool value = true; };

template <int RING, bool KEYED>
__device__ __forceinline__ void
fill_body2(const PairMeta *__restrict__ metas, const int *__restrict__ orderBase, const int *__restrict__ binRange,
           int *__restrict__ queue, const RowRec *__restrict__ rowPool,
           const ColRec *__restrict__ colPool, unsigned char *__restrict__ tbPool,
           const unsigned long long *__restrict__ tbBase, PairOut *__restrict__ outs, const int nGO, const int gapExt) {
    // nGO: -gap_open * SC, from the host (a kernel parameter is an operand, not an instruction)
    const int *order = orderBase + __ldg(binRange);                // the bin's slice of the launch order, as K0 left it
    const int nPairs = __ldg(binRange + 1) - __ldg(binRange);
    constexpr int B = 32, P = F2_WARPS;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int SC = KEYED ? 4 : 1;                         // every weight is multiplied by SC
    constexpr int PRC = KEYED ? 2 : 0, PRI = KEYED ? 1 : 0;   // low bits of a C / I node value (a D node carries 0)
    constexpr int MIN_C = MININT | PRC, MIN_D = MININT, MIN_I = MININT | PRI;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [ P rings of RING records | P x 32 row slots of 64 B | P stash words | pad ]; rings are RING*16-aligned
    const int grp = (int)(threadIdx.x >> 5);
    const int lane = (int)(threadIdx.x & 31);
    const unsigned smem0 = (smem_u32(smem_raw) + RING * 16 - 1) & ~(unsigned)(RING * 16 - 1);
    const unsigned ringAddr = smem0 + grp * RING * 16;
    const unsigned rowSlot = smem0 + P * RING * 16 + (unsigned)(grp * B + lane) * 64u;
    const unsigned stash = smem0 + P * RING * 16 + P * B * 64 + grp * 16;      // lane 31's RB[r+1], see the row switch
    const unsigned keyMask = launder(~3u);                    // (in a register: LOP3 takes one immediate)
    constexpr unsigned RMASK = (unsigned)(RING * 16 - 16);

    for (;;) {
        int slot = 0;
        if (lane == 0) slot = atomicAdd(queue, 1);
        slot = __shfl_sync(FULL, slot, 0);
        if (slot >= nPairs) break;
        const int p = order[slot];
        const PairMeta pm = metas[p];
        const int M = pm.M;
        const RowRec2 *rows = reinterpret_cast<const RowRec2 *>(rowPool + pm.rowBase);
        const ColRec *cols = colPool + pm.colBase;
        unsigned char *tb = tbPool + __ldg(tbBase + p);
        const int KGE = pm.K * gapExt * SC;
        const int nSteps = pm.nSteps;
        // the I node's y and z weights on (ndB, b10): K*ndB opens and K*b10 opens (mz_yama.c:131-137), extension folded in;
        // on the last row only the extension is charged (mz_yama.c:123)
        const unsigned cYI = pack16(pm.K * nGO - KGE, 0), cZI = pack16(-KGE, pm.K * nGO), cLast = pack16(-KGE, 0);

        // ---- row 0 (mz_yama.c:83-94) into the ring -------------------------------------------------------------
        {
            const int RB0 = rows[0].RB16 >> 4;
            const int RB1 = rows[0].RBn;
            int carry = 0;
            for (int base = 0; base <= RB1; base += 32) {
                int c = base + lane;
                int nd = 0;
                if (c >= 1 && c <= RB0) nd = (int)(__ldg(&cols[c].w1) & 0xffu);          // ndB
                int inc = nd;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int o = __shfl_up_sync(FULL, inc, d);
                    if (lane >= d) inc += o;
                }
                unsigned a = and_xor((unsigned)c << 4, RMASK, ringAddr);
                if (c <= RB0) {
                    int I0 = -(carry + inc) * KGE + PRI;
                    if (c == 0) sts128(a, PRC, 0, PRI, 0u);                    // (0,0): all three nodes are 0
                    else sts128(a, MIN_C, MIN_D, I0, 0u);                      // only the I node exists in row 0
                } else if (c <= RB1) {
                    sts128(a, MIN_C, MIN_D, MIN_I, 0u);                        // stale dp[] entry, mz_yama.c:93-94
                }
                carry += __shfl_sync(FULL, inc, 31);
            }
        }

        // ---- per-lane row state -----------------------------------------------------------------------------------
        int r = lane + 1;
        unsigned avXC = 0, wYC = 0, wZC = 0, wXI = 0, wXD = 0, wYD = 0, wZD = 0, wYI = 0, wZI = 0;
        unsigned w01 = 0, w23 = 0, w45 = 0;
        // LBst16: LB16 on the lane that feeds the ring, never reached on the others (one compare decides the ring store)
        int eD = 0, LB16 = 0x7fffffff, RB16 = 0x7fffffff, LBc16 = 0x7fffffff, LBst16 = 0x7fffffff, c16 = 0;
        // cp: address of this lane's column record at the first step of the current group of eight (the loads of the group
        // use immediate offsets; a row switch re-bases it).  Lanes outside their row read whatever lies there -- the column
        // pool is padded on both sides (COL_PAD) -- and discard it.
        const unsigned char *const colBase = reinterpret_cast<const unsigned char *>(cols);
        const unsigned char *cp = colBase;
        const unsigned char *pf = reinterpret_cast<const unsigned char *>(rows + r + B);     // the record the next switch prefetches
        auto unpack_row = [&](const uint4 &q0, const uint4 &q1, const uint4 &q2, const uint4 &q3, int t16, int u) {
            avXC = q0.x; wYC = q0.y; wZC = q0.z; wXI = q0.w;
            wXD = q1.x; wYD = q1.y; wZD = q1.z; eD = (int)q1.w;
            w01 = q2.x; w23 = q2.y; w45 = q2.z; LB16 = (int)q2.w;
            RB16 = (int)q3.x; LBc16 = (int)q3.y;
            c16 = t16 - (int)q3.z;
            cp = colBase + (c16 - 16 * u);
            LBst16 = (lane == B - 1) ? LB16 : 0x7fffffff;
            wYI = r < M ? cYI : cLast;
            wZI = r < M ? cZI : cLast;
            if (lane == B - 1) asm volatile("st.shared.u32 [%0], %1;" ::"r"(stash), "r"(q3.w) : "memory");
        };
        auto prefetch_row = [&]() {                                 // the record at pf -> this lane's slot
            cp_async16(rowSlot, pf); cp_async16(rowSlot + 16, pf + 16);
            cp_async16(rowSlot + 32, pf + 32); cp_async16(rowSlot + 48, pf + 48);
            cp_async_commit();
            pf += B * sizeof(RowRec2);
        };
        if (r <= M) {
            const uint4 *rp = reinterpret_cast<const uint4 *>(rows + r);
            const uint4 q0 = __ldg(rp), q1 = __ldg(rp + 1), q2 = __ldg(rp + 2), q3 = __ldg(rp + 3);
            unpack_row(q0, q1, q2, q3, 0, 0);
            if (r + B <= M) prefetch_row();
        }
        unsigned acc = 0;
        unsigned *tbp = reinterpret_cast<unsigned *>(tb) + lane * 2;      // this lane's two words of an 8-step group (see tb_byte)
        int Cl = MIN_C, Dl = MIN_D, Il = MIN_I;               // grid point (r, c-1); also what this lane hands down
        int Cd = MIN_C, Dd = MIN_D, Id = MIN_I;               // grid point (r-1, c-1)
        __syncwarp();

        // one group of eight steps; CHECK: this group may hold cell (M,N) (F2_SW > 1 only)
        auto group = [&](const int t8, auto checkTag) {
            constexpr bool CHECK = decltype(checkTag)::value;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                // ---- grid point (r-1, c): the last cell of the lane above, or MININT if it is outside its row --------
                int Cu = __shfl_up_sync(FULL, Cl, 1), Du = __shfl_up_sync(FULL, Dl, 1), Iu = __shfl_up_sync(FULL, Il, 1);
                if (lane == 0) {
                    const uint4 up = lds128(and_xor((unsigned)c16, RMASK, ringAddr));
                    Cu = (int)up.x; Du = (int)up.y; Iu = (int)up.z;
                }
#ifdef YB_FORCE_WARPSYNC
                asm volatile("bar.sync %0, 32;" ::"r"(grp + 1) : "memory");     // (sanitize build: this step's ring stores follow the read)
#endif
                if ((F2_SW == 1 || (u % F2_SW) == 0) && c16 > RB16) {
                    // ---- this lane finished its row: move one wavefront width down -----------------------------------
                    // the row below keeps reading us up to its own right bound and must find never-written dp[] entries
                    // there (mz_yama.c:93-94): an idle lane hands down MININT by itself, the ring needs them written
                    if (lane == B - 1) {
                        int RBn;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(RBn) : "r"(stash) : "memory");
                        // (F2_SW > 1: the columns it idled over since the row ended were written step by step, below)
#pragma unroll 1
                        for (int cc = (F2_SW == 1 ? RB16 + 16 : c16) >> 4; cc <= RBn; ++cc)
                            sts128(and_xor((unsigned)cc << 4, RMASK, ringAddr), MIN_C, MIN_D, MIN_I, 0u);
                    }
                    r += B;
                    if (r <= M) {
                        cp_async_wait_all();
                        const uint4 q0 = lds128(rowSlot), q1 = lds128(rowSlot + 16), q2 = lds128(rowSlot + 32), q3 = lds128(rowSlot + 48);
                        unpack_row(q0, q1, q2, q3, (t8 + u) * 16, u);
                        if (r + B <= M) prefetch_row();               // (after the slot's words were consumed)
                    } else {
                        // a lane that runs out of rows idles; the one that just finished row M leaves the final scores
                        if (F2_SW == 1 && r - B == M) { outs[p].C = unkey<KEYED>(Cl); outs[p].D = unkey<KEYED>(Dl); outs[p].I = unkey<KEYED>(Il); }
                        LB16 = 0x7fffffff; RB16 = 0x7fffffff; LBc16 = 0x7fffffff; LBst16 = 0x7fffffff;
                        cp = colBase - 16 * u;                        // (idles over the pair's first columns)
                    }
                }
                // F2_SW > 1: a lane may sit past the end of its row for a few steps; it is outside its row there
                const bool in = F2_SW == 1 || c16 <= RB16;
                const bool active = (c16 >= LB16) && in;
                const uint4 cw = __ldg(reinterpret_cast<const uint4 *>(cp + 16 * u));       // column record of column c
                int vI, vC, vD;
                const bool hasI = (c16 > LB16) && in, hasC = (c16 > LBc16) && in;
                if (!KEYED) acc >>= 8;                       // make room for this cell's byte (bits 24..31)
                // ---- C node (mz_yama.c:169-205) -----------------------------------------------------------
                {
                    const int x = Cd + dp4a_uu(cw.z, avXC, 0) * nGO;
                    const int y = dp2a_hi_su(wYC, cw.z, Dd);
                    const int z = dp2a_lo_su(wZC, cw.w, Id);
                    if (KEYED) {
                        const int m = __vimax3_s32(x, y, z);
                        acc = __funnelshift_r(acc, (unsigned)m, 2);
                        vC = (int)and_or((unsigned)m, keyMask, (unsigned)PRC);
                    } else vC = pick3<0>(x, y, z, hasC, acc);
                    vC = dp2a_lo_su(w01, cw.x, vC);
                    vC = dp2a_hi_su(w23, cw.x, vC);
                    vC = dp2a_lo_su(w45, cw.y, vC);
                }
                vC = hasC ? vC : MIN_C;
                // ---- D node (mz_yama.c:208-242) -----------------------------------------------------------
                {
                    const int x = dp2a_hi_su(wXD, cw.w, Cu);
                    const int y = dp2a_hi_su(wYD, cw.w, Du);
                    const int z = dp2a_hi_su(wZD, cw.w, Iu);
                    if (KEYED) {
                        const int m = __vimax3_s32(x, y, z);
                        acc = __funnelshift_r(acc, (unsigned)m, 2);
                        vD = (int)((unsigned)m & keyMask) - eD;
                    } else vD = pick3<2>(x, y, z, true, acc) - eD;
                }
                vD = active ? vD : MIN_D;
                // ---- I node (mz_yama.c:114-166) -----------------------------------------------------------
                {
                    const int x = dp2a_hi_su(wXI, cw.y, Cl);
                    const int y = dp2a_hi_su(wYI, cw.y, Dl);
                    const int z = dp2a_hi_su(wZI, cw.y, Il);
                    if (KEYED) {
                        const int m = __vimax3_s32(x, y, z);
                        acc = __funnelshift_r(acc, (unsigned)m, 4);                 // (bits 6,7 of the byte: don't care)
                        vI = (int)and_or((unsigned)m, keyMask, (unsigned)PRI);
                    } else vI = pick3<4>(x, y, z, hasI, acc);
                }
                vI = hasI ? vI : MIN_I;
                if (c16 >= LBst16) sts128(and_xor((unsigned)c16, RMASK, ringAddr), vC, vD, vI, 0u);
                if (F2_SW > 1 && CHECK && c16 == RB16 && r == M) {     /* cell (M,N): RB[M] == N */ outs[p].C = unkey<KEYED>(vC); outs[p].D = unkey<KEYED>(vD); outs[p].I = unkey<KEYED>(vI); }
                // four steps of this lane = one 32-bit word of its 8-step group (see tb_byte)
                if (u == 3) tbp[0] = acc;
                if (u == 7) tbp[1] = acc;
                Cl = vC; Dl = vD; Il = vI;
                Cd = Cu; Dd = Du; Id = Iu;
                c16 += 16;
                // The ring is the one shared-memory hand-off left (lane 31 writes a band row, lane 0 reads it a block of rows
                // later; the row slots and the stash word are private to a lane).  Its write and its read are always in
                // different steps, and the steps of a warp are separated by the three full-mask shuffles above, which no lane
                // passes before all have arrived: the LDS of a later step is issued after the STS of an earlier one, and shared
                // memory serves one warp's accesses in issue order.  (A __syncwarp() here is elided by ptxas -- the warp is
                // converged -- but its compiler barrier keeps the next step's column load from being hoisted: not shipped.)
                // The `make sanitize` build separates the steps with a named 32-thread barrier, which is never elided and
                // which compute-sanitizer's racecheck can see: same protocol, tool-checked (profiles/r2_sanitizer.md).
#ifdef YB_FORCE_WARPSYNC
                asm volatile("bar.sync %0, 32;" ::"r"(grp + 1) : "memory");
#endif
            }
            cp += 128;
            tbp += 2 * B;
        };
        {                                                      // (nSteps is a multiple of 8; cell (M,N) lies in its last 9 steps)
            int t8 = 0;
            if (F2_SW > 1) {
                for (; t8 < nSteps - 16; t8 += 8) group(t8, FalseTag{});
                for (; t8 < nSteps; t8 += 8) group(t8, TrueTag{});
            } else {
                for (; t8 < nSteps; t8 += 8) group(t8, FalseTag{});
            }
        }
        __syncwarp();
    }
}

// =================================================================================================
// K2, wide bands (fill_body3): one CTA of G warps per pair, for band rows of hundreds to thousands of cells on long pairs
// (BASELINE configs[4]: K+L = 100, M = 10^4, R = 300) -- pairs whose scores approach the int32 range, so the arithmetic is
// fill_body's (byte counts x gap_open, existence multipliers), while the communication is rebuilt around what ncu showed of
// fill_body<2048,8,1> on cfg5 (profiles/r2_fill_cta_summary.md: 52 % of the issue slots, one bar.sync of 256 threads and one
// LDS.128 + STS.128 per lane and step):
//
//  * inside a warp (C,D,I) of the row above arrive by three shuffles, as in fill_body2; the existence multipliers of a grid
//    point are not handed down at all: the C node of (r-1,c) exists iff c > LB[r-2], its I node iff c > LB[r-1]
//    (mz_yama.c:163-165, :202-204), two compares against numbers the row record carries;
//  * the warps of a CTA are NOT in lock step.  Lane 31 of warp w leaves its (C,D,I) of every step in a small FIFO indexed by
//    the step; lane 0 of warp w+1 picks up the entry of eight steps earlier -- the schedule puts F3_SKEW = 7 extra steps
//    between the two lanes (PairMeta::skew; K0, K1 and K3 know) -- so a warp needs its predecessor to have FINISHED the
//    previous group of eight steps, not to stand at the same step.  Each warp publishes the number of groups it has completed
//    (one shared-memory word, written by lane 31 after a fence); before a group, lane 0 waits until the warp it reads from is
//    far enough and the warp that reads it is not more than F3_BP groups behind (the FIFO and the ring are finite).  No CTA
//    barrier inside a pair; warps drift a few groups apart and hide each other's row switches and column loads;
//  * the last warp hands a band row to the first through the ring, indexed by the column as ever; the first warp works out
//    from the schedule how many groups it may run ahead of the last (the distance in steps between a block of rows and the
//    next is known when lane 0 switches rows);
//  * a lane's next row record is copied global -> shared (cp.async) while it walks its current row.
// =================================================================================================
constexpr int F3_SKEW = 7;          // steps between lane 31 of a warp and lane 0 of the next, beyond the usual one
constexpr int F3_FIFO = 128;        // entries of an inter-warp FIFO: 16 groups of eight steps
constexpr int F3_BP = F3_FIFO / 8 - 3;   // groups a warp may run ahead of the warp that reads it (the FIFO's groups minus the group
                                    // being written, the group being read and the skew)
#ifndef YB_F3_SLEEP
#define YB_F3_SLEEP 64
#endif
// ring entries a band row of `wmax` cells needs: the row, the distance in columns between the lane that writes the ring and
// the lane that reads it (B - 1 lanes + the skews), the drift allowed between their warps, and a margin
__host__ __device__ constexpr int f3_ring_need(int wmax, int G) { return wmax + 32 * G + F3_SKEW * (G - 1) + 8 * (F3_BP + 1) + 16; }
// steps a block of rows starts after the block above, at least: lane 0 reads the ring a finished group (8 steps) after it was written
__host__ __device__ constexpr int f3_min_advance(int G) { return 32 * G + F3_SKEW * (G - 1) + 7; }

__device__ __forceinline__ int ld_volatile_shared(unsigned addr) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_shared(unsigned addr, int v) {
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

template <int RING, int G, bool Y16>
__device__ __forceinline__ void
fill_body3(const PairMeta *__restrict__ metas, const int *__restrict__ orderBase, const int *__restrict__ binRange,
           int *__restrict__ queue, const RowRec *__restrict__ rowPool,
           const ColRec *__restrict__ colPool, unsigned char *__restrict__ tbPool,
           const unsigned long long *__restrict__ tbBase, PairOut *__restrict__ outs, const int gapOpen, const int gapExt) {
    const int *order = orderBase + __ldg(binRange);
    const int nPairs = __ldg(binRange + 1) - __ldg(binRange);
    // (the pair descriptors sit at the start of the wave's input blob: the band rows are reached through them)
    const unsigned char *blob = reinterpret_cast<const unsigned char *>(metas);
    constexpr int B = 32 * G;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr unsigned RMASK = (unsigned)(RING * 16 - 16);
    constexpr unsigned FMASK = (unsigned)(F3_FIFO * 16 - 16);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [ ring: RING records | G FIFOs of F3_FIFO records (FIFO w is written by warp w) | B row slots of 64 B | G progress words | queue slot ]
    const int w = (int)(threadIdx.x >> 5);
    const int lane = (int)(threadIdx.x & 31);
    const int l = (int)threadIdx.x;                               // lane within the pair's wavefront
    const unsigned ringAddr = smem_u32(smem_raw);
    const unsigned fifoW = ringAddr + RING * 16 + (unsigned)w * (F3_FIFO * 16);
    const unsigned fifoR = ringAddr + RING * 16 + (unsigned)((w + G - 1) % G) * (F3_FIFO * 16);
    const unsigned rowSlot = ringAddr + RING * 16 + G * F3_FIFO * 16 + (unsigned)l * 64u;
    const unsigned doneAddr = ringAddr + RING * 16 + G * F3_FIFO * 16 + B * 64;
    const unsigned slotAddr = doneAddr + G * 4;
    const unsigned donePrev = doneAddr + 4u * (unsigned)((w + G - 1) % G), doneNext = doneAddr + 4u * (unsigned)((w + 1) % G);
    const int nGO = -gapOpen;

    for (;;) {
        int slot;
        if (l == 0) st_volatile_shared(slotAddr, atomicAdd(queue, 1));
        if (l < G) st_volatile_shared(doneAddr + 4u * l, 0);
        __syncthreads();
        slot = ld_volatile_shared(slotAddr);
        if (slot >= nPairs) break;
        const int p = order[slot];
        const PairMeta pm = metas[p];
        const int M = pm.M;
        const RowRec *rows = rowPool + pm.rowBase;
        const ColRec *cols = colPool + pm.colBase;
        const int *rbArr = reinterpret_cast<const int *>(blob + pm.offBand2);
        unsigned char *tb = tbPool + __ldg(tbBase + p);
        const unsigned nKGE_lo = launder((unsigned)(-(pm.K * gapExt)) & 0xffffu);   // dp2a.hi weight of byte 2 (ndB)
        const int KGE = pm.K * gapExt;
        const int nSteps = pm.nSteps;
        const int N16 = pm.N * 16;
        const int KnGO = pm.K * nGO;                        // I-node y / z charge per residue / per closing gap of B

        // ---- row 0 (mz_yama.c:83-94) into the ring (first warp) --------------------------------------------------------
        if (w == 0) {
            const int RB0 = rows[0].RB16 >> 4;
            const int RB1 = __ldg(rbArr + 1);
            int carry = 0;
            for (int base = 0; base <= RB1; base += 32) {
                int c = base + lane;
                int nd = 0;
                if (c >= 1 && c <= RB0) nd = (int)((__ldg(&cols[c].w0) >> 16) & 0xffu);
                int inc = nd;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int o = __shfl_up_sync(FULL, inc, d);
                    if (lane >= d) inc += o;
                }
                const unsigned a = ringAddr + (((unsigned)c << 4) & RMASK);
                if (c <= RB0) {
                    int I0 = -(carry + inc) * KGE;
                    if (c == 0) sts128(a, 0, 0, 0, 0u);                    // (0,0): nothing is ever charged
                    else sts128(a, MININT, MININT, I0, 0u);                // only the I node exists in row 0
                } else if (c <= RB1) {
                    sts128(a, MININT, MININT, MININT, 0u);                 // stale dp[] entry, mz_yama.c:93-94
                }
                carry += __shfl_sync(FULL, inc, 31);
            }
        }

        // ---- per-lane row state ------------------------------------------------------------------------------------------
        int r = l + 1;
        unsigned avXC = 0, avYC = 0, avZC = 0, avXI = 0, avXD = 0, avYD = 0, avZD = 0;
        int gIrow = 0, gIz = 0;
        unsigned w01 = 0, w23 = 0, w45 = 0;
        int eD = 0, LB16 = 0x7fffffff, RB16 = 0x7fffffff, LBp16 = 0x7fffffff, LBpp16 = 0x7fffffff, c16 = 0, off16 = 0;
        // the lane that leaves its values for another warp: every step into the FIFO (warps 0..G-2), its row into the ring (warp G-1)
        int LBst16 = 0x7fffffff;
        int needAdj = -(1 << 24);                              // warp 0: groups it must stay behind warp G-1, see the row switch
        const unsigned char *pf = reinterpret_cast<const unsigned char *>(rows + r + B);     // the record the next switch prefetches
        auto unpack_row = [&](const uint4 &q0, const uint4 &q1, const uint4 &q2, const uint4 &q3, int t16) {
            avXC = q0.x; avYC = q0.y; avZC = q0.z; avXI = q0.w;
            avXD = q1.x; avYD = q1.y; avZD = q1.z; eD = (int)q1.w;
            w01 = q2.x; w23 = q2.y; w45 = q2.z; LB16 = (int)q2.w;
            RB16 = (int)q3.x; LBp16 = (int)q3.y; off16 = (int)q3.z; LBpp16 = (int)q3.w;
            c16 = t16 - off16;
            gIrow = r < M ? (Y16 ? (KnGO & 0xffff) : KnGO) : 0;     // mz_yama.c:123: no I-node gap-open on the last row
            gIz = Y16 ? gIrow << 16 : gIrow;                        // the z candidate's weight sits on byte 1 (b10)
            if (lane == 31) LBst16 = (w < G - 1) ? (int)0x80000000 : LB16;
        };
        auto prefetch_row = [&]() {
            cp_async16(rowSlot, pf); cp_async16(rowSlot + 16, pf + 16);
            cp_async16(rowSlot + 32, pf + 32); cp_async16(rowSlot + 48, pf + 48);
            cp_async_commit();
            pf += B * sizeof(RowRec);
        };
        if (lane == 31 && w < G - 1) LBst16 = (int)0x80000000;      // (a FIFO lane stores whether or not it has a row)
        if (r <= M) {
            const uint4 *rp = reinterpret_cast<const uint4 *>(rows + r);
            const uint4 q0 = __ldg(rp), q1 = __ldg(rp + 1), q2 = __ldg(rp + 2), q3 = __ldg(rp + 3);
            unpack_row(q0, q1, q2, q3, 0);
            if (r + B <= M) prefetch_row();
        }
        unsigned acc = 0;
        unsigned *tbp = reinterpret_cast<unsigned *>(tb) + l * 2;      // this lane's two words of an 8-step group (see tb_byte)
        int Cl = MININT, Dl = MININT, Il = MININT, gCl = 0, gIl = 0;     // grid point (r, c-1)
        int Cd = MININT, Dd = MININT, Id = MININT, gCd = 0, gId = 0;     // grid point (r-1, c-1)
        unsigned fr = (unsigned)(-8 * 16) & FMASK, fw = 0;               // FIFO slots (byte offsets) of step t-8 (read) and t (write)
        __syncthreads();

        for (int t8 = 0; t8 < nSteps; t8 += 8) {
            const int g = t8 >> 3;
            // ---- wait: the warp we read from has finished the groups our reads come from; the warp that reads us is near ----
            if (lane == 0) {
                const int needPrev = (w == 0) ? g + needAdj : g;
                while (ld_volatile_shared(donePrev) < needPrev) __nanosleep(YB_F3_SLEEP);
                while (ld_volatile_shared(doneNext) < g - F3_BP) __nanosleep(YB_F3_SLEEP);
                __threadfence_block();
            }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                // ---- grid point (r-1, c) ------------------------------------------------------------------------------
                int Cu = __shfl_up_sync(FULL, Cl, 1), Du = __shfl_up_sync(FULL, Dl, 1), Iu = __shfl_up_sync(FULL, Il, 1);
                if (lane == 0) {
                    const uint4 up = lds128(w == 0 ? ringAddr + ((unsigned)c16 & RMASK) : fifoR + fr);
                    Cu = (int)up.x; Du = (int)up.y; Iu = (int)up.z;
                }
                if (c16 > RB16) {
                    // ---- this lane finished its row ------------------------------------------------------------------
                    if (l == B - 1) {
                        // the row below reads us up to its own right bound and must find never-written dp[] entries there
                        const int RBn = r < M ? __ldg(rbArr + r + 1) : (RB16 >> 4);
#pragma unroll 1
                        for (int cc = (RB16 >> 4) + 1; cc <= RBn; ++cc)
                            sts128(ringAddr + (((unsigned)cc << 4) & RMASK), MININT, MININT, MININT, 0u);
                    }
                    r += B;
                    if (r <= M) {
                        const int offOld = off16;
                        cp_async_wait_all();
                        const uint4 q0 = lds128(rowSlot), q1 = lds128(rowSlot + 16), q2 = lds128(rowSlot + 32), q3 = lds128(rowSlot + 48);
                        unpack_row(q0, q1, q2, q3, (t8 + u) * 16);
                        if (r + B <= M) prefetch_row();
                        if (l == 0) {
                            // The ring entries of the block of rows above were written `gap` steps before we read them
                            // (the schedule: >= 8).  Entries read in group g come from the last warp's groups up to
                            // g + floor((7 - gap) / 8): that many it must have completed.
                            const int gap = ((off16 - offOld) >> 4) - (B - 1) - F3_SKEW * (G - 1);
                            needAdj = 1 + ((7 - gap) >> 3);
                            while (ld_volatile_shared(donePrev) < g + needAdj) __nanosleep(YB_F3_SLEEP);
                            __threadfence_block();
                        }
                    } else {
                        if (r - B == M) { outs[p].C = Cl; outs[p].D = Dl; outs[p].I = Il; }
                        LB16 = 0x7fffffff; RB16 = 0x7fffffff; LBp16 = 0x7fffffff; LBpp16 = 0x7fffffff;
                        if (l == B - 1) LBst16 = 0x7fffffff;
                    }
                }
                const bool active = (c16 >= LB16);
                const bool hasI = c16 > LB16, hasC = c16 > LBp16;
                // existence multipliers of grid point (r-1, c): its C node exists iff c > LB[r-2], its I node iff c > LB[r-1]
                const int gCu = c16 > LBpp16 ? nGO : 0, gIu = hasC ? nGO : 0;
                const uint4 cw = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const unsigned char *>(cols) +
                                                                       (unsigned)__vimin_s32_relu(c16, N16)));
                int vI, vC, vD;
                acc >>= 8;                                   // make room for this cell's byte (bits 24..31)
                // ---- I node (mz_yama.c:114-166) -----------------------------------------------------------
                {
                    int x = Cl + dp4a_uu(cw.x, avXI, 0) * gCl;
                    int y, z;
                    if (Y16) {
                        y = dp2a_hi_su((unsigned)gIrow, cw.x, Dl);                    // K*ndB opens (mz_yama.c:131-134)
                        z = dp2a_lo_su((unsigned)gIl, cw.x, Il);                      // K*b10, if I(r,c-1) exists
                    } else {
                        y = Dl + (int)__byte_perm(cw.x, 0, 0x4442) * gIrow;
                        z = Il + (int)__byte_perm(cw.x, 0, 0x4441) * gIl;
                    }
                    vI = pick3<4>(x, y, z, hasI, acc);
                    vI = dp2a_hi_su(nKGE_lo, cw.x, vI);            // - ndB*K*gap_ext (mz_yama.c:158-161)
                }
                vI = hasI ? vI : MININT;
                // ---- C node (mz_yama.c:169-205) -----------------------------------------------------------
                {
                    int x = Cd + dp4a_uu(cw.w, avXC, 0) * gCd;
                    int y = Y16 ? dp2a_hi_su(avYC, cw.w, Dd) : Dd + dp4a_uu(cw.w, avYC, 0) * nGO;
                    int z = Id + dp4a_uu(cw.w, avZC, 0) * gId;
                    vC = pick3<0>(x, y, z, hasC, acc);
                    vC = dp2a_lo_su(w01, cw.y, vC);
                    vC = dp2a_hi_su(w23, cw.y, vC);
                    vC = dp2a_lo_su(w45, cw.z, vC);
                }
                vC = hasC ? vC : MININT;
                // ---- D node (mz_yama.c:208-242) -----------------------------------------------------------
                {
                    int x = Cu + dp4a_uu(cw.z, avXD, 0) * gCu;
                    int y = Y16 ? dp2a_hi_su(avYD, cw.z, Du) : Du + dp4a_uu(cw.z, avYD, 0) * nGO;
                    int z = Iu + dp4a_uu(cw.z, avZD, 0) * gIu;
                    vD = pick3<2>(x, y, z, true, acc) - eD;
                }
                vD = active ? vD : MININT;
                if (c16 >= LBst16) sts128(w < G - 1 ? fifoW + fw : ringAddr + ((unsigned)c16 & RMASK), vC, vD, vI, 0u);
                if (u == 3) tbp[0] = acc;
                if (u == 7) tbp[1] = acc;
                Cl = vC; Dl = vD; Il = vI;
                gCl = hasC ? nGO : 0; gIl = hasI ? gIz : 0; gCd = gCu; gId = gIu;
                Cd = Cu; Dd = Du; Id = Iu;
                c16 += 16;
                fr = (fr + 16u) & FMASK; fw = (fw + 16u) & FMASK;
            }
            tbp += 2 * B;
            // ---- publish: this warp has finished group g ------------------------------------------------------------------
            __syncwarp();
            if (lane == 31) { __threadfence_block(); st_volatile_shared(doneAddr + 4u * (unsigned)w, g + 1); }
        }
        __syncthreads();
    }
}

// =================================================================================================
// K3: traceback (mz_yama.c:257-291), one thread per pair; threads of a warp get pairs of similar size
// (the launch order is sorted by cell count).  The stored bytes are the reference's own, so a move is ONE
// dependent byte load (>= 4 moves per 32-B sector, see tb_byte) plus branch-free integer work; the
// band is not consulted.  Ops leave as 2-bit codes, 16 per 32-bit store, in the reference's (reversed)
// order: op i sits in bits 2*(i&15) of word i>>4.
// =================================================================================================
// What a 2-bit traceback field means, as a 4 x 2-bit table of node codes (FLAG_C 0, FLAG_I 1, FLAG_D 2; 3 = invalid):
//   fill_body:         e = notC | gt<<1        0,2 -> C   1 -> I   3 -> D
//   fill_body2 KEYED:  low bits of the maximum  2 -> C     1 -> I   0 -> D   (3 never occurs)
constexpr unsigned TB_DECODE_FLAGS = 0x84u, TB_DECODE_KEYED = 0xC6u;
constexpr int TB_LONG = 2048;       // paths of at least this many moves get a warp of their own (the warp-per-path
                                    // kernel is bound by instruction issue, so only where the pointer chase is critical)

__global__ void __launch_bounds__(128)
yb_traceback_kernel(const PairMeta *__restrict__ metas, const int *__restrict__ order, const int *__restrict__ nPairsPtr,
                    const unsigned char *__restrict__ blob, const unsigned char *__restrict__ tbPool,
                    const unsigned long long *__restrict__ tbBase, unsigned *__restrict__ scriptPool,
                    PairOut *__restrict__ outs, int tbLong) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= __ldg(nPairsPtr)) return;
    const int p = order[idx];
    const PairMeta pm = metas[p];
    if (pm.M + pm.N >= tbLong) return;                  // walked by yb_traceback_long_kernel, one warp per pair
    const unsigned decode = pm.cls == 2 ? TB_DECODE_KEYED : TB_DECODE_FLAGS;
    const int *sched = reinterpret_cast<const int *>(blob + pm.offSched);
    const unsigned char *tb = tbPool + __ldg(tbBase + p);
    unsigned *script = scriptPool + pm.scriptBase;
    PairOut o = outs[p];
    int node;
    if (o.C >= o.D && o.C >= o.I) node = FLAG_C;         // mz_yama.c:262-267
    else if (o.D >= o.I) node = FLAG_D;
    else node = FLAG_I;
    int r = pm.M, c = pm.N, n = 0, status = 0;
    const int limit = pm.M + pm.N;
    const unsigned tmax = (unsigned)pm.nSteps - 1u;
    const int lg = pm.lgLanes;
    const unsigned laneMask = (1u << lg) - 1u;
    int blk = -1, offBlk = 0;
    long long lastSec = -1;
    unsigned accw = 0;
    constexpr int TB_AHEAD = 4;
    while (r > 0 || c > 0) {
        if (r < 0 || c < 0 || n >= limit || node == 3) { status = -5; break; }   // mz_yama.c:274-276, :289-290
        unsigned st;
        if (r == 0) {
            st = 1u << 4;                                                   // row 0: from I (mz_yama.c:91), e = 1
        } else {
            if (((r - 1) >> lg) != blk) { blk = (r - 1) >> lg; offBlk = __ldg(sched + blk); }
            const unsigned lane = (unsigned)(r - 1) & laneMask;
            const unsigned t = min((unsigned)(c + offBlk) + lane + (unsigned)pm.skew * (lane >> 5), tmax);   // clamp: stay inside this pair
            const unsigned long long at = tb_byte(lane, t, lg);
            // The bytes were written a whole fill kernel ago: every new 32-B sector is a dependent miss to HBM.  A path
            // is mostly diagonal, so when it enters a sector the one it will need TB_AHEAD sectors later is known:
            // pull it into L2 now and the chain runs at L2 latency.
            const long long sec = (long long)(at >> 5);
            if (sec != lastSec) {
                lastSec = sec;
                const long long ahead = (long long)at - (long long)TB_AHEAD * ((8ll << lg) + 32);   // (l-4, t-8) per sector
                if (ahead >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(tb + ahead));
            }
            st = __ldg(tb + at);
        }
        accw |= (unsigned)node << (2 * (n & 15));
        if ((n & 15) == 15) { script[n >> 4] = accw; accw = 0; }
        ++n;
        // I: left, field 4 | D: up, field 2 | C: diagonal, field 0
        const int shift = node == FLAG_I ? 4 : (node == FLAG_D ? 2 : 0);
        r -= (node != FLAG_I);
        c -= (node != FLAG_D);
        node = (int)((decode >> (2u * ((st >> shift) & 3u))) & 3u);     // TB_DECODE_FLAGS / TB_DECODE_KEYED
    }
    if (n & 15) script[n >> 4] = accw;
    o.m_new = n;
    o.status = status;
    outs[p] = o;
}

// K3 for long paths: the time of the thread-per-pair kernel is (longest path) x (latency of a miss), because a warp
// iteration waits for its slowest lane and some lane misses on every iteration.  A long path therefore gets a warp
// of its own: all 32 lanes walk the same path (uniform control flow, the byte load is a broadcast), and whenever the
// path enters a new 32-B sector, lanes 0..TB_FAN-1 prefetch the next sectors of its diagonal continuation, so that
// the walk finds them in L2 (measured: 20 000-move paths 5.7 -> 2.2 ms).
constexpr int TB_FAN = 8;
__global__ void __launch_bounds__(128)
yb_traceback_long_kernel(const PairMeta *__restrict__ metas, const int *__restrict__ longList, const int *__restrict__ nLongPtr,
                         const unsigned char *__restrict__ blob, const unsigned char *__restrict__ tbPool,
                         const unsigned long long *__restrict__ tbBase, unsigned *__restrict__ scriptPool,
                         PairOut *__restrict__ outs) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= __ldg(nLongPtr)) return;
    const int p = longList[w];
    const PairMeta pm = metas[p];
    const unsigned decode = pm.cls == 2 ? TB_DECODE_KEYED : TB_DECODE_FLAGS;
    const int *sched = reinterpret_cast<const int *>(blob + pm.offSched);
    const unsigned char *tb = tbPool + __ldg(tbBase + p);
    unsigned *script = scriptPool + pm.scriptBase;
    PairOut o = outs[p];
    int node;
    if (o.C >= o.D && o.C >= o.I) node = FLAG_C;         // mz_yama.c:262-267
    else if (o.D >= o.I) node = FLAG_D;
    else node = FLAG_I;
    int r = pm.M, c = pm.N, n = 0, status = 0;
    const int limit = pm.M + pm.N;
    const unsigned tmax = (unsigned)pm.nSteps - 1u;
    const int lg = pm.lgLanes;
    const unsigned laneMask = (1u << lg) - 1u;
    const long long secStride = (8ll << lg) + 32;         // (l-4, t-8): the next sector of a diagonal path
    int blk = -1, offBlk = 0;
    long long lastSec = -1;
    unsigned accw = 0;
    while (r > 0 || c > 0) {
        if (r < 0 || c < 0 || n >= limit || node == 3) { status = -5; break; }   // mz_yama.c:274-276, :289-290
        unsigned st;
        if (r == 0) {
            st = 1u << 4;                                                   // row 0: from I (mz_yama.c:91), e = 1
        } else {
            if (((r - 1) >> lg) != blk) { blk = (r - 1) >> lg; offBlk = __ldg(sched + blk); }
            const unsigned ln = (unsigned)(r - 1) & laneMask;
            const unsigned t = min((unsigned)(c + offBlk) + ln + (unsigned)pm.skew * (ln >> 5), tmax);       // clamp: stay inside this pair
            const unsigned long long at = tb_byte(ln, t, lg);
            const long long sec = (long long)(at >> 5);
            if (sec != lastSec) {
                lastSec = sec;
                const long long ahead = (long long)at - (long long)(lane + 1) * secStride;
                if (lane < TB_FAN && ahead >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(tb + ahead));
            }
            st = __ldg(tb + at);
        }
        accw |= (unsigned)node << (2 * (n & 15));
        if ((n & 15) == 15) { if (lane == 0) script[n >> 4] = accw; accw = 0; }
        ++n;
        const int shift = node == FLAG_I ? 4 : (node == FLAG_D ? 2 : 0);
        r -= (node != FLAG_I);
        c -= (node != FLAG_D);
        node = (int)((decode >> (2u * ((st >> shift) & 3u))) & 3u);     // TB_DECODE_FLAGS / TB_DECODE_KEYED
    }
    if (lane == 0) {
        if (n & 15) script[n >> 4] = accw;
        o.m_new = n;
        o.status = status;
        outs[p] = o;
    }
}

}  // namespace yb
