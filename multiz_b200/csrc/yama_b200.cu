// yama_b200.cu -- host runtime + C ABI (include/yama_b200.h) around the sm_100a kernels.
//
// One context owns 1..8 devices.  A batch of independent block pairs is cut into WAVES (contiguous job
// ranges, ~64 MB of input each).  Waves are handed out dynamically to the devices -- no collective: the
// merge has no cross-pair dependency -- and every device pipelines its waves through a
// ring of staging slots, each with its own stream, pinned buffers and device buffers:
//
//     host threads: analyse (band checks of mz_yama.c:58-71, wavefront schedule) + pack into pinned memory
//     stream:       H2D -> K1 profile -> K2 fill (one launch per ring-size bin) -> K3 traceback -> D2H
//     host threads: expand the 2-bit edit scripts into the caller's result array
//
// so that packing wave w+1 and unpacking wave w-1 overlap the copies and kernels of wave w.
// There is no CPU implementation of the DP in this library.
#include "../../include/yama_b200.h"
#include "yama_kernels.cuh"
#include "score_kernels.cuh"

#include <emmintrin.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace yb;

extern "C" int64_t yb_band_scan(int M, int N, const int32_t *LB, const int32_t *RB, int32_t *wmax, int32_t *sched,
                                int32_t *nSteps, int (*lanesOf)(int, int), int32_t *lanes, int32_t *connected);   // band_scan.cpp

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t need) {
        if (need <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = need + need / 4 + (1u << 20);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { (void)cudaGetLastError(); e = cudaMalloc(&p, need); want = need; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t need) {
        if (need <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = need + need / 4 + (1u << 20);
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// Persistent helper threads of one device: run(n, chunk, fn) calls fn(lo, hi) over [0,n) in dynamic chunks on
// the helpers plus the calling thread, and returns when all of [0,n) is done.
class Pool {
  public:
    explicit Pool(int helpers) {
        for (int t = 0; t < helpers; ++t) th_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> g(mu_); stop_ = true; ++gen_; genFast_.store(gen_, std::memory_order_release); }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    template <class F>
    void run(int64_t n, int64_t chunk, F &&fn) {
        if (n <= 0) return;
        if (chunk < 1) chunk = 1;
        const int64_t nchunks = (n + chunk - 1) / chunk;
        if (th_.empty() || nchunks == 1) { fn((int64_t)0, n); return; }
        std::function<void(int64_t, int64_t)> f = std::ref(fn);
        {
            std::lock_guard<std::mutex> g(mu_);
            fn_ = &f; n_ = n; chunk_ = chunk; nchunks_ = nchunks;
            next_.store(0); pending_ = (int)th_.size();
            pendingFast_.store(pending_, std::memory_order_release);
            ++gen_;
            genFast_.store(gen_, std::memory_order_release);
        }
        cv_.notify_all();
        work();
        for (int spin = 0; spin < 20000 && pendingFast_.load(std::memory_order_acquire) != 0; ++spin) _mm_pause();
        std::unique_lock<std::mutex> g(mu_);
        done_.wait(g, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

  private:
    void work() {
        for (;;) {
            int64_t c = next_.fetch_add(1);
            if (c >= nchunks_) break;
            (*fn_)(c * chunk_, std::min(n_, (c + 1) * chunk_));
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            // runs of one batch follow each other within microseconds: poll briefly before sleeping on the condvar
            for (int spin = 0; spin < 4000 && genFast_.load(std::memory_order_acquire) == seen; ++spin) _mm_pause();
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            work();
            std::lock_guard<std::mutex> g(mu_);
            pendingFast_.fetch_sub(1, std::memory_order_release);
            if (--pending_ == 0) done_.notify_one();
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::function<void(int64_t, int64_t)> *fn_ = nullptr;
    int64_t n_ = 0, chunk_ = 1, nchunks_ = 0;
    std::atomic<int64_t> next_{0};
    int pending_ = 0;
    uint64_t gen_ = 0;
    std::atomic<uint64_t> genFast_{0};
    std::atomic<int> pendingFast_{0};
    bool stop_ = false;
};

constexpr int NBINS = 9;
constexpr int NB = 160;                    // launch-order buckets per ring bin
// Kernel bins: ring entries (>= widest band row + 32), warps per pair G, pairs per CTA P.  Narrow bands and short
// pairs run one warp per pair; wide bands on long pairs run one CTA per pair (one ring per pair -> full occupancy).
// Bins 5..8 are the bulk kernels (fill_body2) for pairs of kernel class 1 (bins 5, 6) and 2 (KEYED; bins 7, 8): same
// rings as bins 0 and 1.
struct BinCfg { int ring, G, P, minRows; };
const BinCfg kBin[NBINS] = {{128, 1, 8, 0}, {512, 1, 8, 0}, {512, 4, 1, 192}, {2048, 8, 1, 0}, {4096, 8, 1, 0},
                            {128, 1, F2_WARPS, 0}, {512, 1, F2_WARPS, 0}, {128, 1, F2_WARPS, 0}, {512, 1, F2_WARPS, 0}};
constexpr int BULK_BIN0 = 5;
constexpr int TB_GROUP = 3;               // waves per traceback launch group
constexpr int NSLOTS = 2 * TB_GROUP;      // one group filling while the previous one drains

struct JobInfo {           // host-side facts about one pair
    int64_t cells = 0;     // tback_size of the reference
    int nSteps = 0;        // wavefront steps (schedule below); traceback bytes = 32 * nSteps
    int wmax = 0;          // widest band row
    int status = YB_OK;
    int bin = 0;           // kernel bin (ring size, warps per pair)
    int bucket = 0;        // launch-order bucket: ring bin * NB + quarter-octave of the cell count, descending
    int connected = 0;     // every band row is reachable from the row above (LB[r] <= RB[r-1] + 1)
    int cls = 0;           // kernel class (PairMeta::cls): 0 fill_body, 1 fill_body2, 2 fill_body2 KEYED
};

// One staging slot = one wave in flight.
struct Slot {
    cudaStream_t stream = nullptr;
    cudaStream_t binStream[NBINS] = {};    // fill kernels of the wider ring bins run beside the main one
    cudaEvent_t binDone[NBINS] = {};
    cudaEvent_t ev[8] = {};
    DevBuf dIn, dRow, dCol, dTb, dScript, dOut, dQueue;
    PinBuf hIn, hOut;
    uint8_t *scriptDst = nullptr;          // where this wave's packed scripts go in the context's pinned store
    bool filled = false;                   // H2D + K1 + K2 queued, K3 not yet
    bool busy = false;                     // K3 + D2H queued, results not yet unpacked
    size_t sentLo = 0, sentHi = 0;         // bytes [sentLo, sentHi) of the blob were queued for H2D while the rest was packed
    double tPack0 = 0, tPack1 = 0, tTbLaunch = 0;   // host times (ms) of the wave: pack start/end, traceback launch
    // the wave it holds
    int64_t first = 0, count = 0;
    std::vector<JobInfo> info;
    std::vector<uint32_t> scriptOff;       // per pair, word offset in the wave's script pool
    struct Off { size_t blob, row, col; uint32_t script; };
    std::vector<Off> off;                  // per pair, dimension-only offsets into the pools
    std::vector<int> bucketCount;
    size_t blobBytes = 0, metaBytes = 0, orderOff = 0, longOff = 0, tbBaseOff = 0, scriptWords = 0;
    int nLong = 0;                         // pairs whose traceback path gets a warp of its own
    int tbLong = TB_LONG;                  // ... those with at least this many moves
    bool y16 = false;                      // every pair has K*gap_open <= 32767: the kernels' 16-bit weight forms
    bool ungated = false;                  // every pair is small enough for the fill variant without existence multipliers
    int nValid = 0;
    int binStart[NBINS + 1] = {};
};

// Block scoring (yb_score_blocks): one wave of blocks in flight per stage, two stages so that packing the next
// wave overlaps the copy and the kernel of the previous one.
struct ScoreStage {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {};
    DevBuf dIn, dSums;
    PinBuf hIn, hSums;
    int64_t first = 0, count = 0;
    bool busy = false;
};

struct Device {
    int id = -1;
    int sms = 0;
    Slot slots[NSLOTS];
    ScoreStage score[2];
    cudaEvent_t evBase = nullptr;          // YB_PROFILE=2: start of the batch on this device (timeline origin)
    double tBase = 0;
    int timeline = 0;
    ScoreConst sc{};                       // the owning context's score tables (kernel arguments)
    int fillBlocks[NBINS] = {};
    int helpers = 1;
    std::unique_ptr<Pool> pool;
    // accumulated stats of the current call
    double kernel_ms = 0, fill_ms = 0, profile_ms = 0, tb_ms = 0, h2d_ms = 0, d2h_ms = 0, pack_ms = 0, unpack_ms = 0;
    int64_t h2d_bytes = 0, d2h_bytes = 0, cells = 0, failed = 0;
    int launches = 0, waves = 0;
    std::string err;
    int64_t errJob = 0;
    bool hasResident = false;
    double t_layout = 0, t_par = 0, t_post = 0, t_reserve = 0, t_wait = 0, t_launch = 0;   // YB_PROFILE breakdown
};

}  // namespace

struct yb_ctx {
    std::vector<Device> devs;
    std::string err;
    bool scoresSet = false;
    ScoreConst sc{};
    int maxDepth = 255;
    int maxAbsS = 1;                        // max |S6|
    bool ungatedOk = true;                  // YB_UNGATED=0 keeps the existence multipliers in every fill kernel
    int maxCls = 2;                         // highest kernel class handed out: YB_FILL2=0 -> 0 (fill_body only), YB_KEYED=0 -> 1
    int nThreads = 1;
    size_t waveInBytes = (size_t)64 << 20;  // input bytes per wave (steady state)
    size_t waveMinBytes = (size_t)16 << 20; // first waves of a batch (the device idles while the first wave is packed)
    size_t waveTailBytes = (size_t)16 << 20; // last waves of a batch (a short last wave shortens the traceback + unpack tail)
    size_t batchBlobBytes = 0;              // input bytes of the current batch (dimension-only estimate)
    size_t waveTbBytes = (size_t)12 << 30;  // traceback bytes per wave (device memory per slot)
    int64_t wavePairs = 1 << 20;
    int tbLong = TB_LONG;                   // paths of at least this many moves: warp-per-path traceback (YB_TB_LONG)
    bool ntStores = true;                   // full-line non-temporal staging stores (YB_NT=0 turns them off)
    bool earlyH2D = true;                   // a wave's parts are copied while the rest is packed (YB_EARLY_H2D=0 turns it off)
    // results of the last batch
    uint8_t *scriptStore = nullptr;         // pinned (portable): the D2H copies of the waves land here directly
    size_t scriptStoreCap = 0;
    std::vector<uint64_t> scriptOff;
    std::vector<uint64_t> blobPrefix;       // input bytes of jobs [0, i): waves are cut by binary search
    // record/replay queue
    std::vector<uint8_t> arena;
    struct QJob { int K, M, L, N; size_t offA, offB, offLB, offRB; };
    std::vector<QJob> queued;
    std::vector<yb_result> queuedRes;
    // resident mode
    std::vector<yb_job> resJobs;
    std::vector<int64_t> resSplit;          // device d owns jobs [resSplit[d], resSplit[d+1])
    int64_t resCells = 0;
};

namespace {

void set_err(yb_ctx *ctx, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    ctx->err = buf;
}

#define CUDA_TRY(dev, call)                                                                       \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            char b_[256];                                                                         \
            snprintf(b_, sizeof b_, "CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__,   \
                     __LINE__);                                                                   \
            (dev).err = b_;                                                                       \
            return YB_ERR_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Non-temporal copy of n bytes to a 64-byte aligned destination whose section is padded to whole 64-byte lines:
// only full lines are streamed (the tail goes through a zero-padded line buffer), so write-combining buffers never
// flush partially and the staging buffer is never read for ownership.
inline void copy_nt64(unsigned char *dst, const unsigned char *src, size_t n) {
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
        __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 16));
        __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 32));
        __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 48), d);
    }
    if (i < n) {
        alignas(64) unsigned char line[64] = {0};
        memcpy(line, src + i, n - i);
        for (int k = 0; k < 64; k += 16)
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + k), _mm_load_si128(reinterpret_cast<const __m128i *>(line + k)));
    }
}
inline void pack_band_nt64(uint32_t *dst, const int32_t *lb, const int32_t *rb, int rowsTotal) {
    int r = 0;
    for (; r + 16 <= rowsTotal; r += 16)
        for (int k = 0; k < 16; k += 4) {
            __m128i l = _mm_loadu_si128(reinterpret_cast<const __m128i *>(lb + r + k));
            __m128i h = _mm_loadu_si128(reinterpret_cast<const __m128i *>(rb + r + k));
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + r + k), _mm_or_si128(l, _mm_slli_epi32(h, 16)));
        }
    if (r < rowsTotal) {
        alignas(64) uint32_t line[16] = {0};
        for (int k = 0; r + k < rowsTotal; ++k) line[k] = (uint32_t)lb[r + k] | ((uint32_t)rb[r + k] << 16);
        for (int k = 0; k < 16; k += 4)
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + r + k), _mm_load_si128(reinterpret_cast<const __m128i *>(line + k)));
    }
}
// band rows as LB | RB<<16 (both < 65536).  (Non-temporal stores were tried for the staging copies and lost:
// the sections are small and rarely cache-line aligned, so write-combining buffers flush partially filled.)
inline void pack_band(uint32_t *dst, const int32_t *lb, const int32_t *rb, int rowsTotal) {
    int r = 0;
    for (; r + 4 <= rowsTotal; r += 4) {
        __m128i l = _mm_loadu_si128(reinterpret_cast<const __m128i *>(lb + r));
        __m128i h = _mm_loadu_si128(reinterpret_cast<const __m128i *>(rb + r));
        _mm_storeu_si128(reinterpret_cast<__m128i *>(dst + r), _mm_or_si128(l, _mm_slli_epi32(h, 16)));
    }
    for (; r < rowsTotal; ++r) dst[r] = (uint32_t)lb[r] | ((uint32_t)rb[r] << 16);
}

// one-off variant for work outside a device's pipeline
template <class F>
void parallel_for(int threads, int64_t n, int64_t chunk, F &&fn) {
    Pool pool(std::max(0, threads - 1));
    pool.run(n, chunk, fn);
}

int bin_of(int wmax, int M) {
    if (wmax + 32 <= kBin[0].ring) return 0;
    if (wmax + 32 <= kBin[1].ring) return M >= kBin[2].minRows ? 2 : 1;
    if (wmax + 32 <= kBin[3].ring) return 3;
    if (wmax + 32 <= kBin[4].ring) return 4;
    return -1;
}
// the bulk kernels take over the warp-per-pair bins for pairs of class 1 / 2
inline int bin_of_cls(int bin, int cls) { return (cls >= 1 && bin >= 0 && bin <= 1) ? BULK_BIN0 + 2 * (cls - 1) + bin : bin; }
int lanes_of(int wmax, int M) {          // wavefront width of the pair's bin (0: no kernel takes it)
    const int b = bin_of(wmax, M);
    return b < 0 ? 0 : 32 * kBin[b].G;
}
inline int lg_of(int lanes) { return 31 - __builtin_clz((unsigned)lanes); }

size_t fill_smem(int bin) {
    // rings (RING*16-aligned, hence the slack) + 32 B of mailbox per lane + the queue slot
    const BinCfg &c = kBin[bin];
    return (size_t)c.P * ((size_t)c.ring * 16 + (size_t)c.G * 1024) + (size_t)c.ring * 16 + 16;
}

}  // namespace

// ---- kernels with a runtime warps-per-CTA: thin wrappers around the template ---------------------
namespace yb {
template <int RING, int G, int P, bool Y16, bool GATED = true>
__global__ void __launch_bounds__(G * P * 32)
yb_fill_kernel_w(const PairMeta *__restrict__ metas, const int *__restrict__ order, int nPairs,
                 int *__restrict__ queue, const RowRec *__restrict__ rowPool,
                 const ColRec *__restrict__ colPool, unsigned char *__restrict__ tbPool,
                 const unsigned long long *__restrict__ tbBase, PairOut *__restrict__ outs, int gapOpen, int gapExt) {
    fill_body<RING, G, P, Y16, GATED>(metas, order, nPairs, queue, rowPool, colPool, tbPool, tbBase, outs, gapOpen, gapExt);
}
// bulk form (fill_body2): one warp per pair, pairs of kernel class 1 / 2
#ifndef YB_F2_MINCTAS
#define YB_F2_MINCTAS 4
#endif
template <int RING, bool KEYED>
__global__ void __launch_bounds__(F2_WARPS * 32, (RING <= 128 ? YB_F2_MINCTAS : 1))
yb_fill2_kernel(const PairMeta *__restrict__ metas, const int *__restrict__ order, int nPairs,
                int *__restrict__ queue, const RowRec *__restrict__ rowPool,
                const ColRec *__restrict__ colPool, unsigned char *__restrict__ tbPool,
                const unsigned long long *__restrict__ tbBase, PairOut *__restrict__ outs, int nGO, int gapExt) {
    fill_body2<RING, KEYED>(metas, order, nPairs, queue, rowPool, colPool, tbPool, tbBase, outs, nGO, gapExt);
}
}  // namespace yb

namespace {

typedef void (*FillFn)(const PairMeta *, const int *, int, int *, const RowRec *, const ColRec *,
                       unsigned char *, const unsigned long long *, PairOut *, int, int);
// ungated: the variant without existence multipliers (bulk bins only), for waves of small enough pairs (Slot::ungated)
FillFn fill_fn(int bin, bool y16, bool ungated = false) {
    if (ungated && y16 && bin == 0) return yb_fill_kernel_w<128, 1, 8, true, false>;
    if (ungated && y16 && bin == 1) return yb_fill_kernel_w<512, 1, 8, true, false>;
    switch (bin) {
        case 0: return y16 ? yb_fill_kernel_w<128, 1, 8, true> : yb_fill_kernel_w<128, 1, 8, false>;
        case 1: return y16 ? yb_fill_kernel_w<512, 1, 8, true> : yb_fill_kernel_w<512, 1, 8, false>;
        case 2: return y16 ? yb_fill_kernel_w<512, 4, 1, true> : yb_fill_kernel_w<512, 4, 1, false>;
        case 3: return y16 ? yb_fill_kernel_w<2048, 8, 1, true> : yb_fill_kernel_w<2048, 8, 1, false>;
        default: return y16 ? yb_fill_kernel_w<4096, 8, 1, true> : yb_fill_kernel_w<4096, 8, 1, false>;
    }
}

FillFn fill_fn2(int bin) {
    switch (bin - BULK_BIN0) {
        case 0: return yb_fill2_kernel<128, false>;
        case 1: return yb_fill2_kernel<512, false>;
        case 2: return yb_fill2_kernel<128, true>;
        default: return yb_fill2_kernel<512, true>;
    }
}
size_t fill2_smem(int bin) {
    // rings (RING*16-aligned, hence the slack) + a 64-B row slot per lane + a stash word per warp
    const size_t ring = (size_t)kBin[bin].ring * 16;
    return F2_WARPS * (ring + 32 * 64 + 16) + ring + 16;
}

int device_init(Device &d) {
    CUDA_TRY(d, cudaSetDevice(d.id));
    cudaDeviceProp prop;
    CUDA_TRY(d, cudaGetDeviceProperties(&prop, d.id));
    d.sms = prop.multiProcessorCount;
    for (auto &s : d.slots) {
        CUDA_TRY(d, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        for (auto &e : s.ev) CUDA_TRY(d, cudaEventCreate(&e));
        for (int b = 1; b < NBINS; ++b) {
            CUDA_TRY(d, cudaStreamCreateWithFlags(&s.binStream[b], cudaStreamNonBlocking));
            CUDA_TRY(d, cudaEventCreateWithFlags(&s.binDone[b], cudaEventDisableTiming));
        }
    }
    for (int b = 0; b < BULK_BIN0; ++b) {
        FillFn fn = fill_fn(b, true);
        size_t sm = fill_smem(b);
        CUDA_TRY(d, cudaFuncSetAttribute(fill_fn(b, false), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        if (fill_fn(b, true, true) != fn)
            CUDA_TRY(d, cudaFuncSetAttribute(fill_fn(b, true, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        CUDA_TRY(d, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        int occ = 0;
        CUDA_TRY(d, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kBin[b].G * kBin[b].P * 32, sm));
        if (occ < 1) occ = 1;
        d.fillBlocks[b] = occ * d.sms;
    }
    for (int b = BULK_BIN0; b < NBINS; ++b) {
        FillFn fn = fill_fn2(b);
        const size_t sm = fill2_smem(b);
        CUDA_TRY(d, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        int occ = 0;
        CUDA_TRY(d, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, F2_WARPS * 32, sm));
        d.fillBlocks[b] = std::max(occ, 1) * d.sms;
    }
    return YB_OK;
}

// Validation loop of mz_yama.c:58-71 in the reference's wording; returns the cell count (tback_size).
int64_t check_band(int M, int N, const int *LB, const int *RB, char *msg, int msglen, int *wmax) {
    if (LB[0] != 0 || RB[M] != N) {
        if (msg) snprintf(msg, msglen, "LB and RB not terminated properly: %d %d %d", LB[0], RB[M], N);
        return YB_ERR_BAND;
    }
    const int need = N < 10 ? N : 10;
    int64_t cells = 0;
    int wm = 0;
    for (int r = 0; r <= M; ++r) {
        int j = RB[r] - LB[r];
        if (j < need) {
            if (msg) snprintf(msg, msglen, "RB[%d] - LB[%d] < %d, %d %d %d", r, r, need, RB[r], LB[r], N);
            return YB_ERR_BAND;
        }
        cells += j + 1;
        if (j + 1 > wm) wm = j + 1;
        if (r > 0 && LB[r] < LB[r - 1]) { if (msg) snprintf(msg, msglen, "LB not monotonic"); return YB_ERR_BAND; }
        if (r > 0 && RB[r] < RB[r - 1]) { if (msg) snprintf(msg, msglen, "RB not monotonic"); return YB_ERR_BAND; }
    }
    if (wmax) *wmax = wm;
    return cells;
}

// Wavefront schedule (see K2): rows Bb+1..Bb+B run on lanes 0..B-1 with column = step - (OFF_b + lane).
// OFF grows per block by at least B (lane 0 stays behind the last lane of the previous block) and by enough
// that a lane starts its next row only after the row below its current one has stopped reading it.
// Returns the step count; sched (may be null) receives ceil(M/B) block offsets.
int make_schedule(int M, const int *LB, const int *RB, int B, int *sched) {
    int off = 0;
    const int nblk = (M + B - 1) / B;
    for (int b = 0; b < nblk; ++b) {
        if (sched) sched[b] = off;
        int need = B;
        const int r0 = B * b + 1, r1 = std::min(M - B, B * b + B);
        for (int r = r0; r <= r1; ++r) need = std::max(need, RB[r + 1] - LB[r + B] + 3);
        if (b == nblk - 1) {
            int lane = (M - 1) % B;
            int last = off + lane + RB[M];                 // step of the last cell
            return ((last + 2) + 7) & ~7;                  // +1 step to publish the final scores, whole 8-step groups
        }
        off += need;
    }
    return 8;
}

// Everything the host needs to know about one job; msg (optional) gets the reference's wording.
void analyse_one(const yb_ctx *ctx, const yb_job &j, JobInfo &ji, int *sched, char *msg, int msglen) {
    ji = JobInfo();
    if (j.K < 1 || j.L < 1 || j.M < 1 || j.N < 1 || !j.A || !j.B || !j.LB || !j.RB) {
        ji.status = YB_ERR_ARG;
        if (msg) snprintf(msg, msglen, "bad dimensions K=%d M=%d L=%d N=%d", j.K, j.M, j.L, j.N);
        return;
    }
    // vectorised scan (band_scan.cpp): validation + cells + widest row + schedule in one pass; on a violation
    // the scalar loop below words the message as the reference does
    int nSteps = 0, lanes = 0;
    int64_t cells = yb_band_scan(j.M, j.N, j.LB, j.RB, &ji.wmax, sched, &nSteps, lanes_of, &lanes, &ji.connected);
    if (cells < 0) {
        ji.connected = 0;
        cells = check_band(j.M, j.N, j.LB, j.RB, msg, msglen, &ji.wmax);
        if (cells < 0) { ji.status = YB_ERR_BAND; return; }
        lanes = lanes_of(ji.wmax, j.M);                        // (unreachable unless the two scans disagree)
        if (lanes > 0) nSteps = make_schedule(j.M, j.LB, j.RB, lanes, sched);
    }
    ji.cells = cells;
    if (j.K > ctx->maxDepth || j.L > 255) {
        ji.status = YB_ERR_LIMIT;
        if (msg) snprintf(msg, msglen, "profile depth K=%d L=%d exceeds the kernel limit (%d/255 rows)", j.K, j.L, ctx->maxDepth);
        return;
    }
    if (lanes <= 0) {
        ji.status = YB_ERR_LIMIT;
        if (msg) snprintf(msg, msglen, "band row of %d cells exceeds the kernel limit (%d)", ji.wmax, kBin[NBINS - 1].ring - 32);
        return;
    }
    ji.bin = bin_of(ji.wmax, j.M);
    ji.nSteps = nSteps;
    // Kernel class.  Without existence multipliers (classes 1, 2) a candidate from a node that does not exist (exactly
    // MININT = -2^30) is charged a gap-open the reference skips.  That cannot change any real value, flag or script as long
    // as real scores stay within +-2^28 and unreal ones within [-2^31, -2^29): both follow from
    // (M+N)*K*L*(gap_open+gap_extend+max|S|) < 2^28 and a connected band.  The KEYED class carries
    // 4*value + priority, hence 2^26; both need every pre-multiplied weight of the bulk row record to fit 16 bits.
    if (ctx->maxCls >= 1 && ji.bin <= 1 && ji.connected) {
        const long double work = (long double)((int64_t)j.M + j.N) * j.K * j.L * (ctx->sc.gap_open + ctx->sc.gap_ext + ctx->maxAbsS);
        const long long w16 = std::max<long long>((long long)j.K * (ctx->sc.gap_open + ctx->sc.gap_ext), 2ll * j.K * ctx->maxAbsS);
        if (ctx->maxCls >= 2 && 4 * w16 <= 32767 && work < (long double)(1 << 26)) ji.cls = 2;
        else if (w16 <= 32767 && work < (long double)(1 << 28)) ji.cls = 1;
        ji.bin = bin_of_cls(ji.bin, ji.cls);
    }
}

inline bool dims_ok(const yb_job &j) { return j.K >= 1 && j.L >= 1 && j.M >= 1 && j.N >= 1 && j.A && j.B && j.LB && j.RB; }
inline int band_fmt(const yb_job &j) { return j.N < 65536 ? 0 : 1; }
inline size_t sched_ints(const yb_job &j) { return (size_t)((j.M + 31) >> 5); }   // upper bound (B >= 32)

// bytes a job takes in the input blob: known from its dimensions alone (wave planning needs no band read)
inline size_t blob_bytes(const yb_job &j) {
    return align_up((size_t)j.K * j.M, 64) + align_up((size_t)j.L * j.N, 64) +
           align_up((size_t)(j.M + 1) * (band_fmt(j) ? 8 : 4), 64) + align_up(sched_ints(j) * 4, 64);
}
// Analyse + pack jobs [first, first+count) into the slot's pinned buffer, one pass over the caller's data:
// everything but the traceback size of a pair follows from its dimensions, so blob / record / script offsets
// are prefix sums taken up front and each helper validates the band (mz_yama.c:58-71), writes the schedule
// and copies A, B and the band while they are hot in its cache.  Traceback offsets are assigned afterwards.
// `maxTb` caps the wave's traceback bytes: the wave is cut short (count shrinks, at least one job stays).
int slot_pack(yb_ctx *ctx, Device &d, Slot &s, const yb_job *jobs, int64_t first, int64_t &count, size_t maxTb,
              bool earlyH2D = false) {
    const double t0 = now_ms();
    s.tPack0 = t0;
    // ---- dimension-only layout (serial, a few ns per job) ------------------------------------------------------
    struct Off { size_t blob, row, col; uint32_t script; };
    s.off.resize((size_t)count);
    size_t blob = 0, rows = 0, cols = 0, words = 0;
    int maxK = 0;
    int64_t maxWork = 0;                     // max over pairs of (M+N)*K*L: bounds every real score and every drift
    for (int64_t i = 0; i < count; ++i) {
        const yb_job &j = jobs[first + i];
        s.off[(size_t)i] = Slot::Off{blob, rows, cols, (uint32_t)words};
        if (!dims_ok(j)) continue;
        maxK = std::max(maxK, j.K);
        maxWork = std::max(maxWork, ((int64_t)j.M + j.N) * j.K * j.L);
        blob += blob_bytes(j); rows += (size_t)j.M + 1; cols += (size_t)j.N + 1;
        words += ((size_t)j.M + j.N + 15) / 16;
    }
    s.first = first;
    s.y16 = (int64_t)std::min(maxK, ctx->maxDepth) * ctx->sc.gap_open <= 32767;
    // Without existence multipliers a candidate from a node that does not exist (exactly MININT = -2^30) is charged a
    // gap-open the reference skips.  That cannot change any real value, flag or script as long as real scores stay
    // within +-2^28 and unreal ones within [-2^31, -2^29): both follow from (M+N)*K*L*(gap_open+gap_extend+max|S|) < 2^28.
    s.ungated = ctx->ungatedOk && (long double)maxWork * (ctx->sc.gap_open + ctx->sc.gap_ext + ctx->maxAbsS) < (long double)(1 << 28);
    s.metaBytes = align_up((size_t)count * sizeof(PairMeta), 256);
    s.orderOff = s.metaBytes;
    s.longOff = s.orderOff + align_up((size_t)count * 4, 256);
    s.tbBaseOff = s.longOff + align_up((size_t)count * 4, 256);
    const size_t dataOff = s.tbBaseOff + align_up((size_t)count * 8, 256);   // 64-B aligned: every section is
    CUDA_TRY(d, s.hIn.reserve(dataOff + blob));
    s.info.resize((size_t)count);
    s.scriptOff.resize((size_t)count);
    unsigned char *h = static_cast<unsigned char *>(s.hIn.p);
    PairMeta *metas = reinterpret_cast<PairMeta *>(h);

    const double t1 = now_ms();
    d.t_layout += t1 - t0;
    // ---- analyse + pack (parallel) -----------------------------------------------------------------------------
    std::mutex errMu;
    // The wave is packed in a few contiguous parts; a part's bytes start their way to the device while the next part
    // is packed (the header -- metas, launch order, traceback offsets -- and the last part follow in slot_launch_fill).
    // Without this a wave's copy starts only when all of it is packed: 1.3 ms per 64 MB that the device idles at the
    // start of a batch.
    s.sentLo = s.sentHi = 0;
    const int nParts = (earlyH2D && blob >= ((size_t)4 << 20) && count >= 64) ? 4 : 1;
    if (nParts > 1) {
        CUDA_TRY(d, s.dIn.reserve(dataOff + blob));
        CUDA_TRY(d, cudaEventRecord(s.ev[0], s.stream));
    }
    int64_t partLo = 0;
    for (int part = 0; part < nParts; ++part) {
        int64_t partHi = count;
        if (part + 1 < nParts) {                                    // cut by bytes
            const size_t want = blob / (size_t)nParts * (size_t)(part + 1);
            partHi = std::lower_bound(s.off.begin() + partLo, s.off.end(), want,
                                      [](const Slot::Off &o, size_t v) { return o.blob < v; }) - s.off.begin();
            partHi = std::max(partLo, std::min<int64_t>(partHi, count));
        }
        const int64_t base = partLo;
        // dynamic chunks: 32 pairs when pairs are small, fewer when a part holds only a few (large) pairs
        const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(32, (partHi - partLo) / (4 * (int64_t)d.helpers)));
        d.pool->run(partHi - partLo, chunk, [&](int64_t lo0, int64_t hi0) {
            for (int64_t i = base + lo0; i < base + hi0; ++i) {
                const yb_job &j = jobs[first + i];
                JobInfo &ji = s.info[(size_t)i];
                const Slot::Off &of = s.off[(size_t)i];
                PairMeta pm;
                memset(&pm, 0, sizeof pm);
                s.scriptOff[(size_t)i] = of.script;
                char msg[256];
                msg[0] = 0;
                // the schedule goes straight to its place in the blob (after A, B and the band)
                size_t o = dataOff + of.blob;
                const bool dimsOk = dims_ok(j);
                size_t oA = o, oB = 0, oBand = 0, oSched = 0;
                if (dimsOk) {
                    oB = oA + align_up((size_t)j.K * j.M, 64);
                    oBand = oB + align_up((size_t)j.L * j.N, 64);
                    oSched = oBand + align_up((size_t)(j.M + 1) * (band_fmt(j) ? 8 : 4), 64);
                }
                analyse_one(ctx, j, ji, dimsOk ? reinterpret_cast<int *>(h + oSched) : nullptr, msg, sizeof msg);
                if (ji.status != YB_OK) {
                    metas[i] = pm;
                    std::lock_guard<std::mutex> g(errMu);
                    if (d.err.empty() || first + i < d.errJob) {
                        char b[400];
                        if (ji.status == YB_ERR_BAND) snprintf(b, sizeof b, "%s", msg);   // reference wording
                        else snprintf(b, sizeof b, "job %lld: %s", (long long)(first + i), msg);
                        d.err = b;
                        d.errJob = first + i;
                    }
                    continue;
                }
                pm.K = j.K; pm.M = j.M; pm.L = j.L; pm.N = j.N;
                pm.offA = oA; pm.offB = oB;
                if (ctx->ntStores) {
                    copy_nt64(h + oA, j.A, (size_t)j.K * j.M);
                    copy_nt64(h + oB, j.B, (size_t)j.L * j.N);
                } else {
                    memcpy(h + oA, j.A, (size_t)j.K * j.M);
                    memcpy(h + oB, j.B, (size_t)j.L * j.N);
                }
                pm.offBand = oBand;
                pm.bandFmt = band_fmt(j);
                if (pm.bandFmt == 0) {
                    if (ctx->ntStores) pack_band_nt64(reinterpret_cast<uint32_t *>(h + oBand), j.LB, j.RB, j.M + 1);
                    else pack_band(reinterpret_cast<uint32_t *>(h + oBand), j.LB, j.RB, j.M + 1);
                } else {
                    memcpy(h + oBand, j.LB, (size_t)(j.M + 1) * 4);
                    memcpy(h + oBand + (size_t)(j.M + 1) * 4, j.RB, (size_t)(j.M + 1) * 4);
                }
                pm.offSched = oSched;
                pm.nSteps = ji.nSteps;
                pm.lgLanes = lg_of(32 * kBin[ji.bin].G);
                pm.cls = ji.cls;
                pm.rowBase = of.row;
                pm.colBase = of.col;
                pm.scriptBase = of.script;
                metas[i] = pm;
                if (ctx->ntStores) _mm_sfence();
                {
                    const int lg = 63 - __builtin_clzll((unsigned long long)std::max<int64_t>(ji.cells, 1));
                    const int frac = lg >= 2 ? (int)((ji.cells >> (lg - 2)) & 3) : 0;
                    ji.bucket = ji.bin * NB + (NB - 1 - std::min(NB - 1, lg * 4 + frac));
                }
            }
        });
        if (part + 1 < nParts && partHi > partLo) {
            const size_t a = dataOff + s.off[(size_t)partLo].blob;
            const size_t z = partHi < count ? dataOff + s.off[(size_t)partHi].blob : dataOff + blob;
            CUDA_TRY(d, cudaMemcpyAsync(static_cast<unsigned char *>(s.dIn.p) + a, h + a, z - a, cudaMemcpyHostToDevice, s.stream));
            if (s.sentHi == 0) s.sentLo = a;
            s.sentHi = z;
        }
        partLo = partHi;
    }

    // ---- traceback offsets, wave cut, launch order (serial, O(count)) -----------------------------------------------
    // launch order: per ring bin, big pairs first (longest-processing-time-first on the warp queue); quarter-octave
    // buckets of the cell count instead of a comparison sort
    const double t2 = now_ms();
    d.t_par += t2 - t1;
    s.bucketCount.assign((size_t)NBINS * NB, 0);
    size_t tb = 0;
    int64_t kept = count;
    unsigned long long *tbBase = reinterpret_cast<unsigned long long *>(h + s.tbBaseOff);
    for (int64_t i = 0; i < count; ++i) {
        const JobInfo &ji = s.info[(size_t)i];
        if (ji.status != YB_OK) continue;
        const size_t need = (size_t)ji.nSteps * 32 * kBin[ji.bin].G;       // one byte per lane and step
        if (i > 0 && tb + need > maxTb) { kept = i; break; }
        tbBase[i] = tb;
        tb += need;
        s.bucketCount[(size_t)ji.bucket]++;
        if (!ji.connected) s.ungated = false;
    }
    count = kept;
    s.count = count;
    if (count < (int64_t)s.off.size()) {           // wave cut short: sizes of the kept prefix
        const Slot::Off &of = s.off[(size_t)count];
        blob = of.blob; rows = of.row; cols = of.col; words = of.script;
    }
    s.blobBytes = dataOff + blob;
    s.scriptWords = words;
    {
        int *order = reinterpret_cast<int *>(h + s.orderOff);
        int acc = 0;
        for (int b = 0; b < NBINS; ++b) {
            s.binStart[b] = acc;
            for (int q = 0; q < NB; ++q) { int c = s.bucketCount[(size_t)b * NB + q]; s.bucketCount[(size_t)b * NB + q] = acc; acc += c; }
        }
        s.binStart[NBINS] = acc;
        s.nValid = acc;
        int *longList = reinterpret_cast<int *>(h + s.longOff);
        s.nLong = 0;
        s.tbLong = ctx->tbLong;
        for (int64_t i = 0; i < count; ++i) {
            const JobInfo &ji = s.info[(size_t)i];
            if (ji.status != YB_OK) continue;
            order[s.bucketCount[(size_t)ji.bucket]++] = (int)i;
            if (jobs[first + i].M + jobs[first + i].N >= s.tbLong) longList[s.nLong++] = (int)i;
        }
    }
    const double t3 = now_ms();
    d.t_post += t3 - t2;
    CUDA_TRY(d, s.dIn.reserve(s.blobBytes));
    CUDA_TRY(d, s.dRow.reserve(rows * sizeof(RowRec) + 64));
    {   // column records sit between two COL_PAD margins (see fill_body2); a fresh buffer is zeroed once so that what idle
        // lanes read there is initialised memory
        const void *before = s.dCol.p;
        CUDA_TRY(d, s.dCol.reserve(cols * sizeof(ColRec) + 2 * COL_PAD + 64));
        if (s.dCol.p != before) CUDA_TRY(d, cudaMemsetAsync(s.dCol.p, 0, s.dCol.cap, s.stream));
    }
    CUDA_TRY(d, s.dTb.reserve(tb + 256));
    CUDA_TRY(d, s.dScript.reserve(words * 4 + 64));
    CUDA_TRY(d, s.dOut.reserve((size_t)count * sizeof(PairOut) + 64));
    CUDA_TRY(d, s.dQueue.reserve(64));
    s.scriptDst = ctx->scriptStore + ctx->scriptOff[(size_t)first];   // the wave's scripts, in job order, in the batch store
    CUDA_TRY(d, s.hOut.reserve((size_t)count * sizeof(PairOut) + 64));
    d.t_reserve += now_ms() - t3;
    d.pack_ms += now_ms() - t0;
    s.tPack1 = now_ms();
    return YB_OK;
}

// Enqueue one wave on its slot's stream.  Nothing here waits for the device.
// Enqueue H2D, K1 and K2 of one wave on its slot's stream.  Nothing here waits for the device.
int slot_launch_fill(Device &d, Slot &s, bool h2d) {
    const PairMeta *metas = static_cast<const PairMeta *>(s.dIn.p);
    const unsigned char *blob = static_cast<const unsigned char *>(s.dIn.p);
    RowRec *rows = static_cast<RowRec *>(s.dRow.p);
    ColRec *cols = reinterpret_cast<ColRec *>(static_cast<unsigned char *>(s.dCol.p) + COL_PAD);
    unsigned char *tb = static_cast<unsigned char *>(s.dTb.p);
    PairOut *outs = static_cast<PairOut *>(s.dOut.p);
    const int *order = reinterpret_cast<const int *>(blob + s.orderOff);
    int *queue = static_cast<int *>(s.dQueue.p);
    cudaStream_t st = s.stream;
    const double tl = now_ms();

    if (h2d && s.sentHi > s.sentLo) {            // part of the blob is already on its way (slot_pack): header + the rest
        unsigned char *dst = static_cast<unsigned char *>(s.dIn.p);
        const unsigned char *src = static_cast<const unsigned char *>(s.hIn.p);
        CUDA_TRY(d, cudaMemcpyAsync(dst, src, std::min(s.sentLo, s.blobBytes), cudaMemcpyHostToDevice, st));
        if (s.blobBytes > s.sentHi)
            CUDA_TRY(d, cudaMemcpyAsync(dst + s.sentHi, src + s.sentHi, s.blobBytes - s.sentHi, cudaMemcpyHostToDevice, st));
    } else {
        CUDA_TRY(d, cudaEventRecord(s.ev[0], st));
        if (h2d) CUDA_TRY(d, cudaMemcpyAsync(s.dIn.p, s.hIn.p, s.blobBytes, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(d, cudaEventRecord(s.ev[1], st));
    CUDA_TRY(d, cudaMemsetAsync(outs, 0, (size_t)s.count * sizeof(PairOut), st));
    CUDA_TRY(d, cudaMemsetAsync(queue, 0, 64, st));
    // a script region holds ceil((M+N)/16) words but a path has m_new <= M+N ops: the unwritten tail is copied back too
    if (s.scriptWords) CUDA_TRY(d, cudaMemsetAsync(s.dScript.p, 0, s.scriptWords * 4, st));
    if (s.nValid > 0) {
        yb_profile_kernel<<<(unsigned)s.count, K1_THREADS, 0, st>>>(metas, blob, rows, cols, s.y16 ? 1 : 0, d.sc);
        d.launches++;
    }
    CUDA_TRY(d, cudaEventRecord(s.ev[2], st));
    // The wide-ring bins hold few, long pairs (one warp each): they start first, on their own streams, and the
    // bulk bin fills the machine around them -- launched back to back they would be a serial tail.
    for (int b = NBINS - 1; b >= 0; --b) {
        int n = s.binStart[b + 1] - s.binStart[b];
        if (n <= 0) continue;
        const BinCfg &bc = kBin[b];
        cudaStream_t bs = b == 0 ? st : s.binStream[b];
        if (b > 0) CUDA_TRY(d, cudaStreamWaitEvent(bs, s.ev[2], 0));
        const bool bulk = b >= BULK_BIN0;
        FillFn fn = bulk ? fill_fn2(b) : fill_fn(b, s.y16, s.ungated);
        const size_t smem = bulk ? fill2_smem(b) : fill_smem(b);
        const int blocks = std::min((n + bc.P - 1) / bc.P, d.fillBlocks[b]);
        fn<<<blocks, bc.G * bc.P * 32, smem, bs>>>(metas, order + s.binStart[b], n, queue + b, rows, cols, tb,
                                                    reinterpret_cast<const unsigned long long *>(blob + s.tbBaseOff), outs,
                                                    bulk ? -d.sc.gap_open * (b >= BULK_BIN0 + 2 ? 4 : 1) : d.sc.gap_open, d.sc.gap_ext);
        if (b > 0) CUDA_TRY(d, cudaEventRecord(s.binDone[b], bs));
        d.launches++;
    }
    for (int b = 1; b < NBINS; ++b)
        if (s.binStart[b + 1] - s.binStart[b] > 0) CUDA_TRY(d, cudaStreamWaitEvent(st, s.binDone[b], 0));
    CUDA_TRY(d, cudaEventRecord(s.ev[3], st));
    CUDA_TRY(d, cudaGetLastError());
    s.filled = true;
    d.waves++;
    if (h2d) d.h2d_bytes += (int64_t)s.blobBytes;
    d.t_launch += now_ms() - tl;
    return YB_OK;
}

// Enqueue K3 and the D2H copies of a wave whose fill is already queued.  K3 is bound by the latency of the
// longest pair's pointer chase, not by throughput, and the fill kernels leave it no registers to co-reside with:
// launching it once per GROUP of waves, on their own streams at the same time, keeps that latency from being
// paid once per wave.
int slot_launch_traceback(Device &d, Slot &s, bool d2h) {
    const PairMeta *metas = static_cast<const PairMeta *>(s.dIn.p);
    const unsigned char *blob = static_cast<const unsigned char *>(s.dIn.p);
    unsigned char *tb = static_cast<unsigned char *>(s.dTb.p);
    unsigned *script = static_cast<unsigned *>(s.dScript.p);
    PairOut *outs = static_cast<PairOut *>(s.dOut.p);
    const int *order = reinterpret_cast<const int *>(blob + s.orderOff);
    cudaStream_t st = s.stream;
    const double tl = now_ms();
    s.tTbLaunch = tl;
    CUDA_TRY(d, cudaEventRecord(s.ev[6], st));
    // the warp-per-path kernel (issue-bound) runs beside the thread-per-pair kernel (latency-bound), on a side stream
    if (s.nLong > 0) {
        cudaStream_t ls = s.binStream[1];
        CUDA_TRY(d, cudaStreamWaitEvent(ls, s.ev[6], 0));
        yb_traceback_long_kernel<<<(unsigned)((s.nLong + 3) / 4), 128, 0, ls>>>(
            metas, reinterpret_cast<const int *>(blob + s.longOff), s.nLong, blob, tb,
            reinterpret_cast<const unsigned long long *>(blob + s.tbBaseOff), script, outs);
        CUDA_TRY(d, cudaEventRecord(s.binDone[1], ls));
        d.launches++;
    }
    if (s.nValid > s.nLong) {
        yb_traceback_kernel<<<(unsigned)((s.nValid + 127) / 128), 128, 0, st>>>(
            metas, order, s.nValid, blob, tb, reinterpret_cast<const unsigned long long *>(blob + s.tbBaseOff), script, outs,
            s.tbLong);
        d.launches++;
    }
    if (s.nLong > 0) CUDA_TRY(d, cudaStreamWaitEvent(st, s.binDone[1], 0));
    CUDA_TRY(d, cudaEventRecord(s.ev[4], st));
    if (d2h) {
        CUDA_TRY(d, cudaMemcpyAsync(s.hOut.p, s.dOut.p, (size_t)s.count * sizeof(PairOut), cudaMemcpyDeviceToHost, st));
        if (s.scriptWords)
            CUDA_TRY(d, cudaMemcpyAsync(s.scriptDst, s.dScript.p, s.scriptWords * 4, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(d, cudaEventRecord(s.ev[5], st));
    CUDA_TRY(d, cudaGetLastError());
    s.filled = false;
    s.busy = true;
    if (d2h) d.d2h_bytes += (int64_t)((size_t)s.count * sizeof(PairOut) + s.scriptWords * 4);
    d.t_launch += now_ms() - tl;
    return YB_OK;
}

int slot_d2h(Device &d, Slot &s) {
    CUDA_TRY(d, cudaMemcpyAsync(s.hOut.p, s.dOut.p, (size_t)s.count * sizeof(PairOut), cudaMemcpyDeviceToHost, s.stream));
    if (s.scriptWords)
        CUDA_TRY(d, cudaMemcpyAsync(s.scriptDst, s.dScript.p, s.scriptWords * 4, cudaMemcpyDeviceToHost, s.stream));
    d.d2h_bytes += (int64_t)((size_t)s.count * sizeof(PairOut) + s.scriptWords * 4);
    return YB_OK;
}

// Wait for the slot's wave and add its device times to the statistics.
int slot_wait(Device &d, Slot &s) {
    const double tw = now_ms();
    CUDA_TRY(d, cudaStreamSynchronize(s.stream));
    d.t_wait += now_ms() - tw;
    CUDA_TRY(d, cudaGetLastError());
    float h = 0, a = 0, b = 0, c = 0, e = 0;
    cudaEventElapsedTime(&h, s.ev[0], s.ev[1]);
    cudaEventElapsedTime(&a, s.ev[1], s.ev[2]);
    cudaEventElapsedTime(&b, s.ev[2], s.ev[3]);
    cudaEventElapsedTime(&c, s.ev[6], s.ev[4]);
    cudaEventElapsedTime(&e, s.ev[4], s.ev[5]);
    d.h2d_ms += h; d.profile_ms += a; d.fill_ms += b; d.tb_ms += c; d.d2h_ms += e;
    d.kernel_ms += a + b + c;
    if (d.timeline && d.evBase) {           // YB_PROFILE=2: where this wave sat on the device's and the host's clocks
        float t[7] = {0};
        const int idx[7] = {0, 1, 2, 3, 6, 4, 5};
        for (int k = 0; k < 7; ++k) cudaEventElapsedTime(&t[k], d.evBase, s.ev[idx[k]]);
        fprintf(stderr, "yama_b200[timeline] dev %d wave %3d pairs %6lld | host: pack %.2f-%.2f tb-launch %.2f wait-done %.2f | device: h2d %.2f-%.2f "
                "K1 -%.2f K2 -%.2f | K3 %.2f-%.2f d2h -%.2f\n", d.id, d.waves, (long long)s.count, s.tPack0 - d.tBase, s.tPack1 - d.tBase,
                s.tTbLaunch - d.tBase, now_ms() - d.tBase, t[0], t[1], t[2], t[3], t[4], t[5], t[6]);
    }
    s.busy = false;
    return YB_OK;
}

// Scores + edit scripts of a finished wave -> results / the context's script store.
void slot_unpack(yb_ctx *ctx, Device &d, Slot &s, yb_result *results) {
    const double t0 = now_ms();
    const PairOut *outs = static_cast<const PairOut *>(s.hOut.p);
    std::atomic<int64_t> cells{0}, failed{0};
    d.pool->run(s.count, 256, [&](int64_t lo, int64_t hi) {
        int64_t c = 0, f = 0;
        for (int64_t i = lo; i < hi; ++i) {
            const int64_t g = s.first + i;
            yb_result &r = results[g];
            const JobInfo &ji = s.info[(size_t)i];
            memset(&r, 0, sizeof r);
            r.status = ji.status;
            r.cells = ji.cells;
            if (ji.status != YB_OK) { ++f; continue; }
            c += ji.cells;
            const PairOut &o = outs[i];
            r.status = o.status;
            if (o.status != YB_OK) ++f;
            r.m_new = o.m_new; r.C = o.C; r.D = o.D; r.I = o.I;
            // the device's 2-bit codes are the ABI's script format and the D2H copy put them in place
            r.script = ctx->scriptStore + ctx->scriptOff[(size_t)g];
        }
        cells += c; failed += f;
    });
    d.cells += cells.load();
    d.failed += failed.load();
    d.unpack_ms += now_ms() - t0;
}

void reset_stats(Device &d) {
    d.kernel_ms = d.fill_ms = d.profile_ms = d.tb_ms = d.h2d_ms = d.d2h_ms = d.pack_ms = d.unpack_ms = 0;
    d.h2d_bytes = d.d2h_bytes = d.cells = d.failed = 0;
    d.launches = d.waves = 0;
    d.err.clear();
    d.errJob = 0;
    d.t_layout = d.t_par = d.t_post = d.t_reserve = d.t_wait = d.t_launch = 0;
}

void collect_stats(yb_ctx *ctx, yb_stats *st, double total_ms, int64_t cells, int64_t pairs) {
    if (!st) return;
    memset(st, 0, sizeof *st);
    for (auto &d : ctx->devs) {
        st->kernel_ms = std::max(st->kernel_ms, d.kernel_ms);
        st->h2d_ms = std::max(st->h2d_ms, d.h2d_ms);
        st->d2h_ms = std::max(st->d2h_ms, d.d2h_ms);
        st->pack_ms = std::max(st->pack_ms, d.pack_ms + d.unpack_ms);
        st->h2d_bytes += d.h2d_bytes;
        st->d2h_bytes += d.d2h_bytes;
        st->kernel_launches += d.launches;
    }
    st->fill_ms = ctx->devs[0].fill_ms;
    st->profile_ms = ctx->devs[0].profile_ms;
    st->traceback_ms = ctx->devs[0].tb_ms;
    st->total_ms = total_ms;
    st->cells = cells;
    st->pairs = pairs;
    st->n_devices = (int)ctx->devs.size();
}

// contiguous, cell-balanced split of [0,n) into nparts ranges (devices of one context, or ranks)
constexpr int64_t kPairOverheadCells = 2000;   // fixed cost of one pair (row-0 setup, queue pop) in cell units
void plan_split(int64_t n, const int64_t *cells, int nparts, int64_t *cut) {
    for (int k = 0; k <= nparts; ++k) cut[k] = n;
    cut[0] = 0;
    long double total = 0;
    for (int64_t i = 0; i < n; ++i) total += (long double)std::max<int64_t>(cells[i], 0) + kPairOverheadCells;
    long double acc = 0;
    int d = 1;
    for (int64_t i = 0; i < n && d < nparts; ++i) {
        const long double before = acc;
        acc += (long double)std::max<int64_t>(cells[i], 0) + kPairOverheadCells;
        while (d < nparts && acc >= total * d / nparts) {
            // boundary d goes to whichever side of job i is nearer to the ideal cost d/nparts
            const long double target = total * d / nparts;
            cut[d++] = (target - before < acc - target) ? i : i + 1;
        }
    }
    for (int k = 1; k <= nparts; ++k) cut[k] = std::max(cut[k], cut[k - 1]);
    cut[nparts] = n;
}

// Hands out waves: contiguous job ranges whose input fits one staging slot (sizes known from dimensions).
// Waves start small (the device idles while the first one is packed), grow to maxBytes, and shrink again
// towards the end of the batch (the host idles while the last one is on the device).
constexpr int64_t kWavePairsWanted = 256;
struct Dispatcher {
    const yb_job *jobs = nullptr;
    const uint64_t *prefix = nullptr;      // input bytes of jobs [0, i), n + 1 entries (prepare_script_store)
    int64_t n = 0, cursor = 0;
    size_t maxBytes = 0, minBytes = 0, tailBytes = 0, remaining = 0;
    int64_t maxPairs = 0;
    int ndev = 1, handed = 0;
    std::mutex mu;
    static size_t bytes_of(const yb_job &j) {
        size_t b = sizeof(PairMeta) + 16;
        if (j.K >= 1 && j.L >= 1 && j.M >= 1 && j.N >= 1) b += blob_bytes(j);
        return b;
    }
    bool grab(int64_t &lo, int64_t &hi) {
        std::lock_guard<std::mutex> g(mu);
        if (cursor >= n) return false;
        lo = cursor;
        const int round = handed / ndev;
        size_t target = std::min(maxBytes, minBytes << std::min(round, 16));           // ramp up
        target = std::min(target, std::max(tailBytes, remaining / (size_t)(2 * ndev))); // ramp down
        // Large pairs (wide bands, deep profiles: a megabyte each) run one CTA per pair, and the device wants a few
        // hundred of them in flight: such a wave is sized by pairs, up to 8x the byte target (cfg5: +50 % end to end).
        const int64_t wantPairs = std::min<int64_t>(kWavePairsWanted, (int64_t)32 << std::min(round, 3));
        const size_t hardBytes = 8 * maxBytes;
        size_t bytes = 0;
        int64_t i = lo;
        if (prefix) {
            // first job whose inclusion would pass the byte target, then the pairs rule, then the hard cap
            auto upto = [&](size_t lim) {        // largest i with prefix[i] - prefix[lo] <= lim
                return (int64_t)(std::upper_bound(prefix + lo, prefix + n + 1, prefix[lo] + lim) - prefix) - 1;
            };
            i = std::max(lo + 1, upto(target));
            if (i - lo < wantPairs) i = std::max(i, std::min(lo + wantPairs, std::max(lo + 1, upto(hardBytes))));
            i = std::min(std::min(i, n), lo + maxPairs);
            bytes = (size_t)(prefix[i] - prefix[lo]);
        } else
        while (i < n && i - lo < maxPairs) {
            size_t b = bytes_of(jobs[i]);
            if (i > lo && bytes + b > target && (i - lo >= wantPairs || bytes + b > hardBytes)) break;
            bytes += b;
            ++i;
        }
        hi = cursor = i;
        remaining -= std::min(remaining, bytes);
        ++handed;
        return true;
    }
};

// One device's share of a batch: grab waves until none are left.  A wave's fill (H2D, K1, K2) is queued as soon as
// it is packed; tracebacks and D2H copies are queued for TB_GROUP waves at a time; a slot is unpacked when the
// ring comes back to it (NSLOTS waves later) or at the end.
int device_run(yb_ctx *ctx, Device &d, Dispatcher &disp, yb_result *results) {
    if (cudaSetDevice(d.id) != cudaSuccess) { d.err = "cudaSetDevice failed"; return YB_ERR_CUDA; }
    int rc = YB_OK, next = 0, nFilled = 0;
    int filledSlots[NSLOTS];
    if (const char *e = getenv("YB_PROFILE")) d.timeline = atoi(e) >= 2;
    if (d.timeline) {
        if (!d.evBase) cudaEventCreate(&d.evBase);
        d.tBase = now_ms();
        cudaEventRecord(d.evBase, d.slots[0].stream);
    }
    int64_t lo = 0, hi = 0;                      // jobs grabbed but not yet packed
    auto flush_group = [&]() -> int {
        for (int k = 0; k < nFilled; ++k) {
            int r2 = slot_launch_traceback(d, d.slots[filledSlots[k]], true);
            if (r2 != YB_OK) return r2;
        }
        nFilled = 0;
        return YB_OK;
    };
    for (;;) {
        if (lo >= hi && !disp.grab(lo, hi)) break;
        Slot &s = d.slots[next];
        if (s.filled && (rc = flush_group()) != YB_OK) break;     // (only if NSLOTS < 2*TB_GROUP)
        if (s.busy) {                            // the wave launched NSLOTS rounds ago
            if ((rc = slot_wait(d, s)) != YB_OK) break;
            slot_unpack(ctx, d, s, results);
        }
        int64_t count = hi - lo;
        if ((rc = slot_pack(ctx, d, s, disp.jobs, lo, count, ctx->waveTbBytes, ctx->earlyH2D)) != YB_OK) break;
        lo += count;
        if ((rc = slot_launch_fill(d, s, true)) != YB_OK) break;
        filledSlots[nFilled++] = next;
        next = (next + 1) % NSLOTS;
        if (nFilled == TB_GROUP && (rc = flush_group()) != YB_OK) break;
    }
    if (rc == YB_OK) rc = flush_group();
    for (int k = 0; k < NSLOTS; ++k) {           // drain, oldest first
        Slot &s = d.slots[(next + k) % NSLOTS];
        if (s.filled) { cudaStreamSynchronize(s.stream); s.filled = false; }
        if (!s.busy) continue;
        int r2 = slot_wait(d, s);
        if (r2 != YB_OK) { if (rc == YB_OK) rc = r2; continue; }
        if (rc == YB_OK) slot_unpack(ctx, d, s, results);
    }
    return rc;
}

int prepare_script_store(yb_ctx *ctx, int64_t n, const yb_job *jobs) {
    // per-job offsets into the script store (packed scripts, whole words: the layout of the waves' device script pools)
    // and the batch's input bytes: a prefix sum over all jobs, done in chunks on the helper threads -- it runs before
    // the first wave can be packed, i.e. while the device idles
    ctx->scriptOff.resize((size_t)n);
    ctx->blobPrefix.resize((size_t)n + 1);
    constexpr int64_t CH = 4096;
    const int64_t nch = (n + CH - 1) / CH;
    std::vector<size_t> chTot((size_t)nch + 1, 0), chBlob((size_t)nch + 1, 0);
    Pool &pool = *ctx->devs[0].pool;
    pool.run(nch, 1, [&](int64_t a, int64_t z) {
        for (int64_t c = a; c < z; ++c) {
            size_t t = 0, b = 0;
            for (int64_t i = c * CH, e = std::min(n, (c + 1) * CH); i < e; ++i) {
                if (dims_ok(jobs[i])) t += (((size_t)jobs[i].M + jobs[i].N + 15) / 16) * 4;
                b += Dispatcher::bytes_of(jobs[i]);
            }
            chTot[(size_t)c + 1] = t; chBlob[(size_t)c + 1] = b;
        }
    });
    for (int64_t c = 0; c < nch; ++c) { chTot[(size_t)c + 1] += chTot[(size_t)c]; chBlob[(size_t)c + 1] += chBlob[(size_t)c]; }
    pool.run(nch, 1, [&](int64_t a, int64_t z) {
        for (int64_t c = a; c < z; ++c) {
            size_t t = chTot[(size_t)c], b = chBlob[(size_t)c];
            for (int64_t i = c * CH, e = std::min(n, (c + 1) * CH); i < e; ++i) {
                ctx->scriptOff[(size_t)i] = t;
                ctx->blobPrefix[(size_t)i] = b;
                if (dims_ok(jobs[i])) t += (((size_t)jobs[i].M + jobs[i].N + 15) / 16) * 4;
                b += Dispatcher::bytes_of(jobs[i]);
            }
        }
    });
    const size_t tot = chTot[(size_t)nch], blob = chBlob[(size_t)nch];
    ctx->blobPrefix[(size_t)n] = blob;
    ctx->batchBlobBytes = blob;
    if (tot + 64 > ctx->scriptStoreCap) {
        if (ctx->scriptStore) cudaFreeHost(ctx->scriptStore);
        ctx->scriptStore = nullptr;
        ctx->scriptStoreCap = 0;
        const size_t want = tot + tot / 4 + (1u << 20);
        void *p = nullptr;
        if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) { (void)cudaGetLastError(); return YB_ERR_CUDA; }
        ctx->scriptStore = static_cast<uint8_t *>(p);
        ctx->scriptStoreCap = want;
    }
    return YB_OK;
}

template <class F>
int for_each_device(yb_ctx *ctx, F &&fn) {
    int ndev = (int)ctx->devs.size();
    std::vector<int> rcs((size_t)ndev, YB_OK);
    if (ndev == 1) rcs[0] = fn(0);
    else {
        std::vector<std::thread> th;
        for (int d = 0; d < ndev; ++d) th.emplace_back([&, d] { rcs[(size_t)d] = fn(d); });
        for (auto &t : th) t.join();
    }
    for (int d = 0; d < ndev; ++d)
        if (rcs[(size_t)d] != YB_OK) {
            ctx->err = ctx->devs[(size_t)d].err.empty() ? "device failure" : ctx->devs[(size_t)d].err;
            return rcs[(size_t)d];
        }
    return YB_OK;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int yb_create(const int *devices, int ndev, yb_ctx **out) {
    if (!out) return YB_ERR_ARG;
    *out = nullptr;
    int avail = 0;
    const double tc0 = now_ms();
    if (cudaGetDeviceCount(&avail) != cudaSuccess || avail < 1) return YB_ERR_CUDA;
    const double tc1 = now_ms();
    yb_ctx *ctx = new yb_ctx();
    std::vector<int> ids;
    if (devices && ndev > 0) ids.assign(devices, devices + ndev);
    else for (int i = 0; i < avail; ++i) ids.push_back(i);
    for (int id : ids) {
        if (id < 0 || id >= avail) { delete ctx; return YB_ERR_ARG; }
        ctx->devs.emplace_back();
        ctx->devs.back().id = id;
    }
    for (auto &d : ctx->devs)
        if (device_init(d) != YB_OK) { fprintf(stderr, "yama_b200: %s\n", d.err.c_str()); yb_destroy(ctx); return YB_ERR_CUDA; }
    if (getenv("YB_PROFILE"))
        fprintf(stderr, "yama_b200[profile] start-up: driver %.0f ms, contexts + streams + kernel attributes %.0f ms\n", tc1 - tc0, now_ms() - tc1);
    int hw = (int)std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    ctx->nThreads = std::min(hw, 32);
    if (const char *e = getenv("YB_THREADS")) ctx->nThreads = std::max(1, atoi(e));
    if (const char *e = getenv("YB_WAVE_MB")) ctx->waveInBytes = std::max<size_t>(1, (size_t)strtoull(e, nullptr, 10)) << 20;
    if (const char *e = getenv("YB_WAVE_MIN_MB")) ctx->waveMinBytes = std::max<size_t>(1, (size_t)strtoull(e, nullptr, 10)) << 20;
    if (const char *e = getenv("YB_WAVE_TAIL_MB")) ctx->waveTailBytes = std::max<size_t>(1, (size_t)strtoull(e, nullptr, 10)) << 20;
    {   // traceback bytes per wave: an eighth of the smallest device's free memory (NSLOTS waves can be in flight)
        size_t cap = (size_t)24 << 30;
        for (auto &d : ctx->devs) {
            size_t fr = 0, tot = 0;
            if (cudaSetDevice(d.id) == cudaSuccess && cudaMemGetInfo(&fr, &tot) == cudaSuccess) cap = std::min(cap, fr / 8);
        }
        ctx->waveTbBytes = std::max<size_t>(cap, (size_t)256 << 20);
    }
    if (const char *e = getenv("YB_WAVE_TB_MB")) ctx->waveTbBytes = std::max<size_t>(1, (size_t)strtoull(e, nullptr, 10)) << 20;
    if (const char *e = getenv("YB_NT")) ctx->ntStores = atoi(e) != 0;
    if (const char *e = getenv("YB_EARLY_H2D")) ctx->earlyH2D = atoi(e) != 0;
    if (const char *e = getenv("YB_UNGATED")) ctx->ungatedOk = atoi(e) != 0;
    if (const char *e = getenv("YB_KEYED")) if (atoi(e) == 0) ctx->maxCls = std::min(ctx->maxCls, 1);
    if (const char *e = getenv("YB_FILL2")) if (atoi(e) == 0) ctx->maxCls = 0;
    if (const char *e = getenv("YB_TB_LONG")) ctx->tbLong = std::max(1, atoi(e));
    if (const char *e = getenv("YB_WAVE_PAIRS")) ctx->wavePairs = std::max<int64_t>(1, atoll(e));
    for (auto &d : ctx->devs) {
        d.helpers = std::max(1, ctx->nThreads / (int)ctx->devs.size());
        d.pool.reset(new Pool(d.helpers - 1));
    }
    *out = ctx;
    return YB_OK;
}

void yb_destroy(yb_ctx *ctx) {
    if (!ctx) return;
    for (auto &d : ctx->devs) {
        cudaSetDevice(d.id);
        for (auto &s : d.slots) {
            for (DevBuf *b : {&s.dIn, &s.dRow, &s.dCol, &s.dTb, &s.dScript, &s.dOut, &s.dQueue}) b->release();
            for (PinBuf *b : {&s.hIn, &s.hOut}) b->release();
            for (auto &e : s.ev) if (e) cudaEventDestroy(e);
            for (auto &e : s.binDone) if (e) cudaEventDestroy(e);
            for (auto &b : s.binStream) if (b) cudaStreamDestroy(b);
            if (s.stream) cudaStreamDestroy(s.stream);
        }
        for (auto &g : d.score) {
            g.dIn.release(); g.dSums.release(); g.hIn.release(); g.hSums.release();
            for (auto &e : g.ev) if (e) cudaEventDestroy(e);
            if (g.stream) cudaStreamDestroy(g.stream);
        }
    }
    if (ctx->scriptStore) cudaFreeHost(ctx->scriptStore);
    delete ctx;
}

const char *yb_last_error(const yb_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }
int yb_device_count(const yb_ctx *ctx) { return ctx ? (int)ctx->devs.size() : 0; }

int yb_set_scores(yb_ctx *ctx, const int32_t *ss, const int32_t *gop, int32_t gap_extend) {
    if (!ctx || !ss || !gop) return YB_ERR_ARG;
    static const unsigned char rep[6] = {'A', 'C', 'G', 'T', 'N', '-'};
    auto cls = [](int ch) {
        int u = ch | 0x20;
        if (ch == '-') return 5;
        if (u == 'a') return 0;
        if (u == 'c') return 1;
        if (u == 'g') return 2;
        if (u == 't') return 3;
        return 4;
    };
    ScoreConst sc;
    int maxabs = 1;
    for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b) {
            sc.S6[a][b] = ss[128 * rep[a] + rep[b]];
            maxabs = std::max(maxabs, std::abs(sc.S6[a][b]));
        }
    for (int c = 0; c < 128; ++c)
        for (int d = 0; d < 128; ++d)
            if (ss[128 * c + d] != sc.S6[cls(c)][cls(d)]) {
                set_err(ctx, "ss[%d][%d]=%d does not follow the 6-class structure of init_scores (expected %d)", c, d, ss[128 * c + d], sc.S6[cls(c)][cls(d)]);
                return YB_ERR_SCORES;
            }
    const int GO = gop[1];
    for (int i = 0; i < 16; ++i) {
        bool open = (i == 1 || i == 2 || i == 6 || i == 9 || i == 13 || i == 14);   // mz_scores.c:78-79
        if (gop[i] != (open ? GO : 0)) {
            set_err(ctx, "gop[%d]=%d does not follow the quasi-natural pattern of init_scores", i, gop[i]);
            return YB_ERR_SCORES;
        }
    }
    if (GO < 0 || GO > 32767) { set_err(ctx, "gap_open %d outside [0,32767]", GO); return YB_ERR_SCORES; }
    if (gap_extend < 0 || gap_extend > 32767) { set_err(ctx, "gap_extend %d outside [0,32767]", gap_extend); return YB_ERR_SCORES; }
    sc.gap_open = GO;
    sc.gap_ext = gap_extend;
    ctx->sc = sc;
    // 16-bit weights in the kernels: sum-of-pairs weights K*max|S6| and the extension weight K*gap_extend
    ctx->maxDepth = std::min(255, 32767 / std::max(maxabs, std::max(1, (int)gap_extend)));
    ctx->maxAbsS = maxabs;
    for (auto &d : ctx->devs) d.sc = sc;
    ctx->scoresSet = true;
    return YB_OK;
}

int yb_pair_facts(const yb_job *job, int64_t *cells, int32_t *wmax, int32_t *nsteps, char *msg, int msglen) {
    if (!job || job->M < 1 || job->N < 1 || !job->LB || !job->RB) return YB_ERR_ARG;
    const int nblk = (job->M + 31) >> 5;
    std::vector<int> s1((size_t)nblk + 1, -1), s2((size_t)nblk + 1, -1);
    int w1 = 0, w2 = 0, n1 = 0;
    int lanes1 = 0;
    const int64_t c1 = yb_band_scan(job->M, job->N, job->LB, job->RB, &w1, s1.data(), &n1, lanes_of, &lanes1, nullptr);
    const int64_t c2 = check_band(job->M, job->N, job->LB, job->RB, msg, msglen, &w2);
    if (c2 < 0) return c1 < 0 ? YB_ERR_BAND : YB_ERR_LIMIT;
    const int lanes2 = lanes_of(w2, job->M);
    if (lanes2 <= 0) return YB_ERR_LIMIT;                      // wider than any kernel bin
    const int n2 = make_schedule(job->M, job->LB, job->RB, lanes2, s2.data());
    for (int b = (job->M + lanes2 - 1) / lanes2; b <= nblk; ++b) s1[(size_t)b] = s2[(size_t)b] = -1;
    if (c1 != c2 || w1 != w2 || n1 != n2 || lanes1 != lanes2 || s1 != s2) return YB_ERR_LIMIT;
    if (cells) *cells = c2;
    if (wmax) *wmax = w2;
    if (nsteps) *nsteps = n2;
    return YB_OK;
}

int yb_plan_split(int64_t n, const int64_t *cells, int nparts, int64_t *cuts) {
    if (n < 0 || nparts < 1 || !cuts || (n > 0 && !cells)) return YB_ERR_ARG;
    plan_split(n, cells, nparts, cuts);
    return YB_OK;
}

int64_t yb_check_band(int32_t M, int32_t N, const int32_t *LB, const int32_t *RB, char *msg, int msglen) {
    if (M < 0 || N < 0 || !LB || !RB) return YB_ERR_ARG;
    return check_band(M, N, LB, RB, msg, msglen, nullptr);
}

int yb_run_batch(yb_ctx *ctx, int64_t n, const yb_job *jobs, yb_result *results, yb_stats *stats) {
    if (!ctx || n < 0 || (n > 0 && (!jobs || !results))) return YB_ERR_ARG;
    if (!ctx->scoresSet) { set_err(ctx, "yb_set_scores has not been called"); return YB_ERR_SCORES; }
    const double t0 = now_ms();
    if (prepare_script_store(ctx, n, jobs) != YB_OK) { set_err(ctx, "cudaHostAlloc failed for the script store"); return YB_ERR_CUDA; }
    const double tPrepared = now_ms();
    for (auto &d : ctx->devs) {
        reset_stats(d);
        if (d.hasResident) {        // a resident batch sized slot 0 for the WHOLE batch: give that memory back before waves
            Slot &s = d.slots[0];
            if (cudaSetDevice(d.id) == cudaSuccess) {
                for (DevBuf *b : {&s.dIn, &s.dRow, &s.dCol, &s.dTb, &s.dScript, &s.dOut}) b->release();
                s.hIn.release();
            }
            d.hasResident = false;
        }
    }
    Dispatcher disp;
    disp.jobs = jobs; disp.n = n;
    disp.prefix = ctx->blobPrefix.data();
    disp.maxBytes = ctx->waveInBytes; disp.maxPairs = ctx->wavePairs;
    disp.minBytes = std::min(ctx->waveInBytes, ctx->waveMinBytes);
    disp.tailBytes = std::min(ctx->waveInBytes, ctx->waveTailBytes);
    disp.ndev = (int)ctx->devs.size();
    disp.remaining = ctx->batchBlobBytes;
    int rc = for_each_device(ctx, [&](int d) { return device_run(ctx, ctx->devs[(size_t)d], disp, results); });
    const double tRan = now_ms();
    int64_t cells = 0;
    for (auto &d : ctx->devs) cells += d.cells;
    collect_stats(ctx, stats, now_ms() - t0, cells, n);
    if (getenv("YB_PROFILE"))
        for (auto &d : ctx->devs)
            fprintf(stderr, "yama_b200[profile] prepare %.2f ms, devices %.2f ms\n", tPrepared - t0, tRan - tPrepared),
            fprintf(stderr, "yama_b200[profile] dev %d: waves %d total %.2f ms | pack %.2f (layout %.2f, analyse+copy %.2f, order %.2f, reserve %.2f) "
                    "launch %.2f wait %.2f unpack %.2f | device: h2d %.2f kernels %.2f d2h %.2f\n", d.id, d.waves, now_ms() - t0, d.pack_ms,
                    d.t_layout, d.t_par, d.t_post, d.t_reserve, d.t_launch, d.t_wait, d.unpack_ms, d.h2d_ms, d.kernel_ms, d.d2h_ms);
    if (rc != YB_OK) return rc;
    // per-pair failures: report the first in job order, in the reference's wording where it has one
    int64_t failed = 0;
    for (auto &d : ctx->devs) failed += d.failed;
    if (failed == 0) return YB_OK;                  // (the usual case: no scan of the results)
    for (int64_t i = 0; i < n; ++i)
        if (results[i].status != YB_OK) {
            bool have = false;
            for (auto &d : ctx->devs) if (!d.err.empty()) { ctx->err = d.err; have = true; break; }
            if (!have || results[i].status == YB_ERR_TRACEBACK) set_err(ctx, "Error generating edit script.");
            return results[i].status;
        }
    return YB_OK;
}

int yb_resident_load(yb_ctx *ctx, int64_t n, const yb_job *jobs) {
    if (!ctx || n < 1 || !jobs) return YB_ERR_ARG;
    if (!ctx->scoresSet) { set_err(ctx, "yb_set_scores has not been called"); return YB_ERR_SCORES; }
    ctx->resJobs.assign(jobs, jobs + n);
    if (prepare_script_store(ctx, n, jobs) != YB_OK) { set_err(ctx, "cudaHostAlloc failed for the script store"); return YB_ERR_CUDA; }
    for (auto &d : ctx->devs) { reset_stats(d); d.hasResident = false; }
    // static, cell-balanced split: needs every pair's cell count first
    std::vector<int64_t> cells((size_t)n, 0);
    std::atomic<int> bad{0};
    parallel_for(ctx->nThreads, n, 256, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            JobInfo ji;
            analyse_one(ctx, jobs[i], ji, nullptr, nullptr, 0);
            cells[(size_t)i] = ji.cells;
            if (ji.status != YB_OK) bad.store(1);
        }
    });
    if (bad.load()) { set_err(ctx, "yb_resident_load: the batch holds invalid pairs (use yb_run_batch for per-pair status)"); return YB_ERR_ARG; }
    ctx->resCells = 0;
    for (auto c : cells) ctx->resCells += c;
    const int ndev = (int)ctx->devs.size();
    ctx->resSplit.assign((size_t)ndev + 1, 0);
    plan_split(n, cells.data(), ndev, ctx->resSplit.data());
    return for_each_device(ctx, [&](int di) -> int {
        Device &d = ctx->devs[(size_t)di];
        if (cudaSetDevice(d.id) != cudaSuccess) return (int)YB_ERR_CUDA;
        int64_t lo = ctx->resSplit[(size_t)di], cnt = ctx->resSplit[(size_t)di + 1] - lo;
        if (cnt <= 0) return (int)YB_OK;
        Slot &s = d.slots[0];
        const int64_t want = cnt;
        int rc = slot_pack(ctx, d, s, ctx->resJobs.data(), lo, cnt, (size_t)-1);
        if (rc != YB_OK) return rc;
        if (cnt != want) { d.err = "resident batch does not fit"; return (int)YB_ERR_LIMIT; }
        CUDA_TRY(d, cudaMemcpyAsync(s.dIn.p, s.hIn.p, s.blobBytes, cudaMemcpyHostToDevice, s.stream));
        CUDA_TRY(d, cudaStreamSynchronize(s.stream));
        d.hasResident = true;
        return (int)YB_OK;
    });
}

int yb_resident_step(yb_ctx *ctx, yb_stats *stats) {
    if (!ctx) return YB_ERR_ARG;
    double t0 = now_ms();
    for (auto &d : ctx->devs) reset_stats(d);
    int rc = for_each_device(ctx, [&](int di) -> int {
        Device &d = ctx->devs[(size_t)di];
        if (!d.hasResident) {
            if (ctx->resSplit.size() > (size_t)di + 1 && ctx->resSplit[(size_t)di + 1] == ctx->resSplit[(size_t)di]) return (int)YB_OK;
            d.err = "no resident batch loaded";
            return (int)YB_ERR_ARG;
        }
        if (cudaSetDevice(d.id) != cudaSuccess) return (int)YB_ERR_CUDA;
        int r = slot_launch_fill(d, d.slots[0], false);
        if (r == YB_OK) r = slot_launch_traceback(d, d.slots[0], false);
        if (r != YB_OK) return r;
        return slot_wait(d, d.slots[0]);
    });
    collect_stats(ctx, stats, now_ms() - t0, ctx->resCells, (int64_t)ctx->resJobs.size());
    return rc;
}

int yb_resident_fetch(yb_ctx *ctx, yb_result *results) {
    if (!ctx || !results) return YB_ERR_ARG;
    return for_each_device(ctx, [&](int di) -> int {
        Device &d = ctx->devs[(size_t)di];
        if (!d.hasResident) return (int)YB_OK;
        if (cudaSetDevice(d.id) != cudaSuccess) return (int)YB_ERR_CUDA;
        Slot &s = d.slots[0];
        int r = slot_d2h(d, s);
        if (r != YB_OK) return r;
        CUDA_TRY(d, cudaStreamSynchronize(s.stream));
        slot_unpack(ctx, d, s, results);
        return (int)YB_OK;
    });
}

int64_t yb_submit(yb_ctx *ctx, const yb_job *job) {
    if (!ctx || !job) return YB_ERR_ARG;
    if (job->K < 1 || job->L < 1 || job->M < 1 || job->N < 1) return YB_ERR_ARG;
    yb_ctx::QJob q;
    q.K = job->K; q.M = job->M; q.L = job->L; q.N = job->N;
    auto put = [&](const void *src, size_t bytes) {
        size_t off = align_up(ctx->arena.size(), 16);
        ctx->arena.resize(off + bytes);
        memcpy(ctx->arena.data() + off, src, bytes);
        return off;
    };
    q.offA = put(job->A, (size_t)job->K * job->M);
    q.offB = put(job->B, (size_t)job->L * job->N);
    q.offLB = put(job->LB, (size_t)(job->M + 1) * 4);
    q.offRB = put(job->RB, (size_t)(job->M + 1) * 4);
    ctx->queued.push_back(q);
    return (int64_t)ctx->queued.size() - 1;
}

int yb_flush(yb_ctx *ctx, yb_stats *stats) {
    if (!ctx) return YB_ERR_ARG;
    size_t n = ctx->queued.size();
    std::vector<yb_job> jobs(n);
    for (size_t i = 0; i < n; ++i) {
        const auto &q = ctx->queued[i];
        jobs[i].K = q.K; jobs[i].M = q.M; jobs[i].L = q.L; jobs[i].N = q.N;
        jobs[i].A = ctx->arena.data() + q.offA;
        jobs[i].B = ctx->arena.data() + q.offB;
        jobs[i].LB = reinterpret_cast<const int32_t *>(ctx->arena.data() + q.offLB);
        jobs[i].RB = reinterpret_cast<const int32_t *>(ctx->arena.data() + q.offRB);
    }
    ctx->queuedRes.assign(n, yb_result{});
    if (n == 0) { if (stats) memset(stats, 0, sizeof *stats); return YB_OK; }
    return yb_run_batch(ctx, (int64_t)n, jobs.data(), ctx->queuedRes.data(), stats);
}

int yb_fetch(yb_ctx *ctx, int64_t id, yb_result *out) {
    if (!ctx || !out || id < 0 || (size_t)id >= ctx->queuedRes.size()) return YB_ERR_ARG;
    *out = ctx->queuedRes[(size_t)id];
    return out->status;
}

void yb_clear(yb_ctx *ctx) {
    if (!ctx) return;
    ctx->queued.clear();
    ctx->queuedRes.clear();
    ctx->arena.clear();
}

int yb_script_unpack(const yb_result *res, uint8_t *ops) {
    if (!res || !ops || (res->m_new > 0 && !res->script)) return YB_ERR_ARG;
    for (int i = 0; i < res->m_new; ++i) ops[i] = (uint8_t)((res->script[i >> 2] >> (2 * (i & 3))) & 3);
    return YB_OK;
}

int yb_assemble(const yb_job *job, const yb_result *res, uint8_t *out) {
    if (!job || !res || !out || !res->script) return YB_ERR_ARG;
    const int K = job->K, L = job->L, W = K + L;
    int i = 0, j = 0, m = 0;
    for (int e = res->m_new - 1; e >= 0; --e) {             // mz_yama.c:300-309
        const int op = (res->script[e >> 2] >> (2 * (e & 3))) & 3;
        uint8_t *dst = out + (size_t)m * W;
        if (op == FLAG_C) { ++i; ++j; }
        else if (op == FLAG_I) ++j;
        else if (op == FLAG_D) ++i;
        else return YB_ERR_TRACEBACK;
        if (i > job->M || j > job->N) return YB_ERR_TRACEBACK;
        if (op == FLAG_I) memset(dst, '-', (size_t)K); else memcpy(dst, job->A + (size_t)(i - 1) * K, (size_t)K);
        if (op == FLAG_D) memset(dst + K, '-', (size_t)L); else memcpy(dst + K, job->B + (size_t)(j - 1) * L, (size_t)L);
        ++m;
    }
    if (i != job->M || j != job->N) return YB_ERR_TRACEBACK;   // mz_yama.c:310-312
    return YB_OK;
}

// ---- block scoring: mafScoreRange (mz_scores.c:124-152), SURVEY 8(f) rank 1 ------------------------------------
// Blocks are cut into waves of about 64 MB of text; a wave's metas, warp units and text go up in ONE copy from a
// pinned buffer, one kernel scores it, one copy brings the per-block int64 sums back.  Two stages alternate, so the
// helpers pack wave w+1 while wave w is on the device.  Device 0 of the context only: a multi-GPU host shards the
// block list itself (blocks are independent), like the yama jobs.
namespace {
struct ScoreLayout { size_t off; int pitch; int units; };
int score_finish(Device &d, ScoreStage &g, double *scores) {
    if (!g.busy) return YB_OK;
    CUDA_TRY(d, cudaStreamSynchronize(g.stream));
    g.busy = false;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, g.ev[0], g.ev[1]) == cudaSuccess) d.h2d_ms += ms;
    if (cudaEventElapsedTime(&ms, g.ev[1], g.ev[2]) == cudaSuccess) d.kernel_ms += ms;
    if (cudaEventElapsedTime(&ms, g.ev[2], g.ev[3]) == cudaSuccess) d.d2h_ms += ms;
    const long long *sums = static_cast<const long long *>(g.hSums.p);
    for (int64_t i = 0; i < g.count; ++i) scores[g.first + i] = (double)sums[i];   // integer-valued, exact below 2^53
    return YB_OK;
}
}  // namespace

int yb_score_blocks(yb_ctx *ctx, int64_t n, const yb_block *blocks, double *scores, yb_stats *stats) {
    if (!ctx || n < 0 || (n > 0 && (!blocks || !scores))) return YB_ERR_ARG;
    if (!ctx->scoresSet) { set_err(ctx, "mafScoreRange: scores not initialized"); return YB_ERR_SCORES; }   // mz_scores.c:133-134
    for (int a = 0; a < 6; ++a)
        for (int b = 0; b < a; ++b)
            if (ctx->sc.S6[a][b] != ctx->sc.S6[b][a]) {
                set_err(ctx, "yb_score_blocks needs a symmetric substitution matrix (class %d/%d: %d vs %d)", a, b, ctx->sc.S6[a][b], ctx->sc.S6[b][a]);
                return YB_ERR_SCORES;
            }
    const double t0 = now_ms();
    Device &d = ctx->devs[0];
    reset_stats(d);
    if (cudaSetDevice(d.id) != cudaSuccess) { set_err(ctx, "cudaSetDevice failed"); return YB_ERR_CUDA; }
    int maxabs = 1;
    for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) maxabs = std::max(maxabs, std::abs(ctx->sc.S6[a][b]));
    // rows up to which a column's quadratic form fits 32 bits: |column| <= (maxabs + gap_open) * rows^2 / 2
    int rows32 = 1;
    while ((double)(rows32 + 1) * (rows32 + 1) * (maxabs + ctx->sc.gap_open) / 2.0 < 2147483647.0 && rows32 < 46340) ++rows32;

    // validation, in the reference's order and wording (mz_scores.c:130-132)
    int64_t pairCols = 0;
    for (int64_t i = 0; i < n; ++i) {
        const yb_block &b = blocks[i];
        if (b.start < 0 || b.size <= 0 || (int64_t)b.start + b.size > b.text_size) {
            set_err(ctx, "mafScoreRange: start = %d, size = %d, textSize = %d\n", b.start, b.size, b.text_size);
            return YB_ERR_ARG;
        }
        if (b.nrows < 0 || (b.nrows > 0 && !b.rows)) { set_err(ctx, "yb_score_blocks: block %lld has no rows", (long long)i); return YB_ERR_ARG; }
        pairCols += (int64_t)b.nrows * (b.nrows - 1) / 2 * b.size;
    }
    const size_t waveMax = (size_t)64 << 20;
    int rc = YB_OK, which = 0, waveNo = 0;
    std::vector<ScoreLayout> lay;
    std::vector<uint32_t> rowBlock;              // per text row of the wave: its block (wave-relative)
    std::vector<uint32_t> rowFirst;              // per block of the wave: index of its first row
    int64_t lo = 0;
    while (lo < n && rc == YB_OK) {
        // ---- cut a wave ------------------------------------------------------------------------------------
        lay.clear(); rowBlock.clear(); rowFirst.clear();
        size_t text = 0;
        int64_t units = 0, hi = lo;
        const size_t waveBytes = std::min(waveMax, ((size_t)16 << 20) << std::min(waveNo, 4));   // small first waves: the
        ++waveNo;                                                                                // device starts early
        while (hi < n && (hi == lo || text < waveBytes)) {
            const yb_block &b = blocks[hi];
            ScoreLayout L;
            L.pitch = 4 + (int)align_up((size_t)b.size, 4);
            L.off = text;
            L.units = (b.size + SCORE_UNIT_COLS - 1) / SCORE_UNIT_COLS;
            text += align_up((size_t)L.pitch * (size_t)b.nrows, 16);
            units += L.units;
            rowFirst.push_back((uint32_t)rowBlock.size());
            rowBlock.insert(rowBlock.end(), (size_t)b.nrows, (uint32_t)(hi - lo));
            lay.push_back(L);
            ++hi;
        }
        const int64_t cnt = hi - lo;
        if (units > 0x7fffffff / 2) { set_err(ctx, "yb_score_blocks: wave too large"); rc = YB_ERR_LIMIT; break; }
        const size_t metaOff = 0, unitOff = align_up((size_t)cnt * sizeof(ScoreMeta), 64);
        const size_t textOff = align_up(unitOff + (size_t)units * sizeof(ScoreUnit), 256);
        const size_t total = textOff + text + 256;
        ScoreStage &g = d.score[which];
        which ^= 1;
        if ((rc = score_finish(d, g, scores)) != YB_OK) break;
        if (!g.stream) {
            CUDA_TRY(d, cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
            for (auto &e : g.ev) CUDA_TRY(d, cudaEventCreate(&e));
        }
        CUDA_TRY(d, g.hIn.reserve(total));
        CUDA_TRY(d, g.dIn.reserve(total));
        CUDA_TRY(d, g.hSums.reserve((size_t)cnt * 8));
        CUDA_TRY(d, g.dSums.reserve((size_t)cnt * 8));
        // ---- pack: metas + units per block, text per row, on the helpers ---------------------------------------
        const double tp = now_ms();
        unsigned char *h = static_cast<unsigned char *>(g.hIn.p);
        ScoreMeta *metas = reinterpret_cast<ScoreMeta *>(h + metaOff);
        ScoreUnit *un = reinterpret_cast<ScoreUnit *>(h + unitOff);
        {
            int64_t u = 0;
            for (int64_t i = 0; i < cnt; ++i) {
                const yb_block &b = blocks[lo + i];
                metas[i].off = textOff + lay[(size_t)i].off;
                metas[i].nrows = b.nrows; metas[i].size = b.size; metas[i].pitch = lay[(size_t)i].pitch;
                metas[i].firstGap = b.start > 0;
                for (int k = 0; k < lay[(size_t)i].units; ++k) { un[u].block = (int)i; un[u].col0 = k * SCORE_UNIT_COLS; ++u; }
            }
        }
        d.pool->run((int64_t)rowBlock.size(), 16, [&](int64_t a, int64_t z) {
            for (int64_t r = a; r < z; ++r) {
                const uint32_t bi = rowBlock[(size_t)r];
                const yb_block &b = blocks[lo + bi];
                const ScoreLayout &L = lay[bi];
                const uint8_t *src = b.rows[r - rowFirst[bi]];
                unsigned char *dst = h + textOff + L.off + (size_t)(r - rowFirst[bi]) * (size_t)L.pitch;
                dst[0] = dst[1] = dst[2] = 0;
                dst[3] = b.start > 0 ? src[b.start - 1] : 0;            // the column before the range (mz_scores.c:143-147)
                memcpy(dst + 4, src + b.start, (size_t)b.size);
                memset(dst + 4 + b.size, 0, (size_t)L.pitch - 4 - (size_t)b.size);
            }
        });
        d.pack_ms += now_ms() - tp;
        // ---- device ---------------------------------------------------------------------------------------------
        unsigned char *dIn = static_cast<unsigned char *>(g.dIn.p);
        CUDA_TRY(d, cudaEventRecord(g.ev[0], g.stream));
        CUDA_TRY(d, cudaMemcpyAsync(dIn, h, total - 256, cudaMemcpyHostToDevice, g.stream));
        CUDA_TRY(d, cudaMemsetAsync(g.dSums.p, 0, (size_t)cnt * 8, g.stream));
        CUDA_TRY(d, cudaEventRecord(g.ev[1], g.stream));
        if (units > 0) {
            const int warps = SCORE_THREADS / 32;
            yb_score_kernel<<<(unsigned)((units + warps - 1) / warps), SCORE_THREADS, 0, g.stream>>>(
                reinterpret_cast<const ScoreMeta *>(dIn + metaOff), reinterpret_cast<const ScoreUnit *>(dIn + unitOff), (int)units,
                dIn, static_cast<unsigned long long *>(g.dSums.p), rows32, d.sc);
            CUDA_TRY(d, cudaGetLastError());
            ++d.launches;
        }
        CUDA_TRY(d, cudaEventRecord(g.ev[2], g.stream));
        CUDA_TRY(d, cudaMemcpyAsync(g.hSums.p, g.dSums.p, (size_t)cnt * 8, cudaMemcpyDeviceToHost, g.stream));
        CUDA_TRY(d, cudaEventRecord(g.ev[3], g.stream));
        g.first = lo; g.count = cnt; g.busy = true;
        d.h2d_bytes += (int64_t)(total - 256);
        d.d2h_bytes += cnt * 8;
        ++d.waves;
        lo = hi;
    }
    for (int k = 0; k < 2; ++k) {                      // drain, oldest first
        ScoreStage &g = d.score[which ^ k];
        int r2 = score_finish(d, g, scores);
        if (rc == YB_OK) rc = r2;
    }
    if (rc != YB_OK) {
        for (auto &g : d.score) if (g.busy) { cudaStreamSynchronize(g.stream); g.busy = false; }
        if (!d.err.empty()) ctx->err = d.err;
        return rc;
    }
    collect_stats(ctx, stats, now_ms() - t0, pairCols, n);
    return YB_OK;
}

}  // extern "C"
