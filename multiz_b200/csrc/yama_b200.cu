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
#include "plan_kernels.cuh"
#include "score_kernels.cuh"

#include <emmintrin.h>
#include <pthread.h>
#include <sched.h>
#include <cctype>

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
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

extern "C" int yb_band_pack(int M, const int32_t *X, uint8_t *out);      // band_scan.cpp: one byte per row, 255 = listed separately
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
    // cpus: the CPUs of the device's NUMA node (may be empty): the helpers stay next to the device's PCIe root, so that the pinned
    // staging they fill and the copies out of it do not cross the socket interconnect
    explicit Pool(int helpers, const cpu_set_t *cpus = nullptr) {
        for (int t = 0; t < helpers; ++t) th_.emplace_back([this] { loop(); });
        if (cpus && CPU_COUNT(cpus) > 0)
            for (auto &t : th_) pthread_setaffinity_np(t.native_handle(), sizeof(cpu_set_t), cpus);
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
// Bins 3 and 4 take the wide bands, a CTA of 8 warps per pair: band rows of up to 992 cells on a 1 024-entry ring (40 KB of
// shared memory per CTA: 4 CTAs per SM -- the 2 048-entry ring of round 1 allowed 3), up to 4 064 on a 4 096-entry ring.
// -DYB_FILL3 (make ../libyama_b200_fill3.so) puts the experimental fill_body3 there instead: 4 decoupled warps per pair, rings
// that also cover the distance between the lane that writes the ring and the lane that reads it (f3_ring_need).
struct BinCfg { int ring, G, P, minRows, skew; };
#ifdef YB_FILL3
const BinCfg kBin[NBINS] = {{128, 1, 8, 0, 0}, {512, 1, 8, 0, 0}, {512, 4, 1, 192, 0}, {1024, 4, 1, 0, F3_SKEW}, {8192, 4, 1, 0, F3_SKEW},
#else
const BinCfg kBin[NBINS] = {{128, 1, 8, 0, 0}, {512, 1, 8, 0, 0}, {512, 4, 1, 192, 0}, {1024, 8, 1, 0, 0}, {4096, 8, 1, 0, 0},
#endif
                            {128, 1, F2_WARPS, 0, 0}, {512, 1, F2_WARPS, 0, 0}, {128, 1, F2_WARPS, 0, 0}, {512, 1, F2_WARPS, 0, 0}};
inline int ring_need(int bin, int wmax) { return kBin[bin].skew ? f3_ring_need(wmax, kBin[bin].G) : wmax + 32; }
constexpr int BULK_BIN0 = 5;
constexpr int NSLOTS = 8;                 // waves in flight per device (each on its own stream, queued in one go)

static_assert(NBINS == PLAN_NBINS && NB == PLAN_NB && BULK_BIN0 == PLAN_BULK_BIN0, "plan_kernels.cuh mirrors the bin table");

// One staging slot = one wave in flight.
struct Slot {
    cudaStream_t stream = nullptr;
    cudaStream_t binStream[NBINS] = {};    // fill kernels of the wider ring bins run beside the main one
    cudaEvent_t binDone[NBINS] = {};
    cudaEvent_t ev[10] = {};                // 0 copy start, 1 copy end, 7 K0 end | 2 K1 start, 3 K1 end, 8 fill end | 6 K3 start, 4 K3 end, 5 results on the host
    DevBuf dIn, dRow, dCol, dTb, dScript, dOut, dQueue;
    PinBuf hIn, hOut, hPlan;               // hIn: pair descriptors + staged input streams; hPlan: the wave summary of K0
    uint8_t *scriptDst = nullptr;          // where this wave's packed scripts go in the context's pinned store
    bool busy = false;                     // the wave is queued; results not yet unpacked
    bool fresh = false;                    // prepared (copies + K0) since the last slot_wait: their times are still to be counted
    double tPack0 = 0, tPack1 = 0, tTbLaunch = 0;   // host times (ms) of the wave: prepare start/end, traceback launch
    // the wave it holds
    int64_t first = 0, count = 0;
    struct Off { size_t a, b, lb, rb, row, col, sched; uint32_t script; size_t pk; };   // pk: the pair's delta-coded band section
    std::vector<Off> off;                  // per pair, dimension-only offsets (streams, pools)
    // layout of dIn: [descriptors | stream A | stream B | stream LB | stream RB] copied from the host, then
    // [traceback offsets | launch order | long-path list | bucket of each pair | schedules | counters | summary] written by K0
    size_t metaBytes = 0, tbBaseOff = 0, orderOff = 0, longOff = 0, bucketOfOff = 0, schedOff = 0, countersOff = 0, summaryOff = 0;
    size_t h2dBytes = 0, scriptWords = 0, tbEst = 0;
    int maxN = 0, minK = 0, nLongMax = 0;  // dimension facts that bound which kernels the wave can need
    unsigned launchMask = 0;               // kernel bins whose fill is queued for this wave
    int gridHint[NBINS] = {};              // pairs to size a bin's grid for (0: unknown, a full grid)
    PlanSummary sum{};                     // what K0 found (read back before the fill is queued)
    int nLong = 0;                         // pairs whose traceback path gets a warp of its own
    int tbLong = TB_LONG;                  // ... those with at least this many moves
    bool y16 = false;                      // every pair has K*gap_open <= 32767: the 16-bit weight forms of fill_body
    int nValid = 0;
    int binStart[NBINS + 1] = {};
    bool binOnSide[NBINS] = {};            // the bin's fill kernel was queued on its side stream
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
    int onlyBins = 0;
    double tbRatio = 1.0;                  // traceback bytes recent waves needed / the one-warp-per-pair estimate from their dimensions
    unsigned recentBins[4] = {0, 0, 0, 0}; // bins that held pairs in the last waves whose summaries came back (ring)
    int recentCount[4][NBINS] = {};        // ... and how many
    int recentAt = 0, summaries = 0;
    int fillSplit = 2;                     // a wave's fill takes 1/fillSplit of the machine, so that consecutive waves' fills overlap
    int maxCls = 2, keyedMaxK = 0;         // kernel classes on offer; deepest first profile the KEYED class takes (yb_set_scores)
    int fillBlocks[NBINS] = {};
    unsigned fillReady[NBINS] = {};        // fill kernels of the bin whose attributes are set (bit 0: Y16 = false / bulk, bit 1: Y16 = true)
    int prioLo = 0;
    double tContext = 0, tStreams = 0;     // start-up times (YB_PROFILE)
    int helpers = 1;
    cpu_set_t cpus;                        // CPUs of the device's NUMA node that this process may use (empty: unknown / YB_NUMA=0)
    int numaNode = -1;
    std::unique_ptr<Pool> pool;
    // accumulated stats of the current call
    double kernel_ms = 0, fill_ms = 0, profile_ms = 0, tb_ms = 0, h2d_ms = 0, d2h_ms = 0, pack_ms = 0, unpack_ms = 0;
    int64_t h2d_bytes = 0, d2h_bytes = 0, cells = 0, failed = 0;
    int launches = 0, waves = 0;
    std::string err;
    int64_t errJob = 0;
    bool hasResident = false;
    double t_layout = 0, t_par = 0, t_post = 0, t_reserve = 0, t_wait = 0, t_launch = 0;   // YB_PROFILE breakdown
    double plan_ms = 0;                    // device time of K0 (plan, scan, scatter)
    int64_t staged_bytes = 0;              // input bytes that went through a host copy (callers outside yb_host_alloc memory)
    int64_t firstFailed = -1;              // lowest failing job of the batch on this device
    std::vector<int64_t> deferred;         // jobs whose traceback matrix did not fit their wave's pool: run again at the end
};

// The pinned blocks handed out by yb_host_alloc: a caller that builds its jobs inside them is copied from directly.
struct HostBlock { const unsigned char *base; size_t bytes; };

}  // namespace

struct yb_ctx {
    std::vector<Device> devs;
    std::string err;
    bool scoresSet = false;
    ScoreConst sc{};
    int maxDepth = 255;
    int maxAbsS = 1;                        // max |S6|
    bool ungatedOk = true;                  // YB_UNGATED=0 keeps the existence multipliers in every fill kernel
    int slackBulk = 3 + F2_SW - 1;          // schedule slack of the shuffle kernels (see YB_F2_SW in yama_kernels.cuh); YB_SLACK (development)
    int maxCls = 2;                         // highest kernel class handed out: YB_FILL2=0 -> 0 (fill_body only), YB_KEYED=0 -> 1
    int nThreads = 1;
    size_t waveInBytes = (size_t)64 << 20;  // input bytes per wave (steady state)
    size_t waveMinBytes = (size_t)16 << 20; // first waves of a batch (the device idles while the first wave is packed)
    size_t waveTailBytes = (size_t)16 << 20; // last waves of a batch (a short last wave shortens the traceback + unpack tail)
    size_t batchBlobBytes = 0;              // input bytes of the current batch (dimension-only estimate)
    bool waveEnv = false;                   // wave sizes were given in the environment: no adaptation to the batch
    size_t waveTbBytes = (size_t)12 << 30;  // traceback bytes per wave (device memory per slot)
    int64_t wavePairs = 1 << 20;
    int tbLong = TB_LONG;                   // paths of at least this many moves: warp-per-path traceback (YB_TB_LONG)
    bool directCopy = true;                 // inputs inside yb_host_alloc memory are copied from where they are (YB_DIRECT=0: always staged)
    int bandPack = -1;                      // delta-coded bands over PCIe: -1 when the device has >= 8 host threads, YB_BAND_PACK=0/1
    std::vector<HostBlock> hostBlocks;      // yb_host_alloc
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

// one-off variant for work outside a device's pipeline
template <class F>
void parallel_for(int threads, int64_t n, int64_t chunk, F &&fn) {
    Pool pool(std::max(0, threads - 1));
    pool.run(n, chunk, fn);
}

int bin_of(int wmax, int M) {
    if (ring_need(0, wmax) <= kBin[0].ring) return 0;
    if (ring_need(1, wmax) <= kBin[1].ring) return M >= kBin[2].minRows ? 2 : 1;
    if (ring_need(3, wmax) <= kBin[3].ring) return 3;
    if (ring_need(4, wmax) <= kBin[4].ring) return 4;
    return -1;
}
int max_band_row() {                      // widest band row any kernel bin takes
    int w = kBin[4].ring;
    while (ring_need(4, w) > kBin[4].ring) --w;
    return w;
}
int lanes_of(int wmax, int M) {          // wavefront width of the pair's bin (0: no kernel takes it)
    const int b = bin_of(wmax, M);
    return b < 0 ? 0 : 32 * kBin[b].G;
}

size_t fill_smem(int bin) {
    const BinCfg &c = kBin[bin];
    // fill_body3: ring + one FIFO per warp + a 64-B row slot per lane + progress words + the queue slot
    if (c.skew) return (size_t)c.ring * 16 + (size_t)c.G * F3_FIFO * 16 + (size_t)c.G * 32 * 64 + (size_t)c.G * 4 + 32;
    // rings (RING*16-aligned, hence the slack) + 32 B of mailbox per lane + the queue slot
    return (size_t)c.P * ((size_t)c.ring * 16 + (size_t)c.G * 1024) + (size_t)c.ring * 16 + 16;
}

}  // namespace

// ---- kernels with a runtime warps-per-CTA: thin wrappers around the template ---------------------
namespace yb {
#ifndef YB_F1_MINCTAS
#define YB_F1_MINCTAS 4
#endif
template <int RING, int G, int P, bool Y16, bool GATED = true>
__global__ void __launch_bounds__(G * P * 32, (RING == 1024 && G == 8) ? YB_F1_MINCTAS : 1)
yb_fill_kernel_w(const PairMeta *__restrict__ metas, const int *__restrict__ order, const int *__restrict__ binRange,
                 int *__restrict__ queue, const RowRec *__restrict__ rowPool,
                 const ColRec *__restrict__ colPool, unsigned char *__restrict__ tbPool,
                 const unsigned long long *__restrict__ tbBase, PairOut *__restrict__ outs, int gapOpen, int gapExt) {
    fill_body<RING, G, P, Y16, GATED>(metas, order, binRange, queue, rowPool, colPool, tbPool, tbBase, outs, gapOpen, gapExt);
}
// wide bands (fill_body3): a CTA of G decoupled warps per pair
template <int RING, int G, bool Y16>
__global__ void __launch_bounds__(G * 32, (RING <= 1024 ? 7 : 1))
yb_fill3_kernel(const PairMeta *__restrict__ metas, const int *__restrict__ order, const int *__restrict__ binRange,
                int *__restrict__ queue, const RowRec *__restrict__ rowPool,
                const ColRec *__restrict__ colPool, unsigned char *__restrict__ tbPool,
                const unsigned long long *__restrict__ tbBase, PairOut *__restrict__ outs, int gapOpen, int gapExt) {
    fill_body3<RING, G, Y16>(metas, order, binRange, queue, rowPool, colPool, tbPool, tbBase, outs, gapOpen, gapExt);
}
// bulk form (fill_body2): one warp per pair, pairs of kernel class 1 / 2
#ifndef YB_F2_MINCTAS
#define YB_F2_MINCTAS 4
#endif
template <int RING, bool KEYED>
__global__ void __launch_bounds__(F2_WARPS * 32, (RING <= 128 ? YB_F2_MINCTAS : 1))
yb_fill2_kernel(const PairMeta *__restrict__ metas, const int *__restrict__ order, const int *__restrict__ binRange,
                int *__restrict__ queue, const RowRec *__restrict__ rowPool,
                const ColRec *__restrict__ colPool, unsigned char *__restrict__ tbPool,
                const unsigned long long *__restrict__ tbBase, PairOut *__restrict__ outs, int nGO, int gapExt) {
    fill_body2<RING, KEYED>(metas, order, binRange, queue, rowPool, colPool, tbPool, tbBase, outs, nGO, gapExt);
}
}  // namespace yb

namespace {

typedef void (*FillFn)(const PairMeta *, const int *, const int *, int *, const RowRec *, const ColRec *,
                       unsigned char *, const unsigned long long *, PairOut *, int, int);
// (the form without existence multipliers is fill_body2, the bulk bins: fill_fn2)
FillFn fill_fn(int bin, bool y16, bool = false) {
    switch (bin) {
        case 0: return y16 ? yb_fill_kernel_w<128, 1, 8, true> : yb_fill_kernel_w<128, 1, 8, false>;
        case 1: return y16 ? yb_fill_kernel_w<512, 1, 8, true> : yb_fill_kernel_w<512, 1, 8, false>;
        case 2: return y16 ? yb_fill_kernel_w<512, 4, 1, true> : yb_fill_kernel_w<512, 4, 1, false>;
#ifdef YB_FILL3
        case 3: return y16 ? yb_fill3_kernel<1024, 4, true> : yb_fill3_kernel<1024, 4, false>;
        default: return y16 ? yb_fill3_kernel<8192, 4, true> : yb_fill3_kernel<8192, 4, false>;
#else
        case 3: return y16 ? yb_fill_kernel_w<1024, 8, 1, true> : yb_fill_kernel_w<1024, 8, 1, false>;
        default: return y16 ? yb_fill_kernel_w<4096, 8, 1, true> : yb_fill_kernel_w<4096, 8, 1, false>;
#endif
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

// The CPUs of the NUMA node the device hangs off (sysfs), intersected with what the process may use.  YB_NUMA=0 turns it off.
void device_cpus(Device &d) {
    CPU_ZERO(&d.cpus);
    d.numaNode = -1;
    if (const char *e = getenv("YB_NUMA")) if (atoi(e) == 0) return;
    char bdf[32] = {0};
    if (cudaDeviceGetPCIBusId(bdf, sizeof bdf, d.id) != cudaSuccess) return;
    for (char *c = bdf; *c; ++c) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bdf);
    FILE *f = fopen(path, "r");
    int node = -1;
    if (f) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
    if (node < 0) return;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return;
    cpu_set_t mine, want;
    CPU_ZERO(&want);
    int a, b;
    char sep;
    while (fscanf(f, "%d", &a) == 1) {                     // "0-15,32-47"
        b = a;
        if (fscanf(f, "%c", &sep) == 1 && sep == '-') { if (fscanf(f, "%d", &b) != 1) b = a; if (fscanf(f, "%c", &sep) != 1) sep = 0; }
        for (int c = a; c <= b && c < CPU_SETSIZE; ++c) CPU_SET(c, &want);
        if (sep != ',') break;
    }
    fclose(f);
    if (sched_getaffinity(0, sizeof mine, &mine) != 0) return;
    CPU_AND(&d.cpus, &mine, &want);
    // (never trade threads for placement: the node must offer a CPU per thread of this device)
    if (CPU_COUNT(&d.cpus) < d.helpers) { CPU_ZERO(&d.cpus); return; }
    d.numaNode = node;
}

int device_init(Device &d) {
    const double t0 = now_ms();
    CUDA_TRY(d, cudaSetDevice(d.id));
    CUDA_TRY(d, cudaFree(nullptr));                          // (the context is created here)
    cudaDeviceProp prop;
    CUDA_TRY(d, cudaGetDeviceProperties(&prop, d.id));
    d.sms = prop.multiProcessorCount;
    d.tContext = now_ms() - t0;
    // A wave's copies, K0, K1, K3 run on its main stream, the fill kernels on the bin streams.  The fill kernels are persistent
    // and fill the machine; the short kernels of the NEXT waves must get the CTA slots that free up first, or every wave's
    // K0 / K1 / K3 waits behind a whole fill: the main streams get the higher priority (YB_PRIO=0: all equal).
    int prioLo = 0, prioHi = 0;
    CUDA_TRY(d, cudaDeviceGetStreamPriorityRange(&prioLo, &prioHi));
    if (const char *e = getenv("YB_PRIO")) if (atoi(e) == 0) prioHi = prioLo;
    d.prioLo = prioLo;
    for (auto &s : d.slots) {
        CUDA_TRY(d, cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, prioHi));
        for (auto &e : s.ev) CUDA_TRY(d, cudaEventCreate(&e));
    }
    d.tStreams = now_ms() - t0 - d.tContext;
    // (the bins' streams, and the fill kernels' attributes -- which make the driver load the kernel -- are set up at a bin's
    //  first launch: a tool invocation uses two or three of the twenty fill kernels, and its start-up is its wall clock)
    return YB_OK;
}

// First launch of a fill kernel on this device: dynamic shared memory limit, grid size from the occupancy.
int ensure_fill(Device &d, int b, bool y16) {
    const bool bulk = b >= BULK_BIN0;
    const unsigned bit = (!bulk && y16) ? 2u : 1u;
    if (d.fillReady[b] & bit) return YB_OK;
    FillFn fn = bulk ? fill_fn2(b) : fill_fn(b, y16, false);
    const size_t sm = bulk ? fill2_smem(b) : fill_smem(b);
    const int threads = bulk ? F2_WARPS * 32 : kBin[b].G * kBin[b].P * 32;
    CUDA_TRY(d, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    int occ = 0;
    CUDA_TRY(d, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, sm));
    d.fillBlocks[b] = std::max(occ, 1) * d.sms;
    d.fillReady[b] |= bit;
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

inline bool dims_ok(const yb_job &j) { return j.K >= 1 && j.L >= 1 && j.M >= 1 && j.N >= 1 && j.A && j.B && j.LB && j.RB; }
inline size_t sched_ints(const yb_job &j) { return (size_t)((j.M + 31) >> 5) + 1; }   // upper bound (B >= 32)

// input bytes of a job: known from its dimensions alone (wave planning needs no band read)
inline size_t blob_bytes(const yb_job &j) {
    return (size_t)j.K * j.M + (size_t)j.L * j.N + (size_t)(j.M + 1) * 8;
}

inline bool in_block(const std::vector<HostBlock> &blocks, const unsigned char *lo, const unsigned char *hi) {
    for (const HostBlock &b : blocks)
        if (lo >= b.base && hi <= b.base + b.bytes) return true;
    return false;
}

// Lay out jobs [first, first+count) from their DIMENSIONS, queue the host->device copies of their four input streams
// (A, B, LB, RB: straight from the caller's buffers when those lie in yb_host_alloc memory, through the slot's pinned
// staging buffer otherwise -- a memcpy either way, no band row is read on the host) and K0 behind them.  Nothing here waits
// for the device.
int slot_prepare(yb_ctx *ctx, Device &d, Slot &s, const yb_job *jobs, int64_t first, int64_t count, size_t tbCapacity, bool allBins = false,
                 bool rawBands = false) {
    const double t0 = now_ms();
    // delta-coded bands (yb_band_expand): worth it when the host has threads to spare -- it reads every band row once so
    // that PCIe carries a quarter of it; with few threads per device the copy of the raw rows is the faster way (measured
    // on cfg2, end to end: 16 threads 12.7 -> 12.2 ms, 8 threads 12.7 -> 11.6 ms, 4 threads 15.3 ms with it -- the coding pass
    // then paces the waves; cfg3 with 16 threads +21 %)
    const bool pack = !rawBands && count > 0 && (ctx->bandPack > 0 || (ctx->bandPack < 0 && d.helpers >= 8));
    s.tPack0 = t0;
    s.first = first;
    s.off.resize((size_t)count + 1);
    // ---- dimension-only layout + the address range of each stream: per chunk of jobs on the helper threads (offsets
    // relative to the chunk), then one serial pass over the chunks ------------------------------------------------------------
    struct Stream { const unsigned char *lo = nullptr, *hi = nullptr; size_t payload = 0; bool direct = false; size_t devOff = 0, bytes = 0; };
    Stream st[4];
    size_t rows = 0, cols = 0, words = 0, sched = 0, pkBytes = 0;
    int maxK = 0, maxN = 0, minK = 0x7fffffff, nLongMax = 0;
    size_t tbEst = 0;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(256, count / (4 * (int64_t)d.helpers)));
    const int64_t nChunks = (count + chunk - 1) / chunk;
    struct Part {
        Slot::Off tot{};
        const unsigned char *lo[4] = {nullptr, nullptr, nullptr, nullptr}, *hi[4] = {nullptr, nullptr, nullptr, nullptr};
        int maxK = 0, maxN = 0, minK = 0x7fffffff, nLong = 0;
        size_t tbEst = 0;
    };
    std::vector<Part> part((size_t)nChunks);
    const int tbLongMoves = ctx->tbLong;
    d.pool->run(count, chunk, [&](int64_t lo0, int64_t hi0) {
        for (int64_t c = lo0 / chunk; c * chunk < hi0; ++c) {                 // (a run without helpers gets [0, count) at once)
            Part P;
            Slot::Off r{};
            for (int64_t i = c * chunk, e = std::min<int64_t>(count, (c + 1) * chunk); i < e; ++i) {
                const yb_job &j = jobs[first + i];
                s.off[(size_t)i] = r;
                if (!dims_ok(j)) continue;
                r.pk += align_up(sizeof(BandPackHdr) + 2 * align_up((size_t)j.M + 1, 4), 16);
                P.maxK = std::max(P.maxK, j.K); P.minK = std::min(P.minK, j.K); P.maxN = std::max(P.maxN, j.N);
                P.tbEst += 32ull * ((size_t)j.M + j.N + 3 * (((size_t)j.M + 31) >> 5) + 40);      // one warp, a diagonal band
                if (j.M + j.N >= tbLongMoves) ++P.nLong;
                const unsigned char *ptr[4] = {j.A, j.B, reinterpret_cast<const unsigned char *>(j.LB), reinterpret_cast<const unsigned char *>(j.RB)};
                const size_t len[4] = {(size_t)j.K * j.M, (size_t)j.L * j.N, (size_t)(j.M + 1) * 4, (size_t)(j.M + 1) * 4};
                for (int k = 0; k < 4; ++k) {
                    if (!P.lo[k] || ptr[k] < P.lo[k]) P.lo[k] = ptr[k];
                    if (!P.hi[k] || ptr[k] + len[k] > P.hi[k]) P.hi[k] = ptr[k] + len[k];
                }
                r.a += align_up(len[0], 64); r.b += align_up(len[1], 64);       // (staged layout: sections of whole 64-byte lines)
                r.lb += align_up(len[2], 64); r.rb += align_up(len[3], 64);
                r.row += (size_t)j.M + 1; r.col += (size_t)j.N + 1; r.sched += sched_ints(j);
                r.script += (uint32_t)(((size_t)j.M + j.N + 15) / 16);
            }
            P.tot = r;
            part[(size_t)c] = P;
        }
    });
    {
        Slot::Off run{};
        for (int64_t c = 0; c < nChunks; ++c) {
            Part &P = part[(size_t)c];
            const Slot::Off t = P.tot;
            P.tot = run;                                                      // now: the chunk's base offsets
            run.a += t.a; run.b += t.b; run.lb += t.lb; run.rb += t.rb; run.row += t.row; run.col += t.col; run.sched += t.sched;
            run.script += t.script; run.pk += t.pk;
            for (int k = 0; k < 4; ++k) {
                if (P.lo[k] && (!st[k].lo || P.lo[k] < st[k].lo)) st[k].lo = P.lo[k];
                if (P.hi[k] && (!st[k].hi || P.hi[k] > st[k].hi)) st[k].hi = P.hi[k];
            }
            maxK = std::max(maxK, P.maxK); minK = std::min(minK, P.minK); maxN = std::max(maxN, P.maxN);
            nLongMax += P.nLong; tbEst += P.tbEst;
        }
        st[0].payload = run.a; st[1].payload = run.b; st[2].payload = run.lb; st[3].payload = run.rb;
        rows = run.row; cols = run.col; sched = run.sched; words = run.script; pkBytes = run.pk;
    }
    s.off[(size_t)count] = Slot::Off{st[0].payload, st[1].payload, st[2].payload, st[3].payload, rows, cols, sched, (uint32_t)words, pkBytes};
    s.y16 = (int64_t)std::min(maxK, ctx->maxDepth) * ctx->sc.gap_open <= 32767;
    s.count = count;
    s.scriptWords = words;
    // a stream is copied straight from the caller's memory when that is pinned (yb_host_alloc) and dense enough
    s.metaBytes = align_up((size_t)count * sizeof(PairMeta), 256);
    size_t devOff = s.metaBytes, stagedBytes = 0;
    size_t stageOff[4] = {0, 0, 0, 0};
    for (int k = 0; k < 4; ++k) {
        Stream &x = st[k];
        if (!x.lo) { x.devOff = devOff; x.bytes = 0; continue; }
        const unsigned char *lo = reinterpret_cast<const unsigned char *>(reinterpret_cast<uintptr_t>(x.lo) & ~(uintptr_t)63);
        const size_t span = (size_t)(x.hi - lo);
        if (pack && k >= 2) {                                      // restored on the device, in the staged layout
            x.direct = false; x.bytes = 0; x.devOff = devOff;
            devOff += align_up(x.payload, 256) + 256;
            continue;
        }
        x.direct = ctx->directCopy && span <= x.payload + x.payload / 2 + 65536 && in_block(ctx->hostBlocks, lo, x.hi);
        if (x.direct) { x.lo = lo; x.bytes = span; }
        else { x.bytes = x.payload; stageOff[k] = s.metaBytes + stagedBytes; stagedBytes += align_up(x.payload, 256); }
        x.devOff = devOff;
        devOff += align_up(x.bytes, 256) + 256;                    // (slack: K1 reads whole words around a row)
    }
    // delta-coded bands: [section offset per pair | sections | exception list], contiguous on the host and on the device
    const size_t excCap = pack ? (size_t)count + 4096 : 0;
    const size_t pkArr = align_up((size_t)count * 8, 256), pkSec = align_up(pkBytes, 256);
    const size_t pkHost = s.metaBytes + stagedBytes, pkDev = devOff;
    if (pack) { stagedBytes += pkArr + pkSec + align_up(excCap * sizeof(BandExc), 256); devOff += pkArr + pkSec + align_up(excCap * sizeof(BandExc), 256); }
    std::atomic<int64_t> excFill{0};
    std::atomic<bool> excOverflow{false};
    const size_t copied = devOff;
    s.tbBaseOff = devOff; devOff += align_up((size_t)count * 8, 256);
    s.orderOff = devOff; devOff += align_up((size_t)count * 4, 256);
    s.longOff = devOff; devOff += align_up((size_t)count * 4, 256);
    s.bucketOfOff = devOff; devOff += align_up((size_t)count * 4, 256);
    s.schedOff = devOff; devOff += align_up(sched * 4, 256);
    s.countersOff = devOff; devOff += align_up((size_t)(2 * NBINS * NB + 16) * 4, 256);
    s.summaryOff = devOff; devOff += 256;
    CUDA_TRY(d, s.hIn.reserve(s.metaBytes + stagedBytes + 256));
    {
        const void *before = s.dIn.p;
        CUDA_TRY(d, s.dIn.reserve(devOff));
        if (s.dIn.p != before) CUDA_TRY(d, cudaMemsetAsync(s.dIn.p, 0, s.dIn.cap, s.stream));   // (slack bytes are read, never used)
    }
    CUDA_TRY(d, s.dOut.reserve((size_t)count * sizeof(PairOut) + 64));
    CUDA_TRY(d, s.hOut.reserve((size_t)count * sizeof(PairOut) + 64));
    CUDA_TRY(d, s.hPlan.reserve(256));
    CUDA_TRY(d, s.dQueue.reserve(64));
    unsigned char *h = static_cast<unsigned char *>(s.hIn.p);
    PairMeta *metas = reinterpret_cast<PairMeta *>(h);
    const double t1 = now_ms();
    d.t_layout += t1 - t0;

    // ---- pair descriptors (+ staged copies) on the helper threads -------------------------------------------------------
    d.pool->run(count, chunk, [&](int64_t lo0, int64_t hi0) {
        for (int64_t i = lo0; i < hi0; ++i) {
            const yb_job &j = jobs[first + i];
            Slot::Off o = s.off[(size_t)i];
            {
                const Slot::Off &b = part[(size_t)(i / chunk)].tot;
                o.a += b.a; o.b += b.b; o.lb += b.lb; o.rb += b.rb; o.row += b.row; o.col += b.col; o.sched += b.sched;
                o.script += b.script; o.pk += b.pk;
            }
            PairMeta pm;
            memset(&pm, 0, sizeof pm);
            if (dims_ok(j)) {
                pm.K = j.K; pm.M = j.M; pm.L = j.L; pm.N = j.N;
                const unsigned char *ptr[4] = {j.A, j.B, reinterpret_cast<const unsigned char *>(j.LB), reinterpret_cast<const unsigned char *>(j.RB)};
                const size_t len[4] = {(size_t)j.K * j.M, (size_t)j.L * j.N, (size_t)(j.M + 1) * 4, (size_t)(j.M + 1) * 4};
                const size_t so[4] = {o.a, o.b, o.lb, o.rb};
                unsigned long long dev[4];
                for (int k = 0; k < 4; ++k) {
                    if (st[k].direct) dev[k] = st[k].devOff + (size_t)(ptr[k] - st[k].lo);
                    else {
                        dev[k] = st[k].devOff + so[k];
                        if (!(pack && k >= 2)) copy_nt64(h + stageOff[k] + so[k], ptr[k], len[k]);
                    }
                }
                if (pack) {
                    unsigned char *sec = h + pkHost + pkArr + o.pk;
                    const size_t a4 = align_up((size_t)j.M + 1, 4);
                    BandPackHdr *hd = reinterpret_cast<BandPackHdr *>(sec);
                    uint8_t *dl = sec + sizeof(BandPackHdr), *dr = dl + a4;
                    memset(dl + a4 - 4, 0, 4); memset(dr + a4 - 4, 0, 4);          // pad bytes past row M
                    const int n1 = yb_band_pack(j.M, j.LB, dl), n2 = yb_band_pack(j.M, j.RB, dr);
                    hd->LB0 = j.LB[0]; hd->RB0 = j.RB[0]; hd->nExc = n1 + n2; hd->excStart = 0;
                    if (n1 + n2 > 0) {
                        const int64_t at = excFill.fetch_add(n1 + n2);
                        if ((size_t)(at + n1 + n2) > excCap) { excOverflow = true; hd->nExc = 0; }
                        else {
                            hd->excStart = (int)at;
                            BandExc *ex = reinterpret_cast<BandExc *>(h + pkHost + pkArr + pkSec) + at;
                            for (int r = 1; r <= j.M; ++r)
                                if (dl[r] == 255) *ex++ = BandExc{r, (int)((uint32_t)j.LB[r] - (uint32_t)j.LB[r - 1])};
                            for (int r = 1; r <= j.M; ++r)
                                if (dr[r] == 255) *ex++ = BandExc{j.M + 1 + r, (int)((uint32_t)j.RB[r] - (uint32_t)j.RB[r - 1])};
                        }
                    }
                    reinterpret_cast<unsigned long long *>(h + pkHost)[i] = pkDev + pkArr + o.pk;
                }
                pm.offA = dev[0]; pm.offB = dev[1]; pm.offBand = dev[2]; pm.offBand2 = dev[3];
                pm.offSched = s.schedOff + o.sched * 4;
                pm.rowBase = o.row; pm.colBase = o.col; pm.scriptBase = o.script;
                pm.lgLanes = 5;
            }
            else if (pack) reinterpret_cast<unsigned long long *>(h + pkHost)[i] = BAND_RAW;
            metas[i] = pm;
        }
        _mm_sfence();
    });
    // more band steps outside 0..254 than the exception list holds (no band pre_yama builds does that): the plain way
    if (excOverflow) return slot_prepare(ctx, d, s, jobs, first, count, tbCapacity, allBins, true);
    const double t2 = now_ms();
    d.t_par += t2 - t1;

    // ---- copies + K0 ------------------------------------------------------------------------------------------------------
    unsigned char *dIn = static_cast<unsigned char *>(s.dIn.p);
    cudaStream_t q = s.stream;
    CUDA_TRY(d, cudaEventRecord(s.ev[0], q));
    CUDA_TRY(d, cudaMemcpyAsync(dIn, h, (size_t)count * sizeof(PairMeta), cudaMemcpyHostToDevice, q));
    s.h2dBytes = (size_t)count * sizeof(PairMeta);
    for (int k = 0; k < 4; ++k) {
        if (!st[k].bytes) continue;
        const void *src = st[k].direct ? static_cast<const void *>(st[k].lo) : static_cast<const void *>(h + stageOff[k]);
        CUDA_TRY(d, cudaMemcpyAsync(dIn + st[k].devOff, src, st[k].bytes, cudaMemcpyHostToDevice, q));
        s.h2dBytes += st[k].bytes;
        if (!st[k].direct) d.staged_bytes += (int64_t)st[k].bytes;
    }
    if (pack) {
        const size_t n = pkArr + pkSec + (size_t)excFill.load() * sizeof(BandExc);
        CUDA_TRY(d, cudaMemcpyAsync(dIn + pkDev, h + pkHost, n, cudaMemcpyHostToDevice, q));
        s.h2dBytes += n;
        d.staged_bytes += (int64_t)n;
    }
    (void)copied;
    CUDA_TRY(d, cudaEventRecord(s.ev[1], q));
    // every other buffer of the wave is sized from the dimensions; the traceback pool has a fixed capacity (pairs that do not
    // fit it any more are deferred by K0 and run again in a later wave)
    CUDA_TRY(d, s.dRow.reserve(rows * sizeof(RowRec) + 64));
    {   // column records sit between two COL_PAD margins (see fill_body2); a fresh buffer is zeroed once so that what idle
        // lanes read there is initialised memory
        const void *before = s.dCol.p;
        CUDA_TRY(d, s.dCol.reserve(cols * sizeof(ColRec) + 2 * COL_PAD + 64));
        if (s.dCol.p != before) CUDA_TRY(d, cudaMemsetAsync(s.dCol.p, 0, s.dCol.cap, q));
    }
    // The traceback pool of the wave: what its pairs would need with one warp each on diagonal bands (known from the
    // dimensions), times what recent waves needed relative to that estimate (wide bands run 4 or 8 warps per pair), with a
    // margin; never more than tbCapacity.  Pairs that do not fit are deferred by K0 and run again at the end -- a fixed
    // large pool per slot would cost every short-lived process seconds of allocation and teardown.
    s.tbEst = tbEst;
    if (!allBins || tbCapacity == 0) tbCapacity = std::min<size_t>(std::max<size_t>(tbCapacity, (size_t)16 << 20),
                                                                    (size_t)((double)tbEst * d.tbRatio * 1.25) + ((size_t)16 << 20));
    CUDA_TRY(d, s.dTb.reserve(std::max(tbCapacity, (size_t)1 << 20) + 256));
    CUDA_TRY(d, s.dScript.reserve(words * 4 + 64));
    s.scriptDst = ctx->scriptStore + ctx->scriptOff[(size_t)first];   // the wave's scripts, in job order, in the batch store
    int *counters = reinterpret_cast<int *>(dIn + s.countersOff);
    CUDA_TRY(d, cudaMemsetAsync(counters, 0, (size_t)(2 * NBINS * NB + 16) * 4, q));
    PlanParams pp;
    for (int b = 0; b < NBINS; ++b) { pp.ring[b] = kBin[b].ring; pp.warps[b] = kBin[b].G; pp.skew[b] = kBin[b].skew; }
    pp.minRows2 = kBin[2].minRows;
    pp.maxDepth = ctx->maxDepth; pp.maxCls = ctx->maxCls; pp.maxAbsS = ctx->maxAbsS;
    pp.gapOpen = ctx->sc.gap_open; pp.gapExt = ctx->sc.gap_ext; pp.tbLong = ctx->tbLong;
    pp.slackBulk = ctx->slackBulk;
    s.tbLong = ctx->tbLong;
    {   // Which kernel bins get a fill launch: those the wave's DIMENSIONS allow (a band row has at most N+1 cells; the KEYED
        // class needs a shallow first profile) and that recent waves actually used -- an empty launch is not free (0.1 ms
        // and more each, measured).  A pair whose bin is left out is deferred by K0 and runs again at the end of the batch.
        const int widest = maxN + 1 + 32;
        unsigned can = 1u;
        if (widest > kBin[0].ring) can |= (1u << 1) | (1u << 2);
        if (widest > kBin[1].ring) can |= 1u << 3;
        if (ring_need(3, maxN + 1) > kBin[3].ring) can |= 1u << 4;
        if (d.maxCls >= 1) can |= (1u << BULK_BIN0) | ((can & 2u) ? 1u << (BULK_BIN0 + 1) : 0u);
        if (d.maxCls >= 2 && minK <= d.keyedMaxK) can |= (1u << (BULK_BIN0 + 2)) | ((can & 2u) ? 1u << (BULK_BIN0 + 3) : 0u);
        unsigned recent = d.recentBins[0] | d.recentBins[1] | d.recentBins[2] | d.recentBins[3];
        if (allBins || d.summaries == 0) recent = ~0u;
        s.launchMask = can & recent;
        if (d.onlyBins) s.launchMask &= (unsigned)d.onlyBins;                  // (development: YB_ONLY_BINS)
        if (!s.launchMask) s.launchMask = can & ~0u;
        pp.launchMask = s.launchMask;
        // grids: a bin that held few pairs lately gets few CTAs (the kernels are persistent: a short grid only means more
        // pairs per CTA should the guess be low) -- a full grid of a bin with three pairs keeps the bulk bin off the SMs
        for (int b = 0; b < NBINS; ++b) {
            int m = 0;
            for (int k = 0; k < 4; ++k) m = std::max(m, d.recentCount[k][b]);
            s.gridHint[b] = (allBins || d.summaries == 0) ? 0 : 2 * m + 8 * kBin[b].P;
        }
    }
    PairOut *outs = static_cast<PairOut *>(s.dOut.p);
    unsigned long long *tbBase = reinterpret_cast<unsigned long long *>(dIn + s.tbBaseOff);
    int *bucketOf = reinterpret_cast<int *>(dIn + s.bucketOfOff);
    int *bucketCount = counters, *bucketFill = counters + NBINS * NB, *longFill = counters + 2 * NBINS * NB;
    PlanSummary *dsum = reinterpret_cast<PlanSummary *>(dIn + s.summaryOff);
    if (count > 0) {
        const int warpsPerCta = PLAN_THREADS / 32;
        const unsigned blocks = (unsigned)std::min<int64_t>((count + warpsPerCta - 1) / warpsPerCta, (int64_t)d.sms * 8);
        if (pack) {
            yb_band_expand<<<blocks, PLAN_THREADS, 0, q>>>(reinterpret_cast<const PairMeta *>(dIn), (int)count, dIn,
                                                           reinterpret_cast<const unsigned long long *>(dIn + pkDev),
                                                           reinterpret_cast<const BandExc *>(dIn + pkDev + pkArr + pkSec));
            d.launches++;
        }
        yb_plan_kernel<<<blocks, PLAN_THREADS, 0, q>>>(reinterpret_cast<PairMeta *>(dIn), (int)count, dIn, outs, tbBase, bucketOf, pp);
        yb_plan_scan<<<1, 1024, 0, q>>>((int)count, tbBase, bucketOf, bucketCount, bucketFill, outs, reinterpret_cast<PairMeta *>(dIn), dsum, pp.tbLong,
                                       (unsigned long long)tbCapacity, s.launchMask);
        yb_plan_scatter<<<(unsigned)((count + 255) / 256), 256, 0, q>>>((int)count, bucketOf, bucketCount, bucketFill, reinterpret_cast<const PairMeta *>(dIn),
                                                                       reinterpret_cast<int *>(dIn + s.orderOff), reinterpret_cast<int *>(dIn + s.longOff), longFill, pp.tbLong);
        d.launches += 3;
    } else {
        CUDA_TRY(d, cudaMemsetAsync(dsum, 0, sizeof(PlanSummary), q));
    }
    CUDA_TRY(d, cudaEventRecord(s.ev[7], q));
    CUDA_TRY(d, cudaGetLastError());
    s.maxN = maxN; s.minK = minK; s.nLongMax = nLongMax;
    d.t_post += now_ms() - t2;
    d.pack_ms += now_ms() - t0;
    d.h2d_bytes += (int64_t)s.h2dBytes;
    s.fresh = true;
    s.tPack1 = now_ms();
    return YB_OK;
}

// Enqueue K1, K2, K3 and the result copies of a prepared wave.  The kernels take their pair ranges from K0's summary in
// device memory, so nothing here waits for the device either: the host never learns the plan before the results arrive.
int slot_launch(Device &d, Slot &s, bool d2h, int fillSplit = 1) {
    const PairMeta *metas = static_cast<const PairMeta *>(s.dIn.p);
    unsigned char *blob = static_cast<unsigned char *>(s.dIn.p);
    RowRec *rows = static_cast<RowRec *>(s.dRow.p);
    ColRec *cols = reinterpret_cast<ColRec *>(static_cast<unsigned char *>(s.dCol.p) + COL_PAD);
    unsigned char *tb = static_cast<unsigned char *>(s.dTb.p);
    unsigned *script = static_cast<unsigned *>(s.dScript.p);
    PairOut *outs = static_cast<PairOut *>(s.dOut.p);
    const int *order = reinterpret_cast<const int *>(blob + s.orderOff);
    const PlanSummary *dsum = reinterpret_cast<const PlanSummary *>(blob + s.summaryOff);
    const unsigned long long *tbBase = reinterpret_cast<const unsigned long long *>(blob + s.tbBaseOff);
    int *queue = static_cast<int *>(s.dQueue.p);
    cudaStream_t st = s.stream;
    const double tl = now_ms();

    CUDA_TRY(d, cudaEventRecord(s.ev[2], st));
    CUDA_TRY(d, cudaMemsetAsync(queue, 0, 64, st));
    // a script region holds ceil((M+N)/16) words but a path has m_new <= M+N ops: the unwritten tail is copied back too
    if (s.scriptWords) CUDA_TRY(d, cudaMemsetAsync(s.dScript.p, 0, s.scriptWords * 4, st));
    if (s.count > 0) {
        yb_profile_kernel<<<(unsigned)s.count, K1_THREADS, 0, st>>>(metas, blob, rows, cols, s.y16 ? 1 : 0, d.sc);
        d.launches++;
    }
    CUDA_TRY(d, cudaEventRecord(s.ev[3], st));
    // One fill launch per kernel bin of the wave's launch mask (slot_prepare).
    // Every bin runs on a stream of its own behind the same event, so that they become runnable together and the hardware
    // starts them in launch order: the wide bins get their CTAs before the bulk bins' persistent CTAs take the machine.
    static const int launchOrder[NBINS] = {4, 3, 2, 1, 6, 8, 0, 5, 7};       // wide rings first, the bulk bins last
    for (int k = 0; k < NBINS && s.count > 0; ++k) {
        const int b = launchOrder[k];
        s.binOnSide[b] = false;
        if (!((s.launchMask >> b) & 1u)) continue;
        const BinCfg &bc = kBin[b];
        if (!s.binStream[b]) {
            CUDA_TRY(d, cudaStreamCreateWithPriority(&s.binStream[b], cudaStreamNonBlocking, d.prioLo));
            CUDA_TRY(d, cudaEventCreateWithFlags(&s.binDone[b], cudaEventDisableTiming));
        }
        if (int rc = ensure_fill(d, b, s.y16)) return rc;
        cudaStream_t bs = s.binStream[b];
        CUDA_TRY(d, cudaStreamWaitEvent(bs, s.ev[3], 0));
        const bool bulk = b >= BULK_BIN0;
        FillFn fn = bulk ? fill_fn2(b) : fill_fn(b, s.y16, false);
        const size_t smem = bulk ? fill2_smem(b) : fill_smem(b);
        const int64_t pairsGuess = s.gridHint[b] > 0 ? std::min<int64_t>(s.count, s.gridHint[b]) : s.count;
        const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((pairsGuess + bc.P - 1) / bc.P, std::max(d.sms, d.fillBlocks[b] / fillSplit)));
        fn<<<blocks, bc.G * bc.P * 32, smem, bs>>>(metas, order, dsum->binStart + b, queue + b, rows, cols, tb, tbBase, outs,
                                                    bulk ? -d.sc.gap_open * (b >= BULK_BIN0 + 2 ? 4 : 1) : d.sc.gap_open, d.sc.gap_ext);
        CUDA_TRY(d, cudaEventRecord(s.binDone[b], bs));
        s.binOnSide[b] = true;
        d.launches++;
    }
    for (int b = 0; b < NBINS; ++b)
        if (s.count > 0 && s.binOnSide[b]) CUDA_TRY(d, cudaStreamWaitEvent(st, s.binDone[b], 0));
    CUDA_TRY(d, cudaEventRecord(s.ev[8], st));

    // K3.  The warp-per-path kernel (issue-bound) runs beside the thread-per-pair kernel (latency-bound), on a side stream.
    s.tTbLaunch = now_ms();
    CUDA_TRY(d, cudaEventRecord(s.ev[6], st));
    if (s.nLongMax > 0) {
        if (!s.binStream[1]) {
            CUDA_TRY(d, cudaStreamCreateWithPriority(&s.binStream[1], cudaStreamNonBlocking, d.prioLo));
            CUDA_TRY(d, cudaEventCreateWithFlags(&s.binDone[1], cudaEventDisableTiming));
        }
        cudaStream_t ls = s.binStream[1];
        CUDA_TRY(d, cudaStreamWaitEvent(ls, s.ev[6], 0));
        yb_traceback_long_kernel<<<(unsigned)((s.nLongMax + 3) / 4), 128, 0, ls>>>(
            metas, reinterpret_cast<const int *>(blob + s.longOff), &dsum->nLong, blob, tb, tbBase, script, outs);
        CUDA_TRY(d, cudaEventRecord(s.binDone[1], ls));
        d.launches++;
    }
    if (s.count > 0) {
        yb_traceback_kernel<<<(unsigned)((s.count + 127) / 128), 128, 0, st>>>(metas, order, &dsum->nValid, blob, tb, tbBase, script, outs, s.tbLong);
        d.launches++;
    }
    if (s.nLongMax > 0) CUDA_TRY(d, cudaStreamWaitEvent(st, s.binDone[1], 0));
    CUDA_TRY(d, cudaEventRecord(s.ev[4], st));
    CUDA_TRY(d, cudaMemcpyAsync(s.hPlan.p, dsum, sizeof(PlanSummary), cudaMemcpyDeviceToHost, st));
    if (d2h) {
        CUDA_TRY(d, cudaMemcpyAsync(s.hOut.p, s.dOut.p, (size_t)s.count * sizeof(PairOut), cudaMemcpyDeviceToHost, st));
        if (s.scriptWords)
            CUDA_TRY(d, cudaMemcpyAsync(s.scriptDst, s.dScript.p, s.scriptWords * 4, cudaMemcpyDeviceToHost, st));
        d.d2h_bytes += (int64_t)((size_t)s.count * sizeof(PairOut) + s.scriptWords * 4);
    }
    CUDA_TRY(d, cudaEventRecord(s.ev[5], st));
    CUDA_TRY(d, cudaGetLastError());
    s.busy = true;
    d.waves++;
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
    float h = 0, k0 = 0, a = 0, b = 0, c = 0, e = 0;
    if (s.fresh) { cudaEventElapsedTime(&h, s.ev[0], s.ev[1]); cudaEventElapsedTime(&k0, s.ev[1], s.ev[7]); s.fresh = false; }
    cudaEventElapsedTime(&a, s.ev[2], s.ev[3]);
    cudaEventElapsedTime(&b, s.ev[3], s.ev[8]);
    cudaEventElapsedTime(&c, s.ev[6], s.ev[4]);
    cudaEventElapsedTime(&e, s.ev[4], s.ev[5]);
    d.h2d_ms += h; d.plan_ms += k0; d.profile_ms += a; d.fill_ms += b; d.tb_ms += c; d.d2h_ms += e;
    d.kernel_ms += k0 + a + b + c;
    s.sum = *static_cast<const PlanSummary *>(s.hPlan.p);           // K0's summary came back with the results
    {
        unsigned used = 0;
        for (int b = 0; b < NBINS; ++b) {
            d.recentCount[d.recentAt][b] = s.sum.binStart[b + 1] - s.sum.binStart[b];
            if (d.recentCount[d.recentAt][b] > 0) used |= 1u << b;
        }
        d.recentBins[d.recentAt] = used;
        if (s.tbEst > 0 && s.sum.tbNeed > 0) {
            const double r = (double)s.sum.tbNeed / (double)s.tbEst;
            d.tbRatio = std::max(1.0, d.summaries == 0 ? r : std::max(r, 0.5 * d.tbRatio + 0.5 * r));
        }
        d.recentAt = (d.recentAt + 1) & 3;
        ++d.summaries;
    }
    if (s.sum.firstFailed >= 0 && (d.firstFailed < 0 || s.first + s.sum.firstFailed < d.firstFailed)) d.firstFailed = s.first + s.sum.firstFailed;
    if (d.timeline && d.evBase) {           // YB_PROFILE=2: where this wave sat on the device's and the host's clocks
        float t[9] = {0};
        const int idx[9] = {0, 1, 7, 2, 3, 8, 6, 4, 5};
        for (int k = 0; k < 9; ++k) cudaEventElapsedTime(&t[k], d.evBase, s.ev[idx[k]]);
        fprintf(stderr, "yama_b200[timeline] dev %d wave %3d pairs %6lld | host: prepare %.2f-%.2f tb-launch %.2f wait-done %.2f | device: h2d %.2f-%.2f "
                "K0 -%.2f | K1 %.2f-%.2f K2 -%.2f | K3 %.2f-%.2f d2h -%.2f\n", d.id, d.waves, (long long)s.count, s.tPack0 - d.tBase, s.tPack1 - d.tBase,
                s.tTbLaunch - d.tBase, now_ms() - d.tBase, t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8]);
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
            const PairOut &o = outs[i];
            memset(&r, 0, sizeof r);
            r.status = o.status;                       // K0: YB_ERR_ARG / _BAND / _LIMIT (or deferred); K3: YB_ERR_TRACEBACK
            r.cells = o.cells;
            if (o.status == YB_DEFERRED) { r.C = o.C; continue; }          // runs again (device_run), with o.C MiB of traceback
            if (o.status != YB_OK) { ++f; if (o.status != YB_ERR_TRACEBACK) continue; }
            else c += o.cells;
            r.m_new = o.m_new; r.C = o.C; r.D = o.D; r.I = o.I;
            // the device's 2-bit codes are the ABI's script format and the D2H copy put them in place
            r.script = ctx->scriptStore + ctx->scriptOff[(size_t)g];
        }
        cells += c; failed += f;
    });
    d.cells += cells.load();
    d.failed += failed.load();
    if (s.sum.nDeferred > 0)
        for (int64_t i = 0; i < s.count; ++i)
            if (outs[i].status == YB_DEFERRED) d.deferred.push_back(s.first + i);
    d.unpack_ms += now_ms() - t0;
}

void reset_stats(Device &d) {
    d.kernel_ms = d.fill_ms = d.profile_ms = d.tb_ms = d.h2d_ms = d.d2h_ms = d.pack_ms = d.unpack_ms = 0;
    d.h2d_bytes = d.d2h_bytes = d.cells = d.failed = 0;
    d.launches = d.waves = 0;
    d.err.clear();
    d.errJob = 0;
    d.plan_ms = 0; d.staged_bytes = 0; d.firstFailed = -1; d.deferred.clear();
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
        st->plan_ms = std::max(st->plan_ms, d.plan_ms);
        st->staged_bytes += d.staged_bytes;
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
        // (no ramp down: a small wave's fill is as long as its longest pair, and nothing on the host waits for the last one)
        if (ndev > 1) target = std::min(target, std::max(tailBytes, remaining / (size_t)ndev));      // ... but devices share the end
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

// One device's share of a batch: grab waves until none are left.  Every wave is queued in one go on its slot's stream
// (copies, K0, K1, K2, K3, result copies); a slot is waited for and unpacked when the ring comes back to it, NSLOTS waves
// later, or at the end.  Pairs that K0 deferred (their traceback matrix no longer fitted their wave's pool) run again at the
// end, one per wave, with a pool of their own size.
int device_run(yb_ctx *ctx, Device &d, Dispatcher &disp, yb_result *results) {
    if (cudaSetDevice(d.id) != cudaSuccess) { d.err = "cudaSetDevice failed"; return YB_ERR_CUDA; }
    int rc = YB_OK, next = 0;
    if (const char *e = getenv("YB_PROFILE")) d.timeline = atoi(e) >= 2;
    if (d.timeline) {
        if (!d.evBase) cudaEventCreate(&d.evBase);
        d.tBase = now_ms();
        cudaEventRecord(d.evBase, d.slots[0].stream);
    }
    int64_t lo = 0, hi = 0;
    while (disp.grab(lo, hi)) {
        Slot &s = d.slots[next];
        if (s.busy) {                            // the wave that used this slot NSLOTS rounds ago
            if ((rc = slot_wait(d, s)) != YB_OK) break;
            slot_unpack(ctx, d, s, results);
        }
        if ((rc = slot_prepare(ctx, d, s, disp.jobs, lo, hi - lo, ctx->waveTbBytes)) != YB_OK) break;
        // a wave's fill is as long as its longest pair whatever its size: with half the machine per wave, the fills of
        // consecutive waves overlap instead (measured on cfg2: 13.8 -> 12.7 ms end to end)
        if ((rc = slot_launch(d, s, true, d.fillSplit)) != YB_OK) break;
        next = (next + 1) % NSLOTS;
    }
    for (int k = 0; k < NSLOTS; ++k) {           // drain, oldest first
        Slot &s = d.slots[(next + k) % NSLOTS];
        if (!s.busy) continue;
        int r2 = slot_wait(d, s);
        if (r2 != YB_OK) { if (rc == YB_OK) rc = r2; continue; }
        if (rc == YB_OK) slot_unpack(ctx, d, s, results);
    }
    // Deferred pairs (rare): their bin had no fill launch in their wave, or the wave's traceback pool was full.  They run
    // again as waves of their own -- every bin launched, traceback pool sized for the largest of them -- through a gathered
    // copy of their jobs; results and scripts land where the first attempt would have put them.
    for (int attempt = 0; attempt < 3 && rc == YB_OK && !d.deferred.empty(); ++attempt) {
        std::vector<int64_t> again;
        again.swap(d.deferred);
        std::sort(again.begin(), again.end());
        size_t needMax = 0;
        for (int64_t g : again) needMax = std::max(needMax, ((size_t)std::max(1, results[g].C) + 1) << 20);
        const size_t cap = std::max(ctx->waveTbBytes, needMax);
        size_t k = 0;
        while (k < again.size() && rc == YB_OK) {
            // contiguous runs of job indices form a wave (scripts go to the batch store by job index)
            size_t e = k + 1, tbSum = ((size_t)std::max(1, results[again[k]].C) + 1) << 20;
            while (e < again.size() && again[e] == again[e - 1] + 1) {
                const size_t nb = ((size_t)std::max(1, results[again[e]].C) + 1) << 20;
                if (tbSum + nb > cap || e - k >= 65536) break;
                tbSum += nb; ++e;
            }
            Slot &s = d.slots[0];
            if ((rc = slot_prepare(ctx, d, s, disp.jobs, again[k], (int64_t)(e - k), std::min(cap, tbSum + ((size_t)16 << 20)), true)) != YB_OK) break;
            if ((rc = slot_launch(d, s, true)) != YB_OK) break;
            if ((rc = slot_wait(d, s)) != YB_OK) break;
            slot_unpack(ctx, d, s, results);
            k = e;
        }
    }
    if (rc == YB_OK && !d.deferred.empty()) { d.err = "a pair's traceback matrix does not fit the device"; rc = YB_ERR_LIMIT; }
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

// the message of a pair the device flagged, in the reference's wording where it has one
void describe_failure(yb_ctx *ctx, const yb_job &j, int64_t index, int status) {
    char msg[256];
    msg[0] = 0;
    if (status == YB_ERR_TRACEBACK) { set_err(ctx, "Error generating edit script."); return; }
    if (!dims_ok(j)) { set_err(ctx, "job %lld: bad dimensions K=%d M=%d L=%d N=%d", (long long)index, j.K, j.M, j.L, j.N); return; }
    int wmax = 0;
    if (check_band(j.M, j.N, j.LB, j.RB, msg, sizeof msg, &wmax) < 0) { set_err(ctx, "%s", msg); return; }   // reference wording
    if (j.K > ctx->maxDepth || j.L > 255)
        set_err(ctx, "job %lld: profile depth K=%d L=%d exceeds the kernel limit (%d/255 rows)", (long long)index, j.K, j.L, ctx->maxDepth);
    else
        set_err(ctx, "job %lld: band row of %d cells exceeds the kernel limit (%d)", (long long)index, wmax, max_band_row());
}

template <class F>
int for_each_device(yb_ctx *ctx, F &&fn) {
    int ndev = (int)ctx->devs.size();
    std::vector<int> rcs((size_t)ndev, YB_OK);
    if (ndev == 1) rcs[0] = fn(0);
    else {
        std::vector<std::thread> th;
        for (int d = 0; d < ndev; ++d)
            th.emplace_back([&, d] {
                const Device &dv = ctx->devs[(size_t)d];          // (a thread of ours: it may sit next to its device)
                if (CPU_COUNT(&dv.cpus) > 0) pthread_setaffinity_np(pthread_self(), sizeof(cpu_set_t), &dv.cpus);
                rcs[(size_t)d] = fn(d);
            });
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
    // A wave runs on ten streams (copies + short kernels, one per fill bin) and eight waves are in flight: with the
    // default of 8 hardware work queues unrelated streams share a queue and wait for each other -- a wave's copy behind
    // another wave's fill (measured on cfg2: 12.2 -> 11.1 ms per end-to-end call with 32).  Only effective when this call
    // is what creates the process's CUDA context; a host that initialises CUDA itself sets the variable itself (bench.py).
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
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
        fprintf(stderr, "yama_b200[profile] start-up: driver %.0f ms, devices %.0f ms (device 0: context %.0f ms, streams + events %.0f ms)\n", tc1 - tc0,
                now_ms() - tc1, ctx->devs[0].tContext, ctx->devs[0].tStreams);
    int hw = (int)std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    ctx->nThreads = std::min(hw, 32);
    if (const char *e = getenv("YB_THREADS")) ctx->nThreads = std::max(1, atoi(e));
    ctx->waveEnv = getenv("YB_WAVE_MB") || getenv("YB_WAVE_MIN_MB") || getenv("YB_WAVE_TAIL_MB");
    if (const char *e = getenv("YB_WAVE_MB")) ctx->waveInBytes = std::max<size_t>(1, (size_t)strtoull(e, nullptr, 10)) << 20;
    if (const char *e = getenv("YB_WAVE_MIN_MB")) ctx->waveMinBytes = std::max<size_t>(1, (size_t)strtoull(e, nullptr, 10)) << 20;
    if (const char *e = getenv("YB_WAVE_TAIL_MB")) ctx->waveTailBytes = std::max<size_t>(1, (size_t)strtoull(e, nullptr, 10)) << 20;
    {   // traceback pool of a wave (one per slot, NSLOTS slots per device): 4 GB, less on a small or crowded device.  A
        // wave whose pairs need more has the overflowing pairs deferred by K0 and run again at the end (device_run).
        size_t cap = (size_t)4 << 30;
        for (auto &d : ctx->devs) {
            size_t fr = 0, tot = 0;
            if (cudaSetDevice(d.id) == cudaSuccess && cudaMemGetInfo(&fr, &tot) == cudaSuccess) cap = std::min(cap, fr / (3 * NSLOTS));
        }
        ctx->waveTbBytes = std::max<size_t>(cap, (size_t)64 << 20);
    }
    if (const char *e = getenv("YB_WAVE_TB_MB")) ctx->waveTbBytes = std::max<size_t>(1, (size_t)strtoull(e, nullptr, 10)) << 20;
    if (const char *e = getenv("YB_DIRECT")) ctx->directCopy = atoi(e) != 0;
    if (const char *e = getenv("YB_BAND_PACK")) ctx->bandPack = atoi(e) != 0;
    if (const char *e = getenv("YB_ONLY_BINS")) for (auto &d : ctx->devs) d.onlyBins = (int)strtol(e, nullptr, 0);
    if (const char *e = getenv("YB_FILL_SPLIT")) for (auto &d : ctx->devs) d.fillSplit = std::max(1, atoi(e));
    if (const char *e = getenv("YB_SLACK")) ctx->slackBulk = std::max(1, atoi(e));
    if (const char *e = getenv("YB_UNGATED")) ctx->ungatedOk = atoi(e) != 0;
    if (!ctx->ungatedOk) ctx->maxCls = 0;   // every pair through the kernels that carry the existence multipliers (fill_body)
    if (const char *e = getenv("YB_KEYED")) if (atoi(e) == 0) ctx->maxCls = std::min(ctx->maxCls, 1);
    if (const char *e = getenv("YB_FILL2")) if (atoi(e) == 0) ctx->maxCls = 0;
    if (const char *e = getenv("YB_TB_LONG")) ctx->tbLong = std::max(1, atoi(e));
    if (const char *e = getenv("YB_WAVE_PAIRS")) ctx->wavePairs = std::max<int64_t>(1, atoll(e));
    for (auto &d : ctx->devs) {
        d.helpers = std::max(1, ctx->nThreads / (int)ctx->devs.size());
        device_cpus(d);
        d.pool.reset(new Pool(d.helpers - 1, &d.cpus));
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
            for (PinBuf *b : {&s.hIn, &s.hOut, &s.hPlan}) b->release();
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
    for (auto &b : ctx->hostBlocks) cudaFreeHost(const_cast<unsigned char *>(b.base));
    delete ctx;
}

void *yb_host_alloc(yb_ctx *ctx, size_t bytes) {
    if (!ctx) return nullptr;
    void *p = nullptr;
    const size_t want = align_up(bytes + 512, 4096);               // (slack: copies are rounded to whole 64-byte lines)
    if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    ctx->hostBlocks.push_back(HostBlock{static_cast<const unsigned char *>(p), want});
    return p;
}

void yb_host_free(yb_ctx *ctx, void *p) {
    if (!ctx || !p) return;
    for (size_t k = 0; k < ctx->hostBlocks.size(); ++k)
        if (ctx->hostBlocks[k].base == p) {
            ctx->hostBlocks.erase(ctx->hostBlocks.begin() + (long)k);
            cudaFreeHost(p);
            return;
        }
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
    {   // deepest first profile whose weights still fit 16 bits times 4 (the KEYED class, see yb_plan_kernel)
        const long long per = std::max<long long>((long long)GO + gap_extend, 2ll * maxabs);
        const int keyedMaxK = (int)std::min<long long>(255, 32767 / (4 * std::max<long long>(per, 1)));
        for (auto &d : ctx->devs) { d.sc = sc; d.maxCls = ctx->maxCls; d.keyedMaxK = keyedMaxK; }
    }
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
    // a wave costs a fixed chain of kernel launches and the latency of its longest pair whatever its size: large batches run
    // in larger waves (about ten per device), bounded by the device memory a wave's traceback and record pools take
    disp.maxBytes = std::max(ctx->waveInBytes, std::min<size_t>((size_t)512 << 20, ctx->batchBlobBytes / (10 * ctx->devs.size())));
    disp.maxPairs = ctx->wavePairs;
    disp.minBytes = std::min(ctx->waveInBytes, ctx->waveMinBytes);
    disp.tailBytes = std::min(ctx->waveInBytes, ctx->waveTailBytes);
    // A small batch (one merge step of a 10 Mb pipeline is 50-80 MB) is one trip through the pipeline, not a stream of waves:
    // cut it finer, so that the copy of one piece overlaps the kernels of the one before (measured on the real merge's
    // batches, profiles/r2_ab_waves.txt: 5.5 -> 4.6 ms per call with 4 / 16 MB waves; 8 MB waves lose it again to launches)
    if (!ctx->waveEnv) {
        const size_t perDev = ctx->batchBlobBytes / std::max<size_t>(1, ctx->devs.size());
        const size_t steady = std::min(ctx->waveInBytes, std::max<size_t>((size_t)16 << 20, perDev / 8));
        if (steady < disp.maxBytes) {
            disp.maxBytes = steady;
            disp.minBytes = disp.tailBytes = std::min(disp.minBytes, std::max<size_t>((size_t)4 << 20, steady / 4));
        }
    }
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
    // per-pair failures: report the first in job order, in the reference's wording where it has one (the device only
    // flags a pair; the words come from the scalar restatement of mz_yama.c:58-71 on that one pair)
    int64_t failed = 0, firstBad = -1;
    for (auto &d : ctx->devs) {
        failed += d.failed;
        if (d.firstFailed >= 0 && (firstBad < 0 || d.firstFailed < firstBad)) firstBad = d.firstFailed;
    }
    if (failed == 0) return YB_OK;                  // (the usual case: no scan of the results)
    if (firstBad < 0 || results[firstBad].status == YB_OK)
        for (firstBad = 0; firstBad < n && results[firstBad].status == YB_OK; ++firstBad) {}
    if (firstBad >= n) return YB_OK;
    describe_failure(ctx, jobs[firstBad], firstBad, results[firstBad].status);
    return results[firstBad].status;
}

int yb_resident_load(yb_ctx *ctx, int64_t n, const yb_job *jobs) {
    if (!ctx || n < 1 || !jobs) return YB_ERR_ARG;
    if (!ctx->scoresSet) { set_err(ctx, "yb_set_scores has not been called"); return YB_ERR_SCORES; }
    ctx->resJobs.assign(jobs, jobs + n);
    if (prepare_script_store(ctx, n, jobs) != YB_OK) { set_err(ctx, "cudaHostAlloc failed for the script store"); return YB_ERR_CUDA; }
    for (auto &d : ctx->devs) { reset_stats(d); d.hasResident = false; }
    // static split over the devices by a dimension-only cost (rows + columns of a pair; the bands are only read on the device)
    std::vector<int64_t> cost((size_t)n, 0);
    for (int64_t i = 0; i < n; ++i) cost[(size_t)i] = dims_ok(jobs[i]) ? 64ll * ((int64_t)jobs[i].M + 1) : 0;
    const int ndev = (int)ctx->devs.size();
    ctx->resSplit.assign((size_t)ndev + 1, 0);
    plan_split(n, cost.data(), ndev, ctx->resSplit.data());
    int rc = for_each_device(ctx, [&](int di) -> int {
        Device &d = ctx->devs[(size_t)di];
        if (cudaSetDevice(d.id) != cudaSuccess) return (int)YB_ERR_CUDA;
        int64_t lo = ctx->resSplit[(size_t)di], cnt = ctx->resSplit[(size_t)di + 1] - lo;
        if (cnt <= 0) return (int)YB_OK;
        Slot &s = d.slots[0];
        // (the whole share in one wave: its traceback pool is as large as the device allows)
        size_t fr = 0, tot = 0;
        CUDA_TRY(d, cudaMemGetInfo(&fr, &tot));
        int rc2 = slot_prepare(ctx, d, s, ctx->resJobs.data(), lo, cnt, fr / 2, true);
        if (rc2 != YB_OK) return rc2;
        PlanSummary sum;
        CUDA_TRY(d, cudaMemcpyAsync(&sum, static_cast<unsigned char *>(s.dIn.p) + s.summaryOff, sizeof sum, cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(d, cudaStreamSynchronize(s.stream));
        s.sum = sum;
        if (sum.nFailed) { d.err = "yb_resident_load: the batch holds invalid pairs (use yb_run_batch for per-pair status)"; return (int)YB_ERR_ARG; }
        if (sum.nDeferred) { d.err = "resident batch does not fit"; return (int)YB_ERR_LIMIT; }
        s.launchMask = 0;
        for (int b = 0; b < NBINS; ++b) {
            s.gridHint[b] = sum.binStart[b + 1] - sum.binStart[b];
            if (s.gridHint[b] > 0) s.launchMask |= 1u << b;
        }
        d.hasResident = true;
        return (int)YB_OK;
    });
    ctx->resCells = 0;
    for (auto &d : ctx->devs) if (d.hasResident) ctx->resCells += d.slots[0].sum.cells;
    return rc;
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
        int r = slot_launch(d, d.slots[0], false);
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
