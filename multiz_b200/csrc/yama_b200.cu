// yama_b200.cu -- host runtime + C ABI (include/yama_b200.h) around the sm_100a kernels.
//
// One context owns 1..8 devices.  A batch of independent block pairs is cut into contiguous,
// cell-balanced ranges (one per device, no collective: the merge has no cross-pair dependency,
// SURVEY §8(e)); each device processes its range in waves sized to its staging buffers:
//   pack (host, pinned) -> H2D -> K1 profile -> K2 fill (per ring-size bin) -> K3 traceback -> D2H.
// There is no CPU implementation of the DP in this library.
#include "../../include/yama_b200.h"
#include "yama_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace yb;

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t need) {
        if (need <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = need + need / 8 + (1u << 20);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, need); want = need; }
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
        size_t want = need + need / 8 + (1u << 20);
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

constexpr int NBINS = 4;
const int kRingOf[NBINS] = {128, 512, 2048, 4096};

struct JobInfo {           // host-side facts about one pair
    int64_t cells = 0;     // tback_size of the reference
    int64_t tbBytes = 0;   // traceback bytes (window-major layout: 32 B per wavefront step)
    int nSteps = 0;        // wavefront steps (schedule below)
    int wmax = 0;          // widest band row
    int status = YB_OK;
};

struct Wave {              // everything needed to (re)launch the kernels of one wave
    int64_t first = 0, count = 0;          // job range [first, first+count)
    size_t blobBytes = 0, metaBytes = 0;
    size_t rowRecs = 0, colRecs = 0, tbBytes = 0, scriptBytes = 0;
    std::vector<int> order;                // pair indices (within wave) grouped by bin, big first
    int binStart[NBINS + 1] = {0, 0, 0, 0, 0};
    std::vector<uint64_t> scriptOff;       // per pair, offset in the wave's script pool
};

struct Device {
    int id = -1;
    int sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    DevBuf dIn, dRow, dCol, dTb, dScript, dOut, dOrder, dQueue;
    PinBuf hIn, hScript, hOut;
    int fillBlocks[NBINS] = {0, 0, 0, 0};
    // accumulated stats of the current call
    double kernel_ms = 0, fill_ms = 0, profile_ms = 0, tb_ms = 0, h2d_ms = 0, d2h_ms = 0, pack_ms = 0;
    int64_t h2d_bytes = 0, d2h_bytes = 0;
    int launches = 0;
    std::string err;
    Wave resident;         // resident mode: the loaded wave
    bool hasResident = false;
};

}  // namespace

struct yb_ctx {
    std::vector<Device> devs;
    std::string err;
    bool scoresSet = false;
    ScoreConst sc{};
    int maxDepth = 255;
    size_t waveBytes = (size_t)24 << 30;    // device working-set budget per wave
    size_t stageBytes = (size_t)1 << 30;    // pinned input budget per wave
    // results of the last batch
    std::vector<uint8_t> scriptStore;
    std::vector<uint64_t> scriptOff;
    // record/replay queue
    std::vector<uint8_t> arena;
    struct QJob { int K, M, L, N; size_t offA, offB, offLB, offRB; };
    std::vector<QJob> queued;
    std::vector<yb_result> queuedRes;
    // resident mode
    std::vector<yb_job> resJobs;
    std::vector<JobInfo> resInfo;
    std::vector<int64_t> resSplit;          // device d owns jobs [resSplit[d], resSplit[d+1])
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

int bin_of(int wmax) {
    for (int b = 0; b < NBINS; ++b)
        if (wmax + 32 <= kRingOf[b]) return b;
    return -1;
}

// warps per CTA follow from the shared memory a ring of that size needs
constexpr int warps_of(int bin) { return bin <= 1 ? 8 : (bin == 2 ? 4 : 1); }

size_t fill_smem(int bin) {
    // rings (RING*16-aligned, hence the slack) + 1 KB of mailboxes per warp
    return (size_t)warps_of(bin) * ((size_t)kRingOf[bin] * 16 + 1024) + (size_t)kRingOf[bin] * 16;
}

}  // namespace

// ---- kernels with a runtime warps-per-CTA: thin wrappers around the template ---------------------
namespace yb {
template <int RING, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
yb_fill_kernel_w(const PairMeta *__restrict__ metas, const int *__restrict__ order, int nPairs,
                 int *__restrict__ queue, const RowRec *__restrict__ rowPool,
                 const ColRec *__restrict__ colPool, unsigned char *__restrict__ tbPool,
                 PairOut *__restrict__ outs) {
    fill_body<RING, WARPS>(metas, order, nPairs, queue, rowPool, colPool, tbPool, outs);
}
}  // namespace yb

namespace {

typedef void (*FillFn)(const PairMeta *, const int *, int, int *, const RowRec *, const ColRec *,
                       unsigned char *, PairOut *);
FillFn fill_fn(int bin) {
    switch (bin) {
        case 0: return yb_fill_kernel_w<128, 8>;
        case 1: return yb_fill_kernel_w<512, 8>;
        case 2: return yb_fill_kernel_w<2048, 4>;
        default: return yb_fill_kernel_w<4096, 1>;
    }
}

int device_init(yb_ctx *ctx, Device &d) {
    CUDA_TRY(d, cudaSetDevice(d.id));
    cudaDeviceProp prop;
    CUDA_TRY(d, cudaGetDeviceProperties(&prop, d.id));
    d.sms = prop.multiProcessorCount;
    CUDA_TRY(d, cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    for (auto &e : d.ev) CUDA_TRY(d, cudaEventCreate(&e));
    for (int b = 0; b < NBINS; ++b) {
        FillFn fn = fill_fn(b);
        size_t sm = fill_smem(b);
        CUDA_TRY(d, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        int occ = 0;
        CUDA_TRY(d, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, warps_of(b) * 32, sm));
        if (occ < 1) occ = 1;
        d.fillBlocks[b] = occ * d.sms;
    }
    (void)ctx;
    return YB_OK;
}

// Host-side per-job facts; mirrors the validation loop of mz_yama.c:58-71.
int64_t check_band(int M, int N, const int *LB, const int *RB, char *msg, int msglen, int64_t *tbBytes, int *wmax) {
    if (LB[0] != 0 || RB[M] != N) {
        if (msg) snprintf(msg, msglen, "LB and RB not terminated properly: %d %d %d", LB[0], RB[M], N);
        return YB_ERR_BAND;
    }
    const int need = N < 10 ? N : 10;
    int64_t cells = 0, tb = 0;
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
    (void)tb;
    if (wmax) *wmax = wm;
    return cells;
}

// Wavefront schedule (see K2): rows 32b+1..32b+32 run on lanes 0..31 with column = step - (OFF_b + lane).
// OFF grows per block by at least 32 (lane 0 stays behind lane 31 of the previous block) and by enough
// that a lane starts its next row only after the row below its current one has stopped reading it.
// Returns the step count; sched (may be null) receives ceil(M/32) block offsets.
int make_schedule(int M, const int *LB, const int *RB, int *sched) {
    int off = 0;
    const int nblk = (M + 31) >> 5;
    for (int b = 0; b < nblk; ++b) {
        if (sched) sched[b] = off;
        int need = 32;
        const int r0 = 32 * b + 1, r1 = std::min(M - 32, 32 * b + 32);
        for (int r = r0; r <= r1; ++r) need = std::max(need, RB[r + 1] - LB[r + 32] + 3);
        if (b == nblk - 1) {
            int lane = (M - 1) & 31;
            int last = off + lane + RB[M];                 // step of the last cell
            return ((last + 2) + 3) & ~3;                  // +1 step to publish the final scores, whole windows
        }
        off += need;
    }
    return 4;
}

int analyse_jobs(yb_ctx *ctx, int64_t n, const yb_job *jobs, std::vector<JobInfo> &info) {
    info.resize((size_t)n);
    int rc = YB_OK;
    for (int64_t i = 0; i < n; ++i) {
        const yb_job &j = jobs[i];
        JobInfo &ji = info[(size_t)i];
        char msg[256];
        if (j.K < 1 || j.L < 1 || j.M < 1 || j.N < 1 || !j.A || !j.B || !j.LB || !j.RB) {
            ji.status = YB_ERR_ARG;
            if (rc == YB_OK) { set_err(ctx, "job %lld: bad dimensions K=%d M=%d L=%d N=%d", (long long)i, j.K, j.M, j.L, j.N); rc = YB_ERR_ARG; }
            continue;
        }
        int64_t cells = check_band(j.M, j.N, j.LB, j.RB, msg, sizeof msg, &ji.tbBytes, &ji.wmax);
        if (cells < 0) {
            ji.status = YB_ERR_BAND;
            if (rc == YB_OK) { ctx->err = msg; rc = YB_ERR_BAND; }
            continue;
        }
        ji.cells = cells;
        ji.nSteps = make_schedule(j.M, j.LB, j.RB, nullptr);
        ji.tbBytes = (int64_t)ji.nSteps * 32;
        if (j.K > ctx->maxDepth || j.L > 255) {
            ji.status = YB_ERR_LIMIT;
            if (rc == YB_OK) { set_err(ctx, "job %lld: profile depth K=%d L=%d exceeds the kernel limit (%d/255 rows)", (long long)i, j.K, j.L, ctx->maxDepth); rc = YB_ERR_LIMIT; }
            continue;
        }
        if (bin_of(ji.wmax) < 0) {
            ji.status = YB_ERR_LIMIT;
            if (rc == YB_OK) { set_err(ctx, "job %lld: band row of %d cells exceeds the kernel limit (%d)", (long long)i, ji.wmax, kRingOf[NBINS - 1] - 32); rc = YB_ERR_LIMIT; }
            continue;
        }
    }
    return rc;
}

struct Need { size_t blob, rows, cols, tb, script; };
inline Need need_of(const yb_job &j, const JobInfo &ji) {
    Need n;
    n.blob = align_up((size_t)j.K * j.M, 16) + align_up((size_t)j.L * j.N, 16) + 2 * align_up((size_t)(j.M + 1) * 4, 16) +
             align_up((size_t)((j.M + 31) >> 5) * 4, 16);
    n.rows = (size_t)j.M + 1;
    n.cols = (size_t)j.N + 1;
    n.tb = align_up((size_t)ji.tbBytes, 16);
    n.script = align_up((size_t)j.M + j.N, 4);
    return n;
}

// Pack jobs [first,first+count) into the device's pinned buffer and upload.  Fills `w`.
int wave_upload(yb_ctx *ctx, Device &d, const yb_job *jobs, const std::vector<JobInfo> &info, int64_t first,
                int64_t count, Wave &w) {
    double t0 = now_ms();
    w.first = first; w.count = count;
    w.metaBytes = align_up((size_t)count * sizeof(PairMeta), 256);
    size_t blob = w.metaBytes, rows = 0, cols = 0, tb = 0, script = 0;
    int nvalid = 0;
    for (int64_t i = 0; i < count; ++i) {
        const JobInfo &ji = info[(size_t)(first + i)];
        if (ji.status != YB_OK) continue;
        Need n = need_of(jobs[first + i], ji);
        blob += n.blob; rows += n.rows; cols += n.cols; tb += n.tb; script += n.script;
        ++nvalid;
    }
    w.blobBytes = blob; w.rowRecs = rows; w.colRecs = cols; w.tbBytes = tb; w.scriptBytes = script;
    CUDA_TRY(d, d.hIn.reserve(blob));
    CUDA_TRY(d, d.dIn.reserve(blob));
    CUDA_TRY(d, d.dRow.reserve(rows * sizeof(RowRec) + 64));
    CUDA_TRY(d, d.dCol.reserve(cols * sizeof(ColRec) + 64));
    CUDA_TRY(d, d.dTb.reserve(tb + 64));
    CUDA_TRY(d, d.dScript.reserve(script + 64));
    CUDA_TRY(d, d.dOut.reserve((size_t)count * sizeof(PairOut) + 64));
    CUDA_TRY(d, d.dOrder.reserve((size_t)count * 4 + 64));
    CUDA_TRY(d, d.dQueue.reserve(64));
    CUDA_TRY(d, d.hScript.reserve(script + 64));
    CUDA_TRY(d, d.hOut.reserve((size_t)count * sizeof(PairOut) + 64));

    unsigned char *h = static_cast<unsigned char *>(d.hIn.p);
    PairMeta *metas = reinterpret_cast<PairMeta *>(h);
    size_t off = w.metaBytes;
    size_t rowBase = 0, colBase = 0, tbBase = 0, scriptBase = 0;
    w.scriptOff.assign((size_t)count, 0);
    std::vector<std::pair<int64_t, int>> binned[NBINS];
    for (int64_t i = 0; i < count; ++i) {
        const yb_job &j = jobs[first + i];
        const JobInfo &ji = info[(size_t)(first + i)];
        PairMeta pm;
        memset(&pm, 0, sizeof pm);
        if (ji.status != YB_OK) { metas[i] = pm; continue; }
        pm.K = j.K; pm.M = j.M; pm.L = j.L; pm.N = j.N;
        pm.offA = off; memcpy(h + off, j.A, (size_t)j.K * j.M); off += align_up((size_t)j.K * j.M, 16);
        pm.offB = off; memcpy(h + off, j.B, (size_t)j.L * j.N); off += align_up((size_t)j.L * j.N, 16);
        pm.offLB = off; memcpy(h + off, j.LB, (size_t)(j.M + 1) * 4); off += align_up((size_t)(j.M + 1) * 4, 16);
        pm.offRB = off; memcpy(h + off, j.RB, (size_t)(j.M + 1) * 4); off += align_up((size_t)(j.M + 1) * 4, 16);
        pm.offSched = off; pm.nSteps = make_schedule(j.M, j.LB, j.RB, reinterpret_cast<int *>(h + off));
        off += align_up((size_t)((j.M + 31) >> 5) * 4, 16);
        Need n = need_of(j, ji);
        pm.rowBase = rowBase; rowBase += n.rows;
        pm.colBase = colBase; colBase += n.cols;
        pm.tbBase = tbBase; tbBase += n.tb;
        pm.scriptBase = scriptBase; w.scriptOff[(size_t)i] = scriptBase; scriptBase += n.script;
        metas[i] = pm;
        binned[bin_of(ji.wmax)].push_back({ji.cells, (int)i});
    }
    // launch order: per ring bin, largest pairs first (longest-processing-time-first on the warp queue)
    w.order.clear();
    for (int b = 0; b < NBINS; ++b) {
        w.binStart[b] = (int)w.order.size();
        std::sort(binned[b].begin(), binned[b].end(), [](const std::pair<int64_t, int> &x, const std::pair<int64_t, int> &y) {
            return x.first != y.first ? x.first > y.first : x.second < y.second;
        });
        for (auto &pr : binned[b]) w.order.push_back(pr.second);
    }
    w.binStart[NBINS] = (int)w.order.size();
    d.pack_ms += now_ms() - t0;

    CUDA_TRY(d, cudaEventRecord(d.ev[0], d.stream));
    CUDA_TRY(d, cudaMemcpyAsync(d.dIn.p, d.hIn.p, blob, cudaMemcpyHostToDevice, d.stream));
    if (!w.order.empty())
        CUDA_TRY(d, cudaMemcpyAsync(d.dOrder.p, w.order.data(), w.order.size() * 4, cudaMemcpyHostToDevice, d.stream));
    CUDA_TRY(d, cudaEventRecord(d.ev[1], d.stream));
    CUDA_TRY(d, cudaStreamSynchronize(d.stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]);
    d.h2d_ms += ms;
    d.h2d_bytes += (int64_t)(blob + w.order.size() * 4);
    (void)ctx; (void)nvalid;
    return YB_OK;
}

// Launch K1,K2,K3 for an uploaded wave; device-timed.
int wave_compute(Device &d, const Wave &w) {
    if (w.order.empty()) return YB_OK;
    const PairMeta *metas = static_cast<const PairMeta *>(d.dIn.p);
    const unsigned char *blob = static_cast<const unsigned char *>(d.dIn.p);
    RowRec *rows = static_cast<RowRec *>(d.dRow.p);
    ColRec *cols = static_cast<ColRec *>(d.dCol.p);
    unsigned char *tb = static_cast<unsigned char *>(d.dTb.p);
    unsigned char *script = static_cast<unsigned char *>(d.dScript.p);
    PairOut *outs = static_cast<PairOut *>(d.dOut.p);
    int *order = static_cast<int *>(d.dOrder.p);
    int *queue = static_cast<int *>(d.dQueue.p);

    CUDA_TRY(d, cudaEventRecord(d.ev[2], d.stream));
    CUDA_TRY(d, cudaMemsetAsync(outs, 0, (size_t)w.count * sizeof(PairOut), d.stream));
    CUDA_TRY(d, cudaMemsetAsync(queue, 0, 64, d.stream));
    yb_profile_kernel<<<(unsigned)w.count, K1_THREADS, 0, d.stream>>>(metas, blob, rows, cols);
    d.launches++;
    CUDA_TRY(d, cudaEventRecord(d.ev[3], d.stream));
    for (int b = 0; b < NBINS; ++b) {
        int n = w.binStart[b + 1] - w.binStart[b];
        if (n <= 0) continue;
        int wpc = warps_of(b);
        int blocks = std::min((n + wpc - 1) / wpc, d.fillBlocks[b]);
        fill_fn(b)<<<blocks, wpc * 32, fill_smem(b), d.stream>>>(metas, order + w.binStart[b], n, queue + b, rows, cols, tb, outs);
        d.launches++;
    }
    CUDA_TRY(d, cudaEventRecord(d.ev[4], d.stream));
    {
        const int nv = (int)w.order.size();
        yb_traceback_kernel<<<(unsigned)((nv + 63) / 64), 64, 0, d.stream>>>(metas, order, nv, blob, tb, script, outs);
    }
    d.launches++;
    CUDA_TRY(d, cudaEventRecord(d.ev[5], d.stream));
    CUDA_TRY(d, cudaStreamSynchronize(d.stream));
    CUDA_TRY(d, cudaGetLastError());
    float a = 0, b = 0, c = 0;
    cudaEventElapsedTime(&a, d.ev[2], d.ev[3]);
    cudaEventElapsedTime(&b, d.ev[3], d.ev[4]);
    cudaEventElapsedTime(&c, d.ev[4], d.ev[5]);
    d.profile_ms += a; d.fill_ms += b; d.tb_ms += c;
    d.kernel_ms += a + b + c;
    return YB_OK;
}

// D2H of scores + scripts of a wave into results / the context's script store.
int wave_download(yb_ctx *ctx, Device &d, const Wave &w, const yb_job *jobs, const std::vector<JobInfo> &info,
                  yb_result *results) {
    CUDA_TRY(d, cudaEventRecord(d.ev[6], d.stream));
    CUDA_TRY(d, cudaMemcpyAsync(d.hOut.p, d.dOut.p, (size_t)w.count * sizeof(PairOut), cudaMemcpyDeviceToHost, d.stream));
    if (w.scriptBytes)
        CUDA_TRY(d, cudaMemcpyAsync(d.hScript.p, d.dScript.p, w.scriptBytes, cudaMemcpyDeviceToHost, d.stream));
    CUDA_TRY(d, cudaEventRecord(d.ev[7], d.stream));
    CUDA_TRY(d, cudaStreamSynchronize(d.stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, d.ev[6], d.ev[7]);
    d.d2h_ms += ms;
    d.d2h_bytes += (int64_t)((size_t)w.count * sizeof(PairOut) + w.scriptBytes);
    const PairOut *outs = static_cast<const PairOut *>(d.hOut.p);
    const unsigned char *hs = static_cast<const unsigned char *>(d.hScript.p);
    for (int64_t i = 0; i < w.count; ++i) {
        int64_t g = w.first + i;
        yb_result &r = results[g];
        const JobInfo &ji = info[(size_t)g];
        memset(&r, 0, sizeof r);
        r.status = ji.status;
        r.cells = ji.cells;
        if (ji.status != YB_OK) continue;
        const PairOut &o = outs[i];
        r.status = o.status;
        r.m_new = o.m_new; r.C = o.C; r.D = o.D; r.I = o.I;
        uint8_t *dst = ctx->scriptStore.data() + ctx->scriptOff[(size_t)g];
        memcpy(dst, hs + w.scriptOff[(size_t)i], (size_t)o.m_new);
        r.script = dst;
        (void)jobs;
    }
    return YB_OK;
}

void reset_stats(Device &d) {
    d.kernel_ms = d.fill_ms = d.profile_ms = d.tb_ms = d.h2d_ms = d.d2h_ms = d.pack_ms = 0;
    d.h2d_bytes = d.d2h_bytes = 0;
    d.launches = 0;
    d.err.clear();
}

void collect_stats(yb_ctx *ctx, yb_stats *st, double total_ms, int64_t cells, int64_t pairs) {
    if (!st) return;
    memset(st, 0, sizeof *st);
    for (auto &d : ctx->devs) {
        st->kernel_ms = std::max(st->kernel_ms, d.kernel_ms);
        st->h2d_ms = std::max(st->h2d_ms, d.h2d_ms);
        st->d2h_ms = std::max(st->d2h_ms, d.d2h_ms);
        st->pack_ms = std::max(st->pack_ms, d.pack_ms);
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
std::vector<int64_t> split_jobs(const std::vector<JobInfo> &info, int ndev) {
    std::vector<int64_t> cells(info.size()), cut((size_t)ndev + 1);
    for (size_t i = 0; i < info.size(); ++i) cells[i] = info[i].cells;
    plan_split((int64_t)info.size(), cells.data(), ndev, cut.data());
    return cut;
}

int run_range(yb_ctx *ctx, Device &d, const yb_job *jobs, const std::vector<JobInfo> &info, int64_t lo, int64_t hi,
              yb_result *results) {
    if (cudaSetDevice(d.id) != cudaSuccess) { d.err = "cudaSetDevice failed"; return YB_ERR_CUDA; }
    int64_t i = lo;
    while (i < hi) {
        // grow the wave until a budget is hit
        size_t blob = 0, dev = 0;
        int64_t j = i;
        while (j < hi) {
            const JobInfo &ji = info[(size_t)j];
            size_t b = sizeof(PairMeta), dv = sizeof(PairOut) + 4;
            if (ji.status == YB_OK) {
                Need n = need_of(jobs[j], ji);
                b += n.blob;
                dv += n.blob + n.rows * sizeof(RowRec) + n.cols * sizeof(ColRec) + n.tb + n.script;
            }
            if (j > i && (blob + b > ctx->stageBytes || dev + dv > ctx->waveBytes || j - i >= (1 << 24))) break;
            blob += b; dev += dv; ++j;
        }
        Wave w;
        int rc = wave_upload(ctx, d, jobs, info, i, j - i, w);
        if (rc != YB_OK) return rc;
        rc = wave_compute(d, w);
        if (rc != YB_OK) return rc;
        rc = wave_download(ctx, d, w, jobs, info, results);
        if (rc != YB_OK) return rc;
        i = j;
    }
    return YB_OK;
}

void prepare_script_store(yb_ctx *ctx, int64_t n, const yb_job *jobs, const std::vector<JobInfo> &info) {
    ctx->scriptOff.assign((size_t)n, 0);
    size_t tot = 0;
    for (int64_t i = 0; i < n; ++i) {
        ctx->scriptOff[(size_t)i] = tot;
        if (info[(size_t)i].status == YB_OK) tot += (size_t)jobs[i].M + jobs[i].N;
    }
    ctx->scriptStore.resize(tot + 16);
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
    if (cudaGetDeviceCount(&avail) != cudaSuccess || avail < 1) return YB_ERR_CUDA;
    yb_ctx *ctx = new yb_ctx();
    std::vector<int> ids;
    if (devices && ndev > 0) ids.assign(devices, devices + ndev);
    else for (int i = 0; i < avail; ++i) ids.push_back(i);
    for (int id : ids) {
        if (id < 0 || id >= avail) { delete ctx; return YB_ERR_ARG; }
        Device d;
        d.id = id;
        ctx->devs.push_back(d);
    }
    for (auto &d : ctx->devs)
        if (device_init(ctx, d) != YB_OK) { fprintf(stderr, "yama_b200: %s\n", d.err.c_str()); yb_destroy(ctx); return YB_ERR_CUDA; }
    if (const char *e = getenv("YB_WAVE_BYTES")) ctx->waveBytes = (size_t)strtoull(e, nullptr, 10);
    if (const char *e = getenv("YB_STAGE_BYTES")) ctx->stageBytes = (size_t)strtoull(e, nullptr, 10);
    *out = ctx;
    return YB_OK;
}

void yb_destroy(yb_ctx *ctx) {
    if (!ctx) return;
    for (auto &d : ctx->devs) {
        cudaSetDevice(d.id);
        for (DevBuf *b : {&d.dIn, &d.dRow, &d.dCol, &d.dTb, &d.dScript, &d.dOut, &d.dOrder, &d.dQueue}) b->release();
        for (PinBuf *b : {&d.hIn, &d.hScript, &d.hOut}) b->release();
        for (auto &e : d.ev) if (e) cudaEventDestroy(e);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
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
    sc.gap_open = GO;
    sc.gap_ext = gap_extend;
    ctx->sc = sc;
    ctx->maxDepth = std::min(255, 32767 / maxabs);
    for (auto &d : ctx->devs) {
        if (cudaSetDevice(d.id) != cudaSuccess || cudaMemcpyToSymbol(c_sc, &sc, sizeof sc) != cudaSuccess) {
            set_err(ctx, "cudaMemcpyToSymbol failed on device %d", d.id);
            return YB_ERR_CUDA;
        }
    }
    ctx->scoresSet = true;
    return YB_OK;
}

int yb_plan_split(int64_t n, const int64_t *cells, int nparts, int64_t *cuts) {
    if (n < 0 || nparts < 1 || !cuts || (n > 0 && !cells)) return YB_ERR_ARG;
    plan_split(n, cells, nparts, cuts);
    return YB_OK;
}

int64_t yb_check_band(int32_t M, int32_t N, const int32_t *LB, const int32_t *RB, char *msg, int msglen) {
    return check_band(M, N, LB, RB, msg, msglen, nullptr, nullptr);
}

int yb_run_batch(yb_ctx *ctx, int64_t n, const yb_job *jobs, yb_result *results, yb_stats *stats) {
    if (!ctx || n < 0 || (n > 0 && (!jobs || !results))) return YB_ERR_ARG;
    if (!ctx->scoresSet) { set_err(ctx, "yb_set_scores has not been called"); return YB_ERR_SCORES; }
    double t0 = now_ms();
    std::vector<JobInfo> info;
    int arc = analyse_jobs(ctx, n, jobs, info);
    prepare_script_store(ctx, n, jobs, info);
    for (auto &d : ctx->devs) reset_stats(d);
    int ndev = (int)ctx->devs.size();
    std::vector<int64_t> cut = split_jobs(info, ndev);
    int rc = for_each_device(ctx, [&](int d) {
        return run_range(ctx, ctx->devs[(size_t)d], jobs, info, cut[(size_t)d], cut[(size_t)d + 1], results);
    });
    int64_t cells = 0;
    for (auto &ji : info) if (ji.status == YB_OK) cells += ji.cells;
    collect_stats(ctx, stats, now_ms() - t0, cells, n);
    if (rc != YB_OK) return rc;
    if (arc != YB_OK) return arc;
    for (int64_t i = 0; i < n; ++i)
        if (results[i].status != YB_OK) { set_err(ctx, "Error generating edit script."); return results[i].status; }
    return YB_OK;
}

int yb_resident_load(yb_ctx *ctx, int64_t n, const yb_job *jobs) {
    if (!ctx || n < 1 || !jobs) return YB_ERR_ARG;
    if (!ctx->scoresSet) { set_err(ctx, "yb_set_scores has not been called"); return YB_ERR_SCORES; }
    ctx->resJobs.assign(jobs, jobs + n);
    int arc = analyse_jobs(ctx, n, jobs, ctx->resInfo);
    if (arc != YB_OK) return arc;
    prepare_script_store(ctx, n, jobs, ctx->resInfo);
    for (auto &d : ctx->devs) reset_stats(d);
    ctx->resSplit = split_jobs(ctx->resInfo, (int)ctx->devs.size());
    return for_each_device(ctx, [&](int di) {
        Device &d = ctx->devs[(size_t)di];
        if (cudaSetDevice(d.id) != cudaSuccess) return (int)YB_ERR_CUDA;
        d.hasResident = false;
        int64_t lo = ctx->resSplit[(size_t)di], hi = ctx->resSplit[(size_t)di + 1];
        int rc = wave_upload(ctx, d, ctx->resJobs.data(), ctx->resInfo, lo, hi - lo, d.resident);
        if (rc == YB_OK) d.hasResident = true;
        return rc;
    });
}

int yb_resident_step(yb_ctx *ctx, yb_stats *stats) {
    if (!ctx) return YB_ERR_ARG;
    double t0 = now_ms();
    for (auto &d : ctx->devs) reset_stats(d);
    int rc = for_each_device(ctx, [&](int di) {
        Device &d = ctx->devs[(size_t)di];
        if (!d.hasResident) { d.err = "no resident batch loaded"; return (int)YB_ERR_ARG; }
        if (cudaSetDevice(d.id) != cudaSuccess) return (int)YB_ERR_CUDA;
        return wave_compute(d, d.resident);
    });
    int64_t cells = 0;
    for (auto &ji : ctx->resInfo) cells += ji.cells;
    collect_stats(ctx, stats, now_ms() - t0, cells, (int64_t)ctx->resInfo.size());
    return rc;
}

int yb_resident_fetch(yb_ctx *ctx, yb_result *results) {
    if (!ctx || !results) return YB_ERR_ARG;
    return for_each_device(ctx, [&](int di) {
        Device &d = ctx->devs[(size_t)di];
        if (!d.hasResident) { d.err = "no resident batch loaded"; return (int)YB_ERR_ARG; }
        if (cudaSetDevice(d.id) != cudaSuccess) return (int)YB_ERR_CUDA;
        return wave_download(ctx, d, d.resident, ctx->resJobs.data(), ctx->resInfo, results);
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

int yb_assemble(const yb_job *job, const yb_result *res, uint8_t *out) {
    if (!job || !res || !out || !res->script) return YB_ERR_ARG;
    const int K = job->K, L = job->L, W = K + L;
    int i = 0, j = 0, m = 0;
    for (int e = res->m_new - 1; e >= 0; --e) {             // mz_yama.c:300-309
        int op = res->script[e];
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

}  // extern "C"
