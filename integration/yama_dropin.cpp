// yama_dropin.cpp -- the reference-side binding: multiz's own `yama` symbol (mz_yama.h:22), implemented on
// libyama_b200.so, plus the speculative record/replay driver that lets the UNMODIFIED reference host
// (multiz.c / multic.c / mz_preyama.c / maf.c / multi_util.c, compiled where they lie under /root/reference)
// hand thousands of block pairs to the GPU per launch instead of one at a time.
//
// Why replay works: the sequence of yama() jobs that multiz() / multih() produce depends
// only on the input files -- pre_yama() derives A, B, LB, RB from its inputs (mz_preyama.c:162-259) and the
// merge loop advances on input coordinates (multiz.c:136-174).  The one exception is v=0, whose second
// yama() call takes the first call's output as its B (mz_preyama.c:335).  So:
//
//   pass k (forked child, stdout/stderr -> /dev/null): run the reference's main().  Every yama() call whose
//          inputs are already in the result table returns the true alignment; an unknown one is shipped to
//          the parent over a pipe and answered with a shape-valid dummy (all of A, then all of B).  A call
//          whose B *is* a dummy we just returned (v=0 stage 2) is "tainted": not shipped.
//   parent: aligns all shipped jobs in ONE yb_run_batch() over all visible GPUs, stores the edit scripts.
//          Repeats while the child saw tainted calls (v=1: one speculative pass; v=0: two).
//   final pass (this process, real stdout): every call hits the table; a miss falls back to a synchronous
//          one-pair GPU call, so the output never depends on speculation being right.
//
// Results are keyed by the CONTENT of the job (dimensions + A + B + LB + RB, 128-bit hash), never by call
// order.  There is no CPU alignment code here: without a usable GPU yama() dies through fatalf(), the
// reference's own error convention (util.c:17-32).
#include "../include/yama_b200.h"
#include "yb_wire.h"

#include <algorithm>
#include <cerrno>
#include <csignal>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_set>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <malloc.h>
#include <stdio_ext.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <unistd.h>

typedef unsigned char uchar;

extern "C" {
// reference globals, common symbols of mz_scores.h:8-11, filled by init_scores70/85 (mz_scores.c:94-122)
extern int **ss;
extern int *gop;
extern int gap_open, gap_extend;
// util.c:21
void fatalf(const char *fmt, ...);
// the reference tool's own main(), renamed at compile time (-Dmain=ref_tool_main); weak so that this file
// can also be linked as a plain library behind a host that keeps its main()
int ref_tool_main(int argc, char **argv) __attribute__((weak));
void yama(uchar **A, int K, int M, uchar **B, int L, int N, int *LB, int *RB, uchar ***OAL, int *OM);
void yb_host_exit(int code) __attribute__((noreturn));
FILE *yb_host_fopen(const char *path, const char *mode);
// block scoring, called by score_dropin.c (the `mafScoreRange` symbol)
int yb_dropin_score_mode(void);
// deferred output (score_dropin.c holds the part that walks the reference's structs)
int yb_dropin_defer_active(void);
void yb_dropin_defer_block(FILE *f, void *ali_copy, int placeholder);
int yb_defer_fill(void *ali_copy, const unsigned char *al, int m_new, int W);
int yb_defer_wants_score(void *ali_copy);
int yb_defer_text_size(void *ali_copy);
int yb_defer_rows(void *ali_copy, const unsigned char **rows, int cap);
double yb_defer_host_score(void *ali_copy);
void yb_defer_write(FILE *f, void *ali_copy, int have_score, double score);
int yb_host_fclose(FILE *f);
// the tool that was linked in: multiz.c defines multiz(), multic.c does not
int multiz(void *, void *, FILE *, FILE *, int) __attribute__((weak));
double yb_dropin_score(int nrows, const unsigned char *const *rows, int text_size, int start, int size);
}

namespace {

enum Mode { DIRECT, RECORD, REPLAY, DEFER };

struct Key {
    uint64_t a, b;
    bool operator==(const Key &o) const { return a == o.a && b == o.b; }
};
struct KeyHash {
    size_t operator()(const Key &k) const { return (size_t)(k.a ^ (k.b * 0x9e3779b97f4a7c15ull)); }
};

// What identifies a job besides its 128-bit key: the dimensions and a second, independently computed digest of the
// content (CRC-32C lanes over A|B and over LB|RB).  A table hit is only trusted when these agree too, so a key
// collision between two jobs degrades to a miss (a synchronous one-pair call), never to a wrong alignment.
struct Proof {
    int32_t K = 0, M = 0, L = 0, N = 0;
    uint64_t chk = 0;
    bool operator==(const Proof &o) const { return K == o.K && M == o.M && L == o.L && N == o.N && chk == o.chk; }
};

struct Entry {            // one aligned job: its edit script (packed 2 bits per op, see yama_b200.h)
    int32_t m_new = 0;
    const uint8_t *script = nullptr;   // into one of G.scriptChunks (a chunk per batch, never reallocated)
    Proof proof;
};

struct Pending {          // a job shipped by a child, waiting for the GPU
    Key key;
    Proof proof;
    int32_t K, M, L, N;
    size_t offA, offB, offLB, offRB;   // into G.arena
};

struct Globals {
    Mode mode = DIRECT;
    yb_ctx *ctx = nullptr;
    std::unordered_map<Key, Entry, KeyHash> table;
    std::deque<std::vector<uint8_t>> scriptChunks;
    // child side
    int pipe_w = -1;
    std::vector<uint8_t> wbuf;
    std::unordered_map<Key, int, KeyHash> shipped;
    uchar **lastDummy = nullptr;           // the placeholder answered last: a later call whose B is this very table, with these
    int lastDummyRows = 0, lastDummyCols = 0;   // dimensions AND this content, is the chained call of v=0 (an address alone
    uint32_t lastDummySum = 0;                  // proves nothing: the host frees the table and malloc hands it out again)
    uint64_t nTainted = 0, nShipped = 0, nHits = 0, collisions = 0;
    // parent side
    std::vector<uint8_t> arena;
    std::vector<Pending> pending;
    // stats
    bool stats = false;
    uint64_t calls = 0, misses = 0, direct = 0;
    double gpu_ms = 0, kernel_ms = 0;
    int64_t cells = 0, batches = 0, jobs = 0, failed = 0;
    int passes = 0;
    double child_ms = 0, final_ms = 0;      // wall time of the speculative passes / of the real pass
    double create_ms = 0, t_start = 0;      // yb_create (CUDA start-up), process start
    double hit_ms = 0;                      // real pass: time inside yama() for table hits (hash + assemble)
    int scoreGpu = -1;                      // YB_SCORE=gpu: mafScoreRange on the device in the real pass
    uint64_t scoreCalls = 0;
    double score_ms = 0;
    bool debug = false;
    std::unordered_map<Key, int, KeyHash> failedKeys;   // debug: batch status of jobs that did not align
} G;

// ---- deferred output (YB_DROPIN=defer, the default): ONE pass of the host for v=1 --------------------------------------
// The host's own loop (multiz.c:60-177 / multic.c:124-196 with mz_preyama.c) runs once, in this process.  yama() keeps a
// copy of every job and answers with a placeholder alignment (see score_dropin.c); what the host prints to stdout goes to
// a memory file, and a merged block is not printed but remembered with its position in that stream.  When main() is done
// all jobs are aligned in one batch and the stream is written out with the merged blocks, now real, at their places: the
// bytes and their order are the reference's, the host ran once, and nothing was forked.
// v=0 (mz_preyama.c:265-335: a second yama() on the first one's output) takes one speculative pass for the first stage
// (a forked child, as in batch mode), then this pass for the second.
struct DeferJob { int32_t K, M, L, N; size_t offA, offB, offLB, offRB; Key key; };
struct DeferBlock { FILE *f; int64_t job; long pos; void *ali; double score; bool haveScore; };   // job < 0: not a placeholder
struct Defer {
    bool active = false, done = false;
    bool chained = false;                // v=0: calls come in pairs, the second consumes the first one's answer
    uint64_t callInPair = 0;
    std::vector<uint8_t> arena;
    std::vector<DeferJob> jobs;
    std::vector<DeferBlock> blocks;          // every block the host "wrote", in its order
    std::vector<FILE *> closed;              // streams the host closed meanwhile: closed for real when their blocks are out
    int64_t lastJob = -1;
    long stdoutClosedAt = -1;                // where the captured stream ends for the caller: the host closed stdout there
    int savedStdout = -1, memFd = -1;
    double emit_ms = 0;
} D;

// ---- streamed replay (YB_DROPIN=stream) ------------------------------------------------------------------------
// The real pass starts at once in the parent and runs BESIDE the speculative child instead of after it: a reader
// thread takes the child's jobs off the pipe, aligns them in chunks as they arrive and publishes the scripts; the
// parent's yama() waits only if its job has not been answered yet.  The child is ahead by construction (its yama()
// returns placeholders immediately), so the tool takes about one host pass instead of two.  A call the child never
// shipped is a miss once the LAST speculative pass has passed it (the child's call counter travels with its jobs) and is
// aligned synchronously -- exact as ever.  v=0 needs two speculative passes (stage 2 consumes stage 1's output): a
// spare child is forked at the start, before any thread or CUDA state exists; if the first child reports placeholder
// inputs, the spare receives the stage-1 results over a pipe and runs the second pass, again streamed.
struct Stream {
    bool active = false;
    std::mutex mu;                       // table, pendingKeys, childCalls, readerDone, starved
    std::condition_variable cv;
    std::mutex backendMu;                // one backend call at a time (reader's chunks, the real pass's direct calls)
    std::unordered_set<Key, KeyHash> pendingKeys;   // received from the child, not aligned yet
    uint64_t childCalls = 0;             // the child's yama() call index as of its last shipped job
    bool readerDone = false, starved = false;
    bool lastPass = false;               // the speculative pass now feeding us is the last one: what it skips is a miss
    std::thread reader;
    size_t chunkBytes = (size_t)16 << 20;   // a chunk is aligned when it holds this much input ...
    size_t minStarved = 256;                // ... or this many jobs while the real pass is waiting (a launch costs the
                                            // same for 1 job and for 100: never feed it job by job)
    uint64_t waits = 0;
} S;


double now_ms() {
    timeval tv;
    gettimeofday(&tv, nullptr);
    return tv.tv_sec * 1e3 + tv.tv_usec * 1e-3;
}

// 128-bit content hash: two independent multiply-xorshift lanes over 8-byte words
struct Hasher {
    uint64_t h1 = 0x243f6a8885a308d3ull, h2 = 0x13198a2e03707344ull;
    void word(uint64_t w) {
        h1 = (h1 ^ w) * 0x9fb21c651e98df25ull; h1 ^= h1 >> 32;
        h2 = (h2 + w) * 0xc2b2ae3d27d4eb4full; h2 ^= h2 >> 29;
    }
    void bytes(const void *p, size_t n) {
        const uint8_t *s = static_cast<const uint8_t *>(p);
        while (n >= 8) { uint64_t w; memcpy(&w, s, 8); word(w); s += 8; n -= 8; }
        uint64_t w = 0;
        memcpy(&w, s, n);
        word(w ^ ((uint64_t)n << 56));
    }
    Key done() {
        word(0x5851f42d4c957f2dull);
        return Key{h1 ^ (h2 >> 7), h2 ^ (h1 << 9)};
    }
};

// Callers build A[1]/B[1] as one contiguous buffer (mz_preyama.c:174-205) but the ABI only promises the
// pointer table (mz_yama.h:8-13), so gather column by column when it is not.
const uint8_t *contiguous(uchar **X, int rows, int cols, std::vector<uint8_t> &tmp) {
    bool contig = true;
    for (int i = 2; i <= cols && contig; ++i) contig = (X[i] == X[i - 1] + rows);
    if (contig) return X[1];
    tmp.resize((size_t)rows * cols);
    for (int i = 1; i <= cols; ++i) memcpy(tmp.data() + (size_t)(i - 1) * rows, X[i], (size_t)rows);
    return tmp.data();
}

Key key_of(int K, int M, int L, int N, const uint8_t *A, const uint8_t *B, const int *LB, const int *RB) {
    Hasher h;
    h.word(((uint64_t)(uint32_t)K << 32) | (uint32_t)M);
    h.word(((uint64_t)(uint32_t)L << 32) | (uint32_t)N);
    h.word(((uint64_t)(uint32_t)gap_open << 32) | (uint32_t)gap_extend);
    h.bytes(A, (size_t)K * M);
    h.bytes(B, (size_t)L * N);
    h.bytes(LB, (size_t)(M + 1) * sizeof(int));
    h.bytes(RB, (size_t)(M + 1) * sizeof(int));
    // test hook: a deliberately weak key (dimensions only) makes same-shape jobs collide, so that the Proof check is
    // what keeps the output right (tests/test_dropin_cpu.py)
    static const bool weak = getenv("YB_DROPIN_WEAK_KEY") != nullptr;
    if (weak) return Key{((uint64_t)(uint32_t)K << 32) | (uint32_t)L, 0x77ull};
    return h.done();
}

// the second digest (see Proof): bitwise CRC-32C tables, nothing shared with Hasher
uint32_t g_crcTab[8][256];
void crc_init() {
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0x82f63b78u & (0u - (c & 1u)));
        g_crcTab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
        for (int t = 1; t < 8; ++t) g_crcTab[t][i] = (g_crcTab[t - 1][i] >> 8) ^ g_crcTab[0][g_crcTab[t - 1][i] & 0xffu];
}
uint32_t crc32c(uint32_t crc, const void *p, size_t n) {
    static const bool once = (crc_init(), true);
    (void)once;
    const uint8_t *s = static_cast<const uint8_t *>(p);
    crc = ~crc;
    while (n >= 8) {                                   // slicing-by-8
        uint64_t w;
        memcpy(&w, s, 8);
        w ^= crc;
        crc = g_crcTab[7][w & 0xff] ^ g_crcTab[6][(w >> 8) & 0xff] ^ g_crcTab[5][(w >> 16) & 0xff] ^ g_crcTab[4][(w >> 24) & 0xff] ^
              g_crcTab[3][(w >> 32) & 0xff] ^ g_crcTab[2][(w >> 40) & 0xff] ^ g_crcTab[1][(w >> 48) & 0xff] ^ g_crcTab[0][w >> 56];
        s += 8; n -= 8;
    }
    while (n--) crc = (crc >> 8) ^ g_crcTab[0][(crc ^ *s++) & 0xffu];
    return ~crc;
}
Proof proof_of(int K, int M, int L, int N, const uint8_t *A, const uint8_t *B, const int *LB, const int *RB) {
    Proof p;
    p.K = K; p.M = M; p.L = L; p.N = N;
    const uint32_t text = crc32c(crc32c(0x59414d41u, A, (size_t)K * M), B, (size_t)L * N);
    const uint32_t band = crc32c(crc32c(0x42414e44u, LB, (size_t)(M + 1) * sizeof(int)), RB, (size_t)(M + 1) * sizeof(int));
    p.chk = ((uint64_t)text << 32) | band;
    return p;
}

// score tables as the child saw them when it shipped its first job (the parent has not run the tool's
// main() -- hence init_scores70/85 -- when it aligns the first batch)
std::vector<int32_t> g_wireScores;     // 128*128 ss + 16 gop + gap_extend

// CUDA start-up (driver + context + module load, around a second) runs on its own thread while the first speculative
// pass reads and walks the inputs: the parent forks first (a child never touches CUDA), then starts this.
std::thread g_warm;
int g_createRc = YB_OK;

void create_ctx() {
    const double t0 = now_ms();
    std::vector<int> devs;
    if (const char *e = getenv("YB_DEVICES")) {
        for (const char *p = e; *p;) {
            devs.push_back((int)strtol(p, const_cast<char **>(&p), 10));
            while (*p == ',' || *p == ' ') ++p;
        }
    }
    g_createRc = yb_create(devs.empty() ? nullptr : devs.data(), (int)devs.size(), &G.ctx);
    G.create_ms = now_ms() - t0;
}

void warm_ctx() {
    if (G.ctx || g_warm.joinable()) return;
    g_warm = std::thread(create_ctx);
}

bool g_scoresStale = false;
void ensure_ctx() {
    static bool ready = false;
    if (ready && !g_scoresStale) return;
    g_scoresStale = false;
    if (g_warm.joinable()) g_warm.join();
    else if (!G.ctx) create_ctx();
    int rc = g_createRc;
    if (rc != YB_OK || !G.ctx) fatalf("yama_b200: no usable CUDA device (this yama has no CPU implementation)");
    if (!g_wireScores.empty()) {
        rc = yb_set_scores(G.ctx, g_wireScores.data(), g_wireScores.data() + 128 * 128, g_wireScores[128 * 128 + 16]);
    } else {
        if (!ss || !gop) fatalf("yama_b200: score tables not initialised (init_scores70/85 must run before yama)");
        std::vector<int32_t> flat(128 * 128);
        for (int c = 0; c < 128; ++c) memcpy(&flat[(size_t)c * 128], ss[c], 128 * sizeof(int));
        rc = yb_set_scores(G.ctx, flat.data(), gop, gap_extend);
    }
    if (rc != YB_OK) fatalf("yama_b200: %s", yb_last_error(G.ctx));
    ready = true;
}

// ---- the resident server (yama_b200d, yama_served.cpp) as the backend -----------------------------------------
// YB_SERVER=auto | 1 | <socket path>: batches and block scores go to a yama_b200d process over a unix socket instead
// of a CUDA context of our own -- no driver start-up, no buffer allocation in this process.  If nobody answers on
// the socket the drop-in starts the server itself (YB_SERVER_SPAWN=0 forbids that) and waits for it.
struct Remote {
    bool enabled = false;
    std::string path, err;
    int fd = -1;
    int devices = 0;
    std::vector<uint8_t> scripts;        // packed scripts of the last batch
    std::vector<ybwire::Job> wjobs;
    std::vector<ybwire::Res> wres;
    std::vector<uint8_t> tmp;
} R;

bool write_full(int fd, const void *src, size_t n) {
    const uint8_t *p = static_cast<const uint8_t *>(src);
    while (n) {
        ssize_t k = write(fd, p, n);
        if (k < 0) { if (errno == EINTR) continue; return false; }
        p += k; n -= (size_t)k;
    }
    return true;
}
bool read_full(int fd, void *dst, size_t n);

void remote_configure() {
    const char *e = getenv("YB_SERVER");
    if (!e || !*e || !strcmp(e, "0") || !strcmp(e, "off")) return;
    R.enabled = true;
    if (!strcmp(e, "auto") || !strcmp(e, "1")) {
        R.path = ybwire::default_socket();
        if (R.path.empty()) fatalf("yama_b200: YB_SERVER=auto needs $XDG_RUNTIME_DIR or a private /tmp/yama_b200-<uid>; name a socket path instead");
    } else R.path = e;
}

int remote_try_connect() {
    sockaddr_un addr{};
    addr.sun_family = AF_UNIX;
    if (R.path.size() >= sizeof addr.sun_path) fatalf("yama_b200: YB_SERVER socket path too long");
    strcpy(addr.sun_path, R.path.c_str());
    int fd = socket(AF_UNIX, SOCK_STREAM, 0);
    if (fd < 0) return -1;
    if (connect(fd, reinterpret_cast<sockaddr *>(&addr), sizeof addr) != 0) { close(fd); return -1; }
    if (!ybwire::peer_is_me(fd)) { close(fd); fatalf("yama_b200: the server on %s belongs to another user", R.path.c_str()); }
    return fd;
}

void remote_spawn() {
    if (const char *e = getenv("YB_SERVER_SPAWN")) if (!strcmp(e, "0")) return;
    std::string bin;
    if (const char *e = getenv("YB_SERVER_BIN")) bin = e;
    else {
        char self[4096];
        ssize_t n = readlink("/proc/self/exe", self, sizeof self - 1);
        if (n <= 0) return;
        self[n] = 0;
        bin = self;
        bin = bin.substr(0, bin.rfind('/') + 1) + "yama_b200d";
    }
    fflush(nullptr);
    pid_t pid = fork();
    if (pid != 0) { if (pid > 0) { int st; while (waitpid(pid, &st, 0) < 0 && errno == EINTR) {} } return; }
    if (fork() != 0) _exit(0);                    // grandchild: not ours to wait for
    setsid();
    const std::string dir = ybwire::private_dir();
    const std::string log = dir.empty() ? std::string("/dev/null") : dir + "/yama_b200d.log";
    int nul = open("/dev/null", O_RDWR), lg = open(log.c_str(), O_WRONLY | O_CREAT | O_APPEND | O_NOFOLLOW, 0600);
    if (nul >= 0) { dup2(nul, 0); dup2(nul, 1); }
    if (lg >= 0) dup2(lg, 2); else if (nul >= 0) dup2(nul, 2);
    for (int fd = 3; fd < 256; ++fd) close(fd);
    const char *idle = getenv("YB_SERVER_IDLE_S");
    if (idle) execl(bin.c_str(), "yama_b200d", "--socket", R.path.c_str(), "--idle", idle, (char *)nullptr);
    else execl(bin.c_str(), "yama_b200d", "--socket", R.path.c_str(), (char *)nullptr);
    _exit(127);
}

// first contact, as early as possible (while the speculative pass runs): start the server if nobody is there
void remote_begin() {
    if (!R.enabled || R.fd >= 0) return;
    R.fd = remote_try_connect();
    if (R.fd < 0) remote_spawn();
}

void remote_hello() {
    ybwire::Hello h{ybwire::MAGIC_HELLO, ybwire::VERSION};
    std::vector<int32_t> sc;
    if (!g_wireScores.empty()) sc = g_wireScores;
    else {
        if (!ss || !gop) fatalf("yama_b200: score tables not initialised (init_scores70/85 must run before yama)");
        sc.resize(ybwire::SCORE_INTS);
        for (int c = 0; c < 128; ++c) memcpy(&sc[(size_t)c * 128], ss[c], 128 * sizeof(int));
        memcpy(&sc[128 * 128], gop, 16 * sizeof(int));
        sc[128 * 128 + 16] = gap_extend;
    }
    if (!write_full(R.fd, &h, sizeof h) || !write_full(R.fd, sc.data(), sc.size() * 4))
        fatalf("yama_b200: lost the server at %s", R.path.c_str());
}

void remote_ensure() {
    static bool greeted = false;
    if (R.fd < 0) {
        double wait_s = 60;
        if (const char *e = getenv("YB_SERVER_WAIT_S")) wait_s = atof(e);
        const double t0 = now_ms();
        for (;;) {
            R.fd = remote_try_connect();
            if (R.fd >= 0 || now_ms() - t0 > wait_s * 1e3) break;
            usleep(20000);
        }
        if (R.fd < 0) fatalf("yama_b200: no server answers on %s (YB_SERVER)", R.path.c_str());
        greeted = false;
    }
    if (!greeted) { remote_hello(); greeted = true; }
}

// jobs whose inputs all lie inside [base, base+baseBytes) go up without a copy; a lone job is packed first
int remote_batch(int64_t n, const yb_job *jobs, yb_result *res, yb_stats *st, const uint8_t *base, size_t baseBytes) {
    remote_ensure();
    if (!base) {
        R.tmp.clear();
        std::vector<size_t> offs;
        auto add = [&](const void *p, size_t bytes) {
            size_t off = (R.tmp.size() + 15) & ~(size_t)15;
            R.tmp.resize(off + bytes);
            memcpy(R.tmp.data() + off, p, bytes);
            offs.push_back(off);
        };
        for (int64_t i = 0; i < n; ++i) {
            add(jobs[i].A, (size_t)jobs[i].K * jobs[i].M); add(jobs[i].B, (size_t)jobs[i].L * jobs[i].N);
            add(jobs[i].LB, (size_t)(jobs[i].M + 1) * 4); add(jobs[i].RB, (size_t)(jobs[i].M + 1) * 4);
        }
        R.wjobs.resize((size_t)n);
        for (int64_t i = 0; i < n; ++i)
            R.wjobs[(size_t)i] = ybwire::Job{jobs[i].K, jobs[i].M, jobs[i].L, jobs[i].N, offs[4 * i], offs[4 * i + 1], offs[4 * i + 2], offs[4 * i + 3]};
        base = R.tmp.data();
        baseBytes = R.tmp.size();
    } else {
        R.wjobs.resize((size_t)n);
        for (int64_t i = 0; i < n; ++i)
            R.wjobs[(size_t)i] = ybwire::Job{jobs[i].K, jobs[i].M, jobs[i].L, jobs[i].N, (uint64_t)(jobs[i].A - base), (uint64_t)(jobs[i].B - base),
                                             (uint64_t)(reinterpret_cast<const uint8_t *>(jobs[i].LB) - base),
                                             (uint64_t)(reinterpret_cast<const uint8_t *>(jobs[i].RB) - base)};
    }
    ybwire::BatchReq rq{ybwire::MAGIC_BATCH, 0, (uint64_t)n, (uint64_t)baseBytes};
    ybwire::BatchResp rp;
    R.wres.resize((size_t)n);
    if (!write_full(R.fd, &rq, sizeof rq) || !write_full(R.fd, R.wjobs.data(), R.wjobs.size() * sizeof(ybwire::Job)) ||
        !write_full(R.fd, base, baseBytes) || !read_full(R.fd, &rp, sizeof rp) || rp.magic != ybwire::MAGIC_RESP || rp.n != (uint64_t)n ||
        !read_full(R.fd, R.wres.data(), R.wres.size() * sizeof(ybwire::Res)))
        fatalf("yama_b200: lost the server at %s", R.path.c_str());
    R.scripts.resize((size_t)rp.scriptBytes + 1);
    R.err.resize(rp.errLen);
    if (!read_full(R.fd, R.scripts.data(), (size_t)rp.scriptBytes) || !read_full(R.fd, &R.err[0], rp.errLen))
        fatalf("yama_b200: lost the server at %s", R.path.c_str());
    for (int64_t i = 0; i < n; ++i) {
        memset(&res[i], 0, sizeof res[i]);
        res[i].status = R.wres[(size_t)i].status;
        res[i].m_new = R.wres[(size_t)i].m_new;
        res[i].script = R.scripts.data() + R.wres[(size_t)i].scriptOff;
    }
    memset(st, 0, sizeof *st);
    st->kernel_ms = rp.kernel_ms; st->total_ms = rp.total_ms; st->cells = rp.cells; st->pairs = n;
    R.devices = rp.devices;
    return rp.rc;
}

const char *backend_error() { return R.enabled ? R.err.c_str() : (G.ctx ? yb_last_error(G.ctx) : "device failure"); }

// one batch through whichever backend this process uses
int backend_batch(int64_t n, const yb_job *jobs, yb_result *res, yb_stats *st, const uint8_t *base = nullptr, size_t baseBytes = 0) {
    if (R.enabled) return remote_batch(n, jobs, res, st, base, baseBytes);
    ensure_ctx();
    return yb_run_batch(G.ctx, n, jobs, res, st);
}

bool is_last_dummy(uchar **B, int L, int N) {
    if (!G.lastDummy || B != G.lastDummy || L != G.lastDummyRows || N != G.lastDummyCols) return false;
    for (int i = 2; i <= N; ++i) if (B[i] != B[i - 1] + L) return false;           // (ours are contiguous)
    return crc32c(0x44554d59u, B[1], (size_t)L * N) == G.lastDummySum;
}

// allocate the reference's output shape: caller frees AL[1] and AL+1 (mz_yama.h:17-18)
uchar **alloc_al(int m_new, int W) {
    uchar **al = static_cast<uchar **>(malloc(sizeof(uchar *) * (size_t)(m_new > 0 ? m_new : 1)));
    uchar *buf = static_cast<uchar *>(malloc((size_t)(m_new > 0 ? m_new : 1) * (size_t)W));
    if (!al || !buf) fatalf("Ran out of memory trying to allocate %lu.", (unsigned long)((size_t)m_new * W));
    al -= 1;
    for (int i = 1; i <= m_new; ++i) al[i] = buf + (size_t)(i - 1) * W;
    if (m_new < 1) al[1] = buf;
    return al;
}

void emit(const yb_job &job, int m_new, const uint8_t *script, uchar ***OAL, int *OM) {
    uchar **al = alloc_al(m_new, job.K + job.L);
    yb_result r;
    memset(&r, 0, sizeof r);
    r.m_new = m_new;
    r.script = script;
    if (yb_assemble(&job, &r, al[1]) != YB_OK)
        fatalf("new_align: edit script does not consume both alignments (M=%d, N=%d, M_new=%d)", job.M, job.N, m_new);
    G.lastDummy = nullptr;          // a real alignment: its address may be one a freed placeholder had
    *OAL = al;
    *OM = m_new;
}

void fail_from_status(int status) {
    if (status == YB_ERR_TRACEBACK) fatalf("Error generating edit script.");
    fatalf("yama_b200: %s", backend_error());
}

void run_direct(const yb_job &job, uchar ***OAL, int *OM) {
    yb_result r;
    yb_stats st;
    std::lock_guard<std::mutex> backend(S.backendMu);       // (the script stays valid while we hold the backend)
    double t0 = now_ms();
    int rc = backend_batch(1, &job, &r, &st);
    G.gpu_ms += now_ms() - t0;
    G.kernel_ms += st.kernel_ms;
    G.cells += st.cells;
    ++G.direct;
    if (rc != YB_OK) fail_from_status(rc);
    emit(job, r.m_new, r.script, OAL, OM);
}

// ---- child side ---------------------------------------------------------------------------------
void flush_pipe() {
    size_t off = 0;
    while (off < G.wbuf.size()) {
        ssize_t n = write(G.pipe_w, G.wbuf.data() + off, G.wbuf.size() - off);
        if (n < 0) { if (errno == EINTR) continue; _exit(3); }
        off += (size_t)n;
    }
    G.wbuf.clear();
}
void put(const void *p, size_t n) {
    const uint8_t *s = static_cast<const uint8_t *>(p);
    G.wbuf.insert(G.wbuf.end(), s, s + n);
    if (G.wbuf.size() > (1u << 20)) flush_pipe();
}
struct WireHdr { uint32_t magic; int32_t K, M, L, N; Key key; uint64_t call; uint64_t chk; };   // call: the child's yama() call index; chk: Proof::chk
struct WireEnd { uint32_t magic; uint32_t pad; uint64_t tainted, shipped, hits; };
constexpr uint32_t MAGIC_JOB = 0x4a4f4231u, MAGIC_END = 0x454e4431u, MAGIC_SCORES = 0x53434f31u;

void ship(const Key &k, const Proof &pr, const yb_job &j) {
    if (G.nShipped == 0) {
        put(&MAGIC_SCORES, 4);
        for (int c = 0; c < 128; ++c) put(ss[c], 128 * sizeof(int));
        put(gop, 16 * sizeof(int));
        put(&gap_extend, sizeof(int));
    }
    WireHdr h{MAGIC_JOB, j.K, j.M, j.L, j.N, k, G.calls, pr.chk};
    put(&h, sizeof h);
    put(j.A, (size_t)j.K * j.M);
    put(j.B, (size_t)j.L * j.N);
    put(j.LB, (size_t)(j.M + 1) * 4);
    put(j.RB, (size_t)(j.M + 1) * 4);
}
void child_finish() {
    WireEnd e{MAGIC_END, 0, G.nTainted, G.nShipped, G.nHits};
    put(&e, sizeof e);
    flush_pipe();
    close(G.pipe_w);
}

// shape-valid placeholder: column i of A beside column i of B, the longer alignment's tail alone -- max(M,N)
// columns, every column of A and of B exactly once and in order (the host rescoring of a placeholder block,
// mafScoreRange, costs rows^2 * columns, so fewer columns is cheaper than "all of A, then all of B")
void emit_dummy(const yb_job &job, uchar ***OAL, int *OM) {
    const int m_new = job.M > job.N ? job.M : job.N, W = job.K + job.L;
    uchar **al = alloc_al(m_new, W);
    for (int i = 1; i <= m_new; ++i) {
        if (i <= job.M) memcpy(al[i], job.A + (size_t)(i - 1) * job.K, (size_t)job.K); else memset(al[i], '-', (size_t)job.K);
        if (i <= job.N) memcpy(al[i] + job.K, job.B + (size_t)(i - 1) * job.L, (size_t)job.L); else memset(al[i] + job.K, '-', (size_t)job.L);
    }
    G.lastDummy = al;
    G.lastDummyRows = W;
    G.lastDummyCols = m_new;
    G.lastDummySum = crc32c(0x44554d59u, al[1], (size_t)W * m_new);
    *OAL = al;
    *OM = m_new;
}

// the deferred pass's placeholder: like emit_dummy, residues replaced by the marker byte of score_dropin.c
constexpr uchar DEFER_MARK = 1;
void emit_placeholder(const yb_job &job, uchar ***OAL, int *OM) {
    const int m_new = job.M > job.N ? job.M : job.N, W = job.K + job.L;
    uchar **al = alloc_al(m_new, W);
    for (int i = 1; i <= m_new; ++i) {
        uchar *c = al[i];
        for (int k = 0; k < job.K; ++k) c[k] = (i <= job.M && job.A[(size_t)(i - 1) * job.K + k] != '-') ? DEFER_MARK : (uchar)'-';
        for (int l = 0; l < job.L; ++l) c[job.K + l] = (i <= job.N && job.B[(size_t)(i - 1) * job.L + l] != '-') ? DEFER_MARK : (uchar)'-';
    }
    G.lastDummy = al;
    G.lastDummyRows = W;
    G.lastDummyCols = m_new;
    G.lastDummySum = crc32c(0x44554d59u, al[1], (size_t)W * m_new);
    *OAL = al;
    *OM = m_new;
}

void defer_record(const Key &key, const yb_job &j) {
    DeferJob q;
    q.K = j.K; q.M = j.M; q.L = j.L; q.N = j.N; q.key = key;
    auto put = [&](const void *src, size_t bytes) {
        const size_t off = (D.arena.size() + 15) & ~(size_t)15;
        if (D.arena.capacity() < off + bytes) D.arena.reserve(std::max<size_t>(2 * D.arena.capacity(), (size_t)64 << 20));
        D.arena.resize(off);                                     // (padding only: the job's bytes are appended, not zero-filled first)
        const uint8_t *p = static_cast<const uint8_t *>(src);
        D.arena.insert(D.arena.end(), p, p + bytes);
        return off;
    };
    q.offA = put(j.A, (size_t)j.K * j.M);
    q.offB = put(j.B, (size_t)j.L * j.N);
    q.offLB = put(j.LB, (size_t)(j.M + 1) * 4);
    q.offRB = put(j.RB, (size_t)(j.M + 1) * 4);
    D.jobs.push_back(q);
    D.lastJob = (int64_t)D.jobs.size() - 1;
}

// stdout of the host goes to a memory file while the deferred pass runs
void defer_begin() {
    fflush(stdout);
    D.savedStdout = dup(1);
    D.memFd = memfd_create("yama_b200_stdout", 0);
    if (D.savedStdout < 0 || D.memFd < 0 || dup2(D.memFd, 1) < 0) fatalf("yama_b200: cannot capture stdout (%s)", strerror(errno));
    D.active = true;
}

// Scores of the blocks that still need one (mafScoreRange over the whole block): one yb_score_blocks() batch when this
// process owns a context, the host's own function on a few threads otherwise (behind the resident server).
void defer_score(std::vector<DeferBlock> &blocks, size_t upto) {
    std::vector<size_t> want;
    for (size_t k = 0; k < upto; ++k) if (yb_defer_wants_score(blocks[k].ali)) want.push_back(k);
    if (want.empty()) return;
    const double t0 = now_ms();
    bool done = false;
    if (!R.enabled && G.ctx) {
        size_t nrows = 0;
        for (size_t k : want) nrows += (size_t)yb_defer_rows(blocks[k].ali, nullptr, 0);
        std::vector<const unsigned char *> rows(nrows + 1);
        std::vector<yb_block> yb(want.size());
        std::vector<double> sc(want.size());
        size_t at = 0;
        for (size_t q = 0; q < want.size(); ++q) {
            void *a = blocks[want[q]].ali;
            const int n = yb_defer_rows(a, rows.data() + at, (int)(nrows - at));
            yb[q].nrows = n; yb[q].text_size = yb_defer_text_size(a); yb[q].start = 0; yb[q].size = yb[q].text_size;
            yb[q].rows = rows.data() + at;
            at += (size_t)n;
        }
        if (yb_score_blocks(G.ctx, (int64_t)yb.size(), yb.data(), sc.data(), nullptr) == YB_OK) {
            for (size_t q = 0; q < want.size(); ++q) { blocks[want[q]].score = sc[q]; blocks[want[q]].haveScore = true; }
            done = true;
        }
    }
    if (!done) {
        unsigned nt = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t)
            th.emplace_back([&, t] {
                for (size_t q = t; q < want.size(); q += nt) {
                    blocks[want[q]].score = yb_defer_host_score(blocks[want[q]].ali);
                    blocks[want[q]].haveScore = true;
                }
            });
        for (auto &x : th) x.join();
    }
    G.scoreCalls += want.size();
    G.score_ms += now_ms() - t0;
}

// Align everything the pass recorded, score what needs a score, and write every block the host handed to mafWrite -- to
// stdout between the bytes the host printed there itself, to out1 / out2 in the host's order.
void defer_finish() {
    if (!D.active || D.done) return;
    D.done = true;
    if (G.debug) fprintf(stderr, "yama_b200[debug]: deferred pass over: %zu jobs, %zu blocks, %zu streams closed by the host\n", D.jobs.size(), D.blocks.size(), D.closed.size());
    const double t0 = now_ms();
    fflush(stdout);
    struct stat sb;
    const off_t total = fstat(D.memFd, &sb) == 0 ? sb.st_size : 0;
    std::vector<char> text((size_t)(total > 0 ? total : 0));
    if (total > 0 && pread(D.memFd, text.data(), (size_t)total, 0) != (ssize_t)total) fatalf("yama_b200: cannot read the captured stdout");
    dup2(D.savedStdout, 1);                                   // fd 1 is the caller's again
    FILE *out = fdopen(D.savedStdout, "w");                   // ... and so is this stream (the host may have closed `stdout`)
    if (!out) fatalf("yama_b200: cannot reopen stdout");
    __fsetlocking(out, FSETLOCKING_BYCALLER);
    close(D.memFd);
    G.mode = REPLAY;
    D.active = false;
    // ---- one batch ------------------------------------------------------------------------------------------------------
    const size_t n = D.jobs.size();
    std::vector<yb_job> jobs(n);
    for (size_t i = 0; i < n; ++i) {
        const DeferJob &q = D.jobs[i];
        jobs[i].K = q.K; jobs[i].M = q.M; jobs[i].L = q.L; jobs[i].N = q.N;
        jobs[i].A = D.arena.data() + q.offA; jobs[i].B = D.arena.data() + q.offB;
        jobs[i].LB = reinterpret_cast<const int32_t *>(D.arena.data() + q.offLB);
        jobs[i].RB = reinterpret_cast<const int32_t *>(D.arena.data() + q.offRB);
    }
    // YB_DUMP_JOBS=<file>: the jobs of this invocation as yama() received them, for the benchmark's replay of a real merge
    // (bench.py --workload cfg2real): "YBJ1", n, then per job K M L N (int32) and the four byte streams A, B, LB, RB
    if (const char *path = getenv("YB_DUMP_JOBS")) {
        if (FILE *df = fopen(path, "wb")) {
            const uint64_t n64 = n;
            fwrite("YBJ1", 1, 4, df); fwrite(&n64, 8, 1, df);
            for (size_t i = 0; i < n; ++i) {
                const int32_t dims[4] = {jobs[i].K, jobs[i].M, jobs[i].L, jobs[i].N};
                fwrite(dims, 4, 4, df);
            }
            for (int k = 0; k < 4; ++k)
                for (size_t i = 0; i < n; ++i) {
                    const yb_job &j = jobs[i];
                    if (k == 0) fwrite(j.A, 1, (size_t)j.K * j.M, df);
                    else if (k == 1) fwrite(j.B, 1, (size_t)j.L * j.N, df);
                    else fwrite(k == 2 ? j.LB : j.RB, 4, (size_t)j.M + 1, df);
                }
            fclose(df);
        }
    }
    std::vector<yb_result> res(n);
    int rc = YB_OK;
    size_t bad = n;                                           // first job that did not align
    if (n > 0) {
        yb_stats st;
        const double tb = now_ms();
        rc = backend_batch((int64_t)n, jobs.data(), res.data(), &st, D.arena.data(), D.arena.size());
        G.gpu_ms += now_ms() - tb;
        G.kernel_ms += st.kernel_ms; G.cells += st.cells; G.jobs += (int64_t)n; ++G.batches;
        if (rc == YB_ERR_CUDA || rc == YB_ERR_SCORES || rc == YB_ERR_ARG) fatalf("yama_b200: %s", backend_error());
        if (rc != YB_OK) { bad = 0; while (bad < n && res[bad].status == YB_OK) ++bad; }
    }
    // the reference would have died inside the failing yama() call, after writing everything before it
    size_t upto = D.blocks.size();
    if (rc != YB_OK)
        for (size_t k = 0; k < D.blocks.size(); ++k) if (D.blocks[k].job >= (int64_t)bad) { upto = k; break; }
    // ---- the merged blocks get their text; then every block that needs a score gets one ---------------------------------------
    std::vector<uchar> al;
    for (size_t k = 0; k < upto; ++k) {
        DeferBlock &b = D.blocks[k];
        if (b.job < 0) continue;
        const yb_job &j = jobs[(size_t)b.job];
        const yb_result &r = res[(size_t)b.job];
        al.resize((size_t)r.m_new * (size_t)(j.K + j.L) + 1);
        if (yb_assemble(&j, &r, al.data()) != YB_OK || yb_defer_fill(b.ali, al.data(), r.m_new, j.K + j.L) != 0)
            fatalf("new_align: edit script does not consume both alignments (M=%d, N=%d, M_new=%d)", j.M, j.N, r.m_new);
    }
    if (G.debug) fprintf(stderr, "yama_b200[debug]: batch rc=%d, filled; scoring\n", rc);
    defer_score(D.blocks, upto);
    if (G.debug) fprintf(stderr, "yama_b200[debug]: scored; writing\n");
    // ---- out: stdout's blocks go between the bytes the host printed itself, the others to their files in order --------------
    long at = 0;
    const long end = D.stdoutClosedAt >= 0 ? std::min((long)text.size(), D.stdoutClosedAt) : (long)text.size();
    for (size_t k = 0; k < upto; ++k) {
        DeferBlock &b = D.blocks[k];
        FILE *f = b.f;
        if (f == stdout) {
            const long pos = std::min(b.pos, end);
            if (pos > at) { fwrite(text.data() + at, 1, (size_t)(pos - at), out); at = pos; }
            f = out;
        }
        yb_defer_write(f, b.ali, b.haveScore ? 1 : 0, b.score);
    }
    const long last = upto < D.blocks.size() ? std::min(end, std::max(at, D.blocks[upto].f == stdout ? D.blocks[upto].pos : at)) : end;
    if (last > at) fwrite(text.data() + at, 1, (size_t)(last - at), out);
    fflush(out);
    for (FILE *f : D.closed) if (f != stdout) fclose(f);
    D.emit_ms = now_ms() - t0;
    if (rc != YB_OK) fail_from_status(bad < n ? res[bad].status : rc);
}

// ---- parent side --------------------------------------------------------------------------------
bool read_full(int fd, void *dst, size_t n) {
    uint8_t *d = static_cast<uint8_t *>(dst);
    while (n) {
        ssize_t k = read(fd, d, n);
        if (k == 0) return false;
        if (k < 0) { if (errno == EINTR) continue; return false; }
        d += k; n -= (size_t)k;
    }
    return true;
}

// returns false if the stream ended without a trailer (child died: fatal() in the host, bad input, ...)
void align_pending();
bool drain_child(int fd, WireEnd &end) {
    for (;;) {
        uint32_t magic;
        if (!read_full(fd, &magic, 4)) return false;
        if (magic == MAGIC_END) {
            end.magic = magic;
            return read_full(fd, reinterpret_cast<uint8_t *>(&end) + 4, sizeof end - 4);
        }
        if (magic == MAGIC_SCORES) {
            std::vector<int32_t> sc(128 * 128 + 17);
            if (!read_full(fd, sc.data(), sc.size() * 4)) return false;
            if (!g_wireScores.empty() && sc != g_wireScores) {       // tables changed between passes: start over
                g_scoresStale = true;                       // ensure_ctx() uploads the new tables before the next batch
                G.table.clear();
                G.scriptChunks.clear();
            }
            g_wireScores.swap(sc);
            continue;
        }
        if (magic != MAGIC_JOB) return false;
        WireHdr h;
        h.magic = magic;
        if (!read_full(fd, reinterpret_cast<uint8_t *>(&h) + 4, sizeof h - 4)) return false;
        Pending p;
        p.key = h.key; p.K = h.K; p.M = h.M; p.L = h.L; p.N = h.N;
        p.proof.K = h.K; p.proof.M = h.M; p.proof.L = h.L; p.proof.N = h.N; p.proof.chk = h.chk;
        auto grab = [&](size_t bytes, size_t &off) {
            off = (G.arena.size() + 15) & ~(size_t)15;
            G.arena.resize(off + bytes);
            return read_full(fd, G.arena.data() + off, bytes);
        };
        if (!grab((size_t)h.K * h.M, p.offA) || !grab((size_t)h.L * h.N, p.offB) ||
            !grab((size_t)(h.M + 1) * 4, p.offLB) || !grab((size_t)(h.M + 1) * 4, p.offRB))
            return false;
        G.pending.push_back(p);
        if (S.active) {
            bool go;
            {
                std::lock_guard<std::mutex> g(S.mu);
                S.pendingKeys.insert(p.key);
                S.childCalls = h.call;
                go = G.arena.size() >= S.chunkBytes || (S.starved && G.pending.size() >= S.minStarved);
            }
            S.cv.notify_all();
            if (go) align_pending();
        }
    }
}

void align_pending() {
    if (G.pending.empty()) return;
    const size_t n = G.pending.size();
    std::vector<yb_job> jobs(n);
    for (size_t i = 0; i < n; ++i) {
        const Pending &p = G.pending[i];
        jobs[i].K = p.K; jobs[i].M = p.M; jobs[i].L = p.L; jobs[i].N = p.N;
        jobs[i].A = G.arena.data() + p.offA;
        jobs[i].B = G.arena.data() + p.offB;
        jobs[i].LB = reinterpret_cast<const int32_t *>(G.arena.data() + p.offLB);
        jobs[i].RB = reinterpret_cast<const int32_t *>(G.arena.data() + p.offRB);
    }
    std::vector<yb_result> res(n);
    yb_stats st;
    std::unique_lock<std::mutex> backend(S.backendMu);
    double t0 = now_ms();
    int rc = backend_batch((int64_t)n, jobs.data(), res.data(), &st, G.arena.data(), G.arena.size());
    G.gpu_ms += now_ms() - t0;
    G.kernel_ms += st.kernel_ms;
    G.cells += st.cells;
    G.jobs += (int64_t)n;
    ++G.batches;
    if (rc == YB_ERR_CUDA || rc == YB_ERR_SCORES || rc == YB_ERR_ARG) fatalf("yama_b200: %s", backend_error());
    // per-pair failures (band / limit / traceback) are not fatal here: the final pass meets the same job
    // as a miss and reports it at the point where the reference would
    size_t bytes = 0;
    for (size_t i = 0; i < n; ++i) if (res[i].status == YB_OK) bytes += (size_t)(res[i].m_new + 3) / 4;
    std::vector<uint8_t> chunk(bytes + 1);
    std::vector<Entry> entries(n);
    size_t off = 0;
    for (size_t i = 0; i < n; ++i) {
        if (res[i].status != YB_OK) continue;
        const size_t len = (size_t)(res[i].m_new + 3) / 4;                 // packed, 2 bits per op
        memcpy(chunk.data() + off, res[i].script, len);
        entries[i].m_new = res[i].m_new;
        entries[i].script = chunk.data() + off;
        entries[i].proof = G.pending[i].proof;
        off += len;
    }
    backend.unlock();
    {
        std::lock_guard<std::mutex> g(S.mu);
        G.scriptChunks.push_back(std::move(chunk));                        // (moving a vector keeps its buffer: pointers stay valid)
        for (size_t i = 0; i < n; ++i) {
            if (S.active) S.pendingKeys.erase(G.pending[i].key);
            if (res[i].status != YB_OK) {
                ++G.failed;
                if (G.debug) {
                    G.failedKeys.emplace(G.pending[i].key, res[i].status);
                    fprintf(stderr, "yama_b200[debug]: batch job %zu failed status=%d K=%d M=%d L=%d N=%d\n", i, res[i].status,
                            G.pending[i].K, G.pending[i].M, G.pending[i].L, G.pending[i].N);
                }
                continue;
            }
            G.table.emplace(G.pending[i].key, entries[i]);
        }
        S.starved = false;
    }
    S.cv.notify_all();
    G.pending.clear();
    G.arena.clear();
}

[[noreturn]] void run_record_child(int argc, char **argv, int pipe_w);

int run_batched(int argc, char **argv) {
    const int maxPasses = 4;
    for (int pass = 1; pass <= maxPasses; ++pass) {
        int fds[2];
        if (pipe(fds) != 0) break;
        const double tc = now_ms();
        fflush(nullptr);
        pid_t pid = fork();
        if (pid < 0) { close(fds[0]); close(fds[1]); break; }
        if (pid == 0) {
            close(fds[0]);
            G.mode = RECORD;
            G.pipe_w = fds[1];
            int nul = open("/dev/null", O_WRONLY);
            if (nul >= 0) { dup2(nul, 1); dup2(nul, 2); close(nul); }
            ref_tool_main(argc, argv);
            yb_host_exit(0);
        }
        close(fds[1]);
        if (R.enabled) remote_begin(); else warm_ctx();
        WireEnd end{};
        const bool clean = drain_child(fds[0], end);
        close(fds[0]);
        int status = 0;
        while (waitpid(pid, &status, 0) < 0 && errno == EINTR) {}
        ++G.passes;
        G.child_ms += now_ms() - tc;
        const bool any = !G.pending.empty();
        align_pending();
        if (!clean || !any || end.tainted == 0) break;
    }
    G.mode = REPLAY;
    const double tf = now_ms();
    const int rc = ref_tool_main(argc, argv);
    G.final_ms = now_ms() - tf;
    return rc;
}

// YB_DROPIN=defer: one pass with deferred output (see Defer above); v=0 runs one speculative pass for the first stage first
int run_deferred(int argc, char **argv, bool chained) {
    if (chained) {
        int fds[2];
        if (pipe(fds) != 0) return run_batched(argc, argv);
        const double tc = now_ms();
        fflush(nullptr);
        pid_t pid = fork();
        if (pid < 0) { close(fds[0]); close(fds[1]); return run_batched(argc, argv); }
        if (pid == 0) { close(fds[0]); run_record_child(argc, argv, fds[1]); }
        close(fds[1]);
        if (R.enabled) remote_begin(); else warm_ctx();
        WireEnd end{};
        drain_child(fds[0], end);
        close(fds[0]);
        int status = 0;
        while (waitpid(pid, &status, 0) < 0 && errno == EINTR) {}
        ++G.passes;
        G.child_ms += now_ms() - tc;
        align_pending();
    } else {
        if (R.enabled) remote_begin(); else warm_ctx();
    }
    D.chained = chained;
    G.mode = DEFER;
    defer_begin();
    const double tf = now_ms();
    const int rc = ref_tool_main(argc, argv);
    G.final_ms = now_ms() - tf;
    defer_finish();
    return rc;
}

// YB_DROPIN=stream: speculative children, the real pass beside them (see Stream above)
constexpr uint32_t MAGIC_TABLE = 0x54424c31u;

[[noreturn]] void run_record_child(int argc, char **argv, int pipe_w) {
    G.mode = RECORD;
    G.pipe_w = pipe_w;
    int nul = open("/dev/null", O_WRONLY);
    if (nul >= 0) { dup2(nul, 1); dup2(nul, 2); close(nul); }
    ref_tool_main(argc, argv);
    yb_host_exit(0);
}

// the spare: sleeps until the parent sends the stage-1 results (then it is the second speculative pass) or hangs up
[[noreturn]] void run_spare_child(int argc, char **argv, int ctl_r, int pipe_w) {
    uint32_t magic = 0;
    uint64_t n = 0;
    if (!read_full(ctl_r, &magic, 4) || magic != MAGIC_TABLE || !read_full(ctl_r, &n, 8)) _exit(0);
    std::vector<int32_t> sc(128 * 128 + 17);
    if (!read_full(ctl_r, sc.data(), sc.size() * 4)) _exit(0);
    g_wireScores.swap(sc);
    uint64_t bytes = 0;
    if (!read_full(ctl_r, &bytes, 8)) _exit(0);
    G.scriptChunks.emplace_back((size_t)bytes + 1);
    uint8_t *base = G.scriptChunks.back().data();
    size_t off = 0;
    for (uint64_t i = 0; i < n; ++i) {
        Key k;
        int32_t m_new;
        Proof pr;
        if (!read_full(ctl_r, &k, sizeof k) || !read_full(ctl_r, &pr, sizeof pr) || !read_full(ctl_r, &m_new, 4)) _exit(0);
        const size_t len = (size_t)(m_new + 3) / 4;
        if (off + len > bytes || !read_full(ctl_r, base + off, len)) _exit(0);
        Entry e;
        e.m_new = m_new;
        e.script = base + off;
        e.proof = pr;
        off += len;
        G.table.emplace(k, e);
    }
    close(ctl_r);
    run_record_child(argc, argv, pipe_w);
}

bool send_table(int fd) {
    std::lock_guard<std::mutex> g(S.mu);
    uint64_t n = G.table.size(), bytes = 0;
    for (auto &kv : G.table) bytes += (uint64_t)(kv.second.m_new + 3) / 4;
    if (g_wireScores.size() != 128 * 128 + 17) return false;
    bool ok = write_full(fd, &MAGIC_TABLE, 4) && write_full(fd, &n, 8) && write_full(fd, g_wireScores.data(), g_wireScores.size() * 4) &&
              write_full(fd, &bytes, 8);
    for (auto it = G.table.begin(); ok && it != G.table.end(); ++it)
        ok = write_full(fd, &it->first, sizeof(Key)) && write_full(fd, &it->second.proof, sizeof(Proof)) && write_full(fd, &it->second.m_new, 4) &&
             write_full(fd, it->second.script, (size_t)(it->second.m_new + 3) / 4);
    return ok;
}

int run_streamed(int argc, char **argv) {
    int p1[2], p2[2], ctl[2];
    if (pipe(p1) != 0) return run_batched(argc, argv);
    if (pipe(p2) != 0 || pipe(ctl) != 0) { close(p1[0]); close(p1[1]); return run_batched(argc, argv); }
    const double tc = now_ms();
    fflush(nullptr);
    pid_t spare = fork();                          // before any thread or CUDA state exists in this process
    if (spare == 0) {
        close(p1[0]); close(p1[1]); close(p2[0]); close(ctl[1]);
        run_spare_child(argc, argv, ctl[0], p2[1]);
    }
    pid_t first = spare < 0 ? -1 : fork();
    if (first == 0) {
        close(p1[0]); close(p2[0]); close(p2[1]); close(ctl[0]); close(ctl[1]);
        run_record_child(argc, argv, p1[1]);
    }
    close(p1[1]); close(p2[1]); close(ctl[0]);
    if (spare < 0 || first < 0) {                  // could not fork: the classic way
        close(p1[0]); close(p2[0]); close(ctl[1]);
        int st;
        if (spare > 0) while (waitpid(spare, &st, 0) < 0 && errno == EINTR) {}
        return run_batched(argc, argv);
    }
    signal(SIGPIPE, SIG_IGN);
    if (R.enabled) remote_begin(); else warm_ctx();
    S.active = true;
    if (const char *e = getenv("YB_STREAM_CHUNK_MB")) S.chunkBytes = (size_t)std::max(1, atoi(e)) << 20;
    if (const char *e = getenv("YB_STREAM_MIN_JOBS")) S.minStarved = (size_t)std::max(1, atoi(e));
    S.reader = std::thread([=] {
        int status = 0;
        WireEnd end{};
        const bool clean = drain_child(p1[0], end);
        close(p1[0]);
        align_pending();                                   // what is left of pass 1
        while (waitpid(first, &status, 0) < 0 && errno == EINTR) {}
        ++G.passes;
        bool second = clean && end.tainted > 0;
        if (second) {
            { std::lock_guard<std::mutex> g(S.mu); S.childCalls = 0; }
            second = send_table(ctl[1]);
        }
        close(ctl[1]);                                     // (without a table the spare just leaves)
        {
            std::lock_guard<std::mutex> g(S.mu);
            S.lastPass = true;
        }
        S.cv.notify_all();
        if (second) {
            WireEnd end2{};
            drain_child(p2[0], end2);
            align_pending();
            ++G.passes;
        }
        close(p2[0]);
        while (waitpid(spare, &status, 0) < 0 && errno == EINTR) {}
        G.child_ms = now_ms() - tc;
        {
            std::lock_guard<std::mutex> g(S.mu);
            S.readerDone = true;
        }
        S.cv.notify_all();
    });
    G.mode = REPLAY;
    const double tf = now_ms();
    const int rc = ref_tool_main(argc, argv);
    G.final_ms = now_ms() - tf;
    if (S.reader.joinable()) S.reader.join();
    return rc;
}

void print_stats();
// End of the process: everything the tool wrote is flushed, then _exit -- tearing the CUDA context down buffer by
// buffer (yb_destroy + the runtime's atexit handlers) costs a few hundred milliseconds that no caller needs; the
// driver reclaims the context with the process.  The reference registers no atexit handlers of its own.
[[noreturn]] void finish_process(int code) {
    defer_finish();                                // (the host called exit() inside the deferred pass: write what it had)
    if (S.reader.joinable()) {                     // (the host called exit() inside the streamed real pass)
        if (S.reader.get_id() == std::this_thread::get_id()) S.reader.detach();   // a fatal error on the reader itself
        else if (code == 0) S.reader.join();
        else S.reader.detach();                    // dying: do not wait for a speculative child
    }
    if (g_warm.joinable()) g_warm.join();
    fflush(nullptr);
    print_stats();
    fflush(nullptr);
    _exit(code);
}

void print_stats() {
    if (!G.stats) return;
    fprintf(stderr,
            "yama_b200: passes=%d batches=%lld jobs=%lld failed=%lld cells=%lld calls=%llu misses=%llu direct=%llu "
            "gpu_ms=%.2f kernel_ms=%.2f create_ms=%.0f speculative_ms=%.0f final_ms=%.0f emit_ms=%.0f wall_ms=%.0f hit_ms=%.0f score_calls=%llu score_ms=%.1f stream_waits=%llu devices=%d\n",
            G.passes, (long long)G.batches, (long long)G.jobs, (long long)G.failed, (long long)G.cells, (unsigned long long)G.calls,
            (unsigned long long)G.misses, (unsigned long long)G.direct, G.gpu_ms, G.kernel_ms, G.create_ms, G.child_ms, G.final_ms, D.emit_ms,
            now_ms() - G.t_start, G.hit_ms, (unsigned long long)G.scoreCalls, G.score_ms, (unsigned long long)S.waits, R.enabled ? R.devices : (G.ctx ? yb_device_count(G.ctx) : 0));
}

}  // namespace

extern "C" {

// fopen() of the reference objects (compiled with -Dfopen=yb_host_fopen; util.c:36 ckopen, multi_util.c:107, multiz.c:242-243).  The host
// reads and writes MAF a character at a time (fgetc/fputc, maf.c); once a process has ever started a thread -- ours
// does: the packing helpers, the CUDA runtime -- glibc takes the stream lock on every one of those calls, which
// made the real pass 1.7x slower than the same pass in a thread-free process.  Only one thread ever runs host
// code, so its streams need no locking.
FILE *yb_host_fopen(const char *path, const char *mode) {
    // A speculative pass must never touch the tool's real outputs: multiz.c:242-243 opens out1/out2 with "w", which
    // would truncate what the real pass (running beside it in streamed mode) has already written.  Its output goes
    // nowhere anyway (mafWrite is skipped, stdout is /dev/null).
    if (G.mode == RECORD && mode && strpbrk(mode, "wa+")) path = "/dev/null";
    FILE *f = fopen(path, mode);
    if (f) __fsetlocking(f, FSETLOCKING_BYCALLER);
    return f;
}

// exit() of the reference objects (compiled with -Dexit=yb_host_exit): a forked speculative pass must not
// run the parent's atexit handlers (the CUDA runtime's among them) and must flush its job pipe.
void yb_host_exit(int code) {
    if (G.mode == RECORD && G.pipe_w >= 0) {
        if (code == 0) child_finish(); else flush_pipe();
        fflush(nullptr);
        _exit(code);
    }
    finish_process(code);
}

// mafScoreRange (score_dropin.c): skipped in a speculative pass, the host's own function by default, the device
// with YB_SCORE=gpu
int yb_dropin_defer_active(void) { return D.active ? 1 : 0; }

// score_dropin.c hands over its copy of a block the host just "wrote" (deferred pass)
void yb_dropin_defer_block(FILE *f, void *ali_copy, int placeholder) {
    if (placeholder && D.lastJob < 0) fatalf("yama_b200: a placeholder block without a job");
    if (placeholder && f != stdout) fatalf("yama_b200: a merged block is written to a file other than stdout");
    D.blocks.push_back(DeferBlock{f, placeholder ? D.lastJob : (int64_t)-1, f == stdout ? (long)ftello(stdout) : 0L, ali_copy, 0.0, false});
}

// fclose() of the reference objects (-Dfclose=yb_host_fclose): in the deferred pass the host's output streams are kept
// open until their blocks have been written (multiz.c:288-291 closes out1 / out2 -- and stdout, when they are not given)
int yb_host_fclose(FILE *f) {
    if (D.active && f) {
        // (the reference loses what it prints to stdout after closing it -- the "##eof maf" line of multiz.c:292 when
        //  out1 / out2 are not given: so do we)
        if (f == stdout) { if (D.stdoutClosedAt < 0) D.stdoutClosedAt = (long)ftello(stdout); return 0; }
        for (const DeferBlock &b : D.blocks)
            if (b.f == f) {
                if (std::find(D.closed.begin(), D.closed.end(), f) == D.closed.end()) D.closed.push_back(f);
                return 0;
            }
    }
    return fclose(f);
}

int yb_dropin_score_mode(void) {
    if (G.mode == RECORD) return 1;
    if (G.scoreGpu < 0) {
        const char *e = getenv("YB_SCORE");
        G.scoreGpu = (e && strcmp(e, "gpu") == 0) ? 1 : 0;
    }
    return G.scoreGpu ? 2 : 0;
}

double yb_dropin_score(int nrows, const unsigned char *const *rows, int text_size, int start, int size) {
    double score = 0.0;
    std::lock_guard<std::mutex> backend(S.backendMu);
    const double t0 = now_ms();
    int rc;
    if (R.enabled) {
        remote_ensure();
        ybwire::ScoreReq rq{ybwire::MAGIC_SCORE, nrows, text_size, start, size, 0};
        ybwire::ScoreResp rp;
        bool ok = write_full(R.fd, &rq, sizeof rq);
        for (int j = 0; ok && j < nrows; ++j) ok = write_full(R.fd, rows[j], (size_t)text_size);
        ok = ok && read_full(R.fd, &rp, sizeof rp) && rp.magic == ybwire::MAGIC_RESP;
        if (ok) { R.err.resize(rp.errLen); ok = read_full(R.fd, &R.err[0], rp.errLen); }
        if (!ok) fatalf("yama_b200: lost the server at %s", R.path.c_str());
        rc = rp.rc;
        score = rp.score;
    } else {
        ensure_ctx();
        yb_block blk;
        blk.nrows = nrows; blk.text_size = text_size; blk.start = start; blk.size = size; blk.rows = rows;
        rc = yb_score_blocks(G.ctx, 1, &blk, &score, nullptr);
    }
    G.score_ms += now_ms() - t0;
    ++G.scoreCalls;
    if (rc != YB_OK) fatalf("%s", backend_error());          // (a bad range carries the reference's own message)
    return score;
}

void yama(uchar **A, int K, int M, uchar **B, int L, int N, int *LB, int *RB, uchar ***OAL, int *OM) {
    ++G.calls;
    const double tHit0 = (G.stats && (G.mode == REPLAY || G.mode == DEFER)) ? now_ms() : 0.0;
    std::vector<uint8_t> tmpA, tmpB;
    yb_job job;
    job.K = K; job.M = M; job.L = L; job.N = N;
    job.LB = LB; job.RB = RB;

    // v=0 stage 2 of a pair whose stage 1 we answered with a placeholder: its inputs are meaningless
    if (G.mode == DEFER && is_last_dummy(B, L, N))
        fatalf("yama_b200: a second-stage yama() call (v=0) reached the deferred pass with a placeholder as its input");
    if (G.mode == RECORD && is_last_dummy(B, L, N)) {
        job.A = contiguous(A, K, M, tmpA);
        job.B = contiguous(B, L, N, tmpB);
        ++G.nTainted;
        emit_dummy(job, OAL, OM);
        return;
    }

    // the reference validates first (mz_yama.c:58-71) and dies with these words
    char msg[256];
    if (M < 1 || N < 1 || K < 1 || L < 1) fatalf("yama_b200: empty alignment K=%d M=%d L=%d N=%d", K, M, L, N);
    if (yb_check_band(M, N, LB, RB, msg, sizeof msg) < 0) fatalf("%s", msg);
    job.A = contiguous(A, K, M, tmpA);
    job.B = contiguous(B, L, N, tmpB);

    if (G.mode == DIRECT) { run_direct(job, OAL, OM); return; }

    // The deferred pass of v=1 has no table to consult -- nothing ran before it -- and every job is simply recorded: the two
    // digests over the job's bytes (a third of the time this function takes there) are not computed
    if (G.mode == DEFER && !D.chained && G.table.empty()) {
        defer_record(Key{0, 0}, job);
        emit_placeholder(job, OAL, OM);
        if (G.stats) G.hit_ms += now_ms() - tHit0;
        return;
    }

    const Key key = key_of(K, M, L, N, job.A, job.B, LB, RB);
    const Proof proof = proof_of(K, M, L, N, job.A, job.B, LB, RB);
    bool collided = false;
    {
        std::unique_lock<std::mutex> lk(S.mu, std::defer_lock);
        if (S.active) lk.lock();                // (only the streamed real pass shares the table with another thread)
        for (;;) {
            auto it = G.table.find(key);
            if (it != G.table.end() && !(it->second.proof == proof)) {       // same key, another job: never trust it
                collided = true;
                ++G.collisions;
                break;
            }
            if (it != G.table.end()) {
                const Entry e = it->second;
                if (lk.owns_lock()) lk.unlock();
                ++G.nHits;
                if (G.mode == DEFER && D.chained) ++D.callInPair;
                emit(job, e.m_new, e.script, OAL, OM);
                if (G.stats && G.mode == REPLAY) G.hit_ms += now_ms() - tHit0;
                return;
            }
            // streamed: not answered yet?  Wait while the child may still ship it: it has not reached this call, or
            // the job sits in the reader's queue.  Once the child's call counter has passed ours without shipping it
            // (v=0 stage 2), or the child is gone, it is a miss.
            if (!S.active || G.mode != REPLAY || S.readerDone) break;
            if (S.lastPass && S.childCalls > G.calls && !S.pendingKeys.count(key)) break;
            S.starved = true;
            ++S.waits;
            S.cv.wait(lk);
        }
    }
    if (G.mode == RECORD) {
        if (!collided && G.shipped.emplace(key, 1).second) { ship(key, proof, job); ++G.nShipped; }
        emit_dummy(job, OAL, OM);
        return;
    }
    if (G.mode == DEFER) {
        // v=0: the first call of a pair must have been answered by the speculative pass (a table hit, above); if it was
        // not, it is aligned now, alone -- its answer is the second call's input
        if (D.chained && (D.callInPair++ & 1) == 0) { ++G.misses; run_direct(job, OAL, OM); return; }
        defer_record(key, job);
        emit_placeholder(job, OAL, OM);
        if (G.stats) G.hit_ms += now_ms() - tHit0;          // (deferred pass: the time inside yama())
        return;
    }
    ++G.misses;                 // REPLAY miss: speculation did not cover this call; still exact
    if (G.debug) {
        auto f = G.failedKeys.find(key);
        fprintf(stderr, "yama_b200[debug]: replay miss at call %llu K=%d M=%d L=%d N=%d batch_status=%d\n",
                (unsigned long long)G.calls, K, M, L, N, f == G.failedKeys.end() ? 1 : f->second);
    }
    run_direct(job, OAL, OM);
}

}  // extern "C"

// The tool's entry point.  YB_DROPIN=direct keeps the reference's one-pair-at-a-time behaviour (each
// yama() call is a synchronous GPU launch); the default batches through record/replay (YB_DROPIN=batch), streamed
// (YB_DROPIN=stream) when the resident server is the backend.
int main(int argc, char **argv) {
    if (!ref_tool_main) {
        fprintf(stderr, "yama_dropin: linked without a reference tool (ref_tool_main)\n");
        return 2;
    }
    G.t_start = now_ms();
    // The host allocates and frees an alignment block per yama() call; above glibc's 128 KB mmap threshold each of
    // them is an mmap/munmap pair, and every munmap of a process that holds a CUDA context runs the driver's
    // mmu-notifier.  Keep those blocks on the heap.  (YB_MALLOPT=0 leaves malloc alone.)
    if (const char *e = getenv("YB_MALLOPT"); !e || strcmp(e, "0") != 0) {
        mallopt(M_MMAP_THRESHOLD, 1 << 30);
        mallopt(M_TRIM_THRESHOLD, 1 << 30);
        mallopt(M_TOP_PAD, 64 << 20);
    }
    for (FILE *f : {stdin, stdout, stderr}) __fsetlocking(f, FSETLOCKING_BYCALLER);     // see yb_host_fopen
    remote_configure();
    G.stats = getenv("YB_DROPIN_STATS") != nullptr;
    G.debug = getenv("YB_DROPIN_DEBUG") != nullptr;
    const char *m = getenv("YB_DROPIN");
    // Record/replay runs the tool's main() more than once, so its inputs must be re-readable.  The reference also
    // takes /dev/stdin (maf.c:343), FIFOs and /dev/fd/N: with such an input the tool runs once, one pair per launch.
    bool rereadable = true;
    for (int i = 1, files = 0; i < argc && files < 2; ++i) {
        if (argv[i][0] && argv[i][1] == '=') continue;              // R= M= L= S= ... flags (multiz.c:205, multic.c:290)
        ++files;                                                    // file1 file2 are the first two positional arguments
        struct stat sb;
        if (stat(argv[i], &sb) == 0 && !S_ISREG(sb.st_mode)) rereadable = false;
    }
    int rc;
    // v (multiz.c:247, multic.c: the third positional argument): 0 chains two yama() calls per overlap
    int v = -1;
    for (int i = 1, pos = 0; i < argc; ++i) {
        if (argv[i][0] && argv[i][1] == '=') continue;
        if (++pos == 3) { v = atoi(argv[i]); break; }
    }
    // multic decides whether to print a merged block by its WIDTH (multic.c:100), which a placeholder does not have:
    // with a minimum width (M=) it keeps the two-pass mode
    bool widthMatters = false;
    if (!multiz) for (int i = 1; i < argc; ++i) if (argv[i][0] == 'M' && argv[i][1] == '=' && atoi(argv[i] + 2) > 1) widthMatters = true;
    if ((m && strcmp(m, "direct") == 0) || (!rereadable && v != 1)) {
        G.mode = DIRECT;
        rc = ref_tool_main(argc, argv);
    } else if ((!m || strcmp(m, "defer") == 0) && (v == 0 || v == 1) && !widthMatters && (rereadable || v == 1)) {
        rc = run_deferred(argc, argv, v == 0);  // one pass of the host (v=1; inputs may be streams), two for v=0
    } else if (!rereadable) {
        G.mode = DIRECT;
        rc = ref_tool_main(argc, argv);
    } else if ((m && strcmp(m, "stream") == 0) || (R.enabled && !(m && strcmp(m, "batch") == 0))) {
        rc = run_streamed(argc, argv);          // the real pass beside the speculative one (behind the resident server)
    } else {
        rc = run_batched(argc, argv);
    }
    finish_process(rc);
}
