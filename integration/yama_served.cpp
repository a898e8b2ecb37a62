// yama_served.cpp -- yama_b200d: a resident owner of the GPU context for the drop-in.
//
// A multiz invocation through the drop-in spends 0.5-2 s starting the CUDA driver and creating a context, then
// allocating pinned staging and device buffers -- often more than all of its host work.  tba and roast run multiz
// once per tree node.  This process pays that once: it creates the yb context, listens on a unix socket and answers
// the drop-in's batches (protocol: yb_wire.h) with yb_run_batch / yb_score_blocks of libyama_b200.so.  It has no
// alignment logic of its own; any number of clients stay connected, their requests are answered one at a time in
// arrival order (each client's score tables are installed before its request if they differ); exits after --idle
// seconds without a client.  Started by hand, or by the drop-in itself when YB_SERVER names a socket nobody answers on.
//
//   yama_b200d [--socket PATH] [--idle SECONDS]
#include "../include/yama_b200.h"
#include "yb_wire.h"

#include <cerrno>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <fcntl.h>
#include <poll.h>
#include <sys/file.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

using namespace ybwire;

namespace {

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
bool write_full(int fd, const void *src, size_t n) {
    const uint8_t *s = static_cast<const uint8_t *>(src);
    while (n) {
        ssize_t k = write(fd, s, n);
        if (k < 0) { if (errno == EINTR) continue; return false; }
        s += k; n -= (size_t)k;
    }
    return true;
}

struct Server {
    yb_ctx *ctx = nullptr;
    std::vector<int32_t> scores;          // the tables the context currently holds
    std::vector<uint8_t> arena, scripts;
    std::vector<Job> wjobs;
    std::vector<yb_job> jobs;
    std::vector<yb_result> res;
    std::vector<Res> wres;
    std::vector<const uint8_t *> rows;
    uint64_t batches = 0, pairs = 0;

    // a client's greeting: its score tables (kept with the client, installed before each of its requests)
    bool hello(int fd, std::vector<int32_t> &sc) {
        Hello h;
        if (!read_full(fd, &h, sizeof h) || h.magic != MAGIC_HELLO || h.version != VERSION) return false;
        sc.resize(SCORE_INTS);
        return read_full(fd, sc.data(), sc.size() * 4);
    }
    void install(const std::vector<int32_t> &sc) {
        if (sc == scores) return;
        if (yb_set_scores(ctx, sc.data(), sc.data() + 128 * 128, sc[128 * 128 + 16]) != YB_OK) {
            fprintf(stderr, "yama_b200d: %s\n", yb_last_error(ctx));
            scores.clear();                  // the request will fail with YB_ERR_SCORES and carry the message
            return;
        }
        scores = sc;
    }

    bool batch(int fd, const BatchReq &rq) {
        if (rq.n > (1ull << 31) || rq.arenaBytes > (1ull << 40)) return false;
        wjobs.resize((size_t)rq.n);
        arena.resize((size_t)rq.arenaBytes + 64);
        if (!read_full(fd, wjobs.data(), wjobs.size() * sizeof(Job)) || !read_full(fd, arena.data(), (size_t)rq.arenaBytes)) return false;
        jobs.resize(wjobs.size());
        res.resize(wjobs.size());
        BatchResp rp{};
        rp.magic = MAGIC_RESP; rp.n = rq.n;
        bool ok = true;
        for (size_t i = 0; i < wjobs.size(); ++i) {
            const Job &w = wjobs[i];
            const uint64_t a = (uint64_t)(w.K > 0 ? w.K : 0) * (uint64_t)(w.M > 0 ? w.M : 0), b = (uint64_t)(w.L > 0 ? w.L : 0) * (uint64_t)(w.N > 0 ? w.N : 0);
            const uint64_t band = ((uint64_t)(w.M > 0 ? w.M : 0) + 1) * 4;
            auto inside = [&](uint64_t off, uint64_t len) { return off <= rq.arenaBytes && len <= rq.arenaBytes - off; };   // (no wrap)
            if (!inside(w.offA, a) || !inside(w.offB, b) || !inside(w.offLB, band) || !inside(w.offRB, band) ||
                (w.offLB & 3) || (w.offRB & 3)) { ok = false; break; }
            jobs[i].K = w.K; jobs[i].M = w.M; jobs[i].L = w.L; jobs[i].N = w.N;
            jobs[i].A = arena.data() + w.offA; jobs[i].B = arena.data() + w.offB;
            jobs[i].LB = reinterpret_cast<const int32_t *>(arena.data() + w.offLB);
            jobs[i].RB = reinterpret_cast<const int32_t *>(arena.data() + w.offRB);
        }
        std::string err;
        wres.assign(wjobs.size(), Res{YB_ERR_ARG, 0, 0});
        scripts.clear();
        if (!ok) { rp.rc = YB_ERR_ARG; err = "yama_b200d: malformed batch (offsets outside the arena)"; }
        else {
            yb_stats st;
            memset(&st, 0, sizeof st);
            rp.rc = yb_run_batch(ctx, (int64_t)jobs.size(), jobs.data(), res.data(), &st);
            rp.kernel_ms = st.kernel_ms; rp.total_ms = st.total_ms; rp.cells = st.cells;
            if (rp.rc != YB_OK) err = yb_last_error(ctx);
            if (rp.rc != YB_ERR_CUDA && rp.rc != YB_ERR_SCORES && rp.rc != YB_ERR_ARG)
                for (size_t i = 0; i < jobs.size(); ++i) {
                    wres[i].status = res[i].status; wres[i].m_new = res[i].m_new; wres[i].scriptOff = scripts.size();
                    if (res[i].status == YB_OK && res[i].script)
                        scripts.insert(scripts.end(), res[i].script, res[i].script + (res[i].m_new + 3) / 4);
                }
        }
        rp.scriptBytes = scripts.size();
        rp.errLen = (uint32_t)err.size();
        rp.devices = yb_device_count(ctx);
        ++batches; pairs += rq.n;
        return write_full(fd, &rp, sizeof rp) && write_full(fd, wres.data(), wres.size() * sizeof(Res)) &&
               write_full(fd, scripts.data(), scripts.size()) && write_full(fd, err.data(), err.size());
    }

    bool score(int fd, const ScoreReq &rq) {
        if (rq.nrows < 0 || rq.text_size < 0 || (uint64_t)rq.nrows * (uint64_t)rq.text_size > (1ull << 36)) return false;
        arena.resize((size_t)rq.nrows * (size_t)rq.text_size + 64);
        if (!read_full(fd, arena.data(), (size_t)rq.nrows * (size_t)rq.text_size)) return false;
        rows.resize((size_t)rq.nrows);
        for (int j = 0; j < rq.nrows; ++j) rows[(size_t)j] = arena.data() + (size_t)j * (size_t)rq.text_size;
        yb_block blk;
        blk.nrows = rq.nrows; blk.text_size = rq.text_size; blk.start = rq.start; blk.size = rq.size; blk.rows = rows.data();
        ScoreResp rp{};
        rp.magic = MAGIC_RESP;
        rp.rc = yb_score_blocks(ctx, 1, &blk, &rp.score, nullptr);
        std::string err = rp.rc == YB_OK ? "" : yb_last_error(ctx);
        rp.errLen = (uint32_t)err.size();
        return write_full(fd, &rp, sizeof rp) && write_full(fd, err.data(), err.size());
    }

    // one request of a greeted client; false: the client is gone (or sent garbage) and is dropped
    bool request(int fd, const std::vector<int32_t> &sc) {
        uint32_t magic;
        if (!read_full(fd, &magic, 4)) return false;
        install(sc);
        if (magic == MAGIC_BATCH) {
            BatchReq rq;
            rq.magic = magic;
            return read_full(fd, reinterpret_cast<uint8_t *>(&rq) + 4, sizeof rq - 4) && batch(fd, rq);
        }
        if (magic == MAGIC_SCORE) {
            ScoreReq rq;
            rq.magic = magic;
            return read_full(fd, reinterpret_cast<uint8_t *>(&rq) + 4, sizeof rq - 4) && score(fd, rq);
        }
        return false;
    }
};

std::string g_sock;
void on_signal(int) {
    if (!g_sock.empty()) unlink(g_sock.c_str());
    _exit(0);
}

}  // namespace

int main(int argc, char **argv) {
    std::string path = default_socket();
    int idle_s = 300;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--socket") && i + 1 < argc) path = argv[++i];
        else if (!strcmp(argv[i], "--idle") && i + 1 < argc) idle_s = atoi(argv[++i]);
        else { fprintf(stderr, "usage: yama_b200d [--socket PATH] [--idle SECONDS]\n"); return 2; }
    }
    if (path.empty()) { fprintf(stderr, "yama_b200d: no private directory for the default socket; give --socket PATH\n"); return 2; }
    sockaddr_un addr{};
    addr.sun_family = AF_UNIX;
    if (path.size() >= sizeof addr.sun_path) { fprintf(stderr, "yama_b200d: socket path too long\n"); return 2; }
    strcpy(addr.sun_path, path.c_str());
    {   // somebody already serving there?
        int probe = socket(AF_UNIX, SOCK_STREAM, 0);
        if (probe >= 0 && connect(probe, reinterpret_cast<sockaddr *>(&addr), sizeof addr) == 0) { close(probe); return 0; }
        if (probe >= 0) close(probe);
    }
    // one server per socket: whoever holds the lock file serves; a second one started at the same moment (several
    // clients found nobody at once) leaves, and its client finds the first
    {
        const std::string lock = path + ".lock";
        int lf = open(lock.c_str(), O_CREAT | O_RDWR | O_NOFOLLOW | O_CLOEXEC, 0600);
        if (lf < 0 || flock(lf, LOCK_EX | LOCK_NB) != 0) return 0;      // (kept open, hence locked, for our lifetime)
    }
    signal(SIGPIPE, SIG_IGN);
    Server S;
    if (yb_create(nullptr, 0, &S.ctx) != YB_OK) {          // (YB_DEVICES-style selection: CUDA_VISIBLE_DEVICES)
        fprintf(stderr, "yama_b200d: no usable CUDA device\n");
        return 1;
    }
    unlink(path.c_str());
    int ls = socket(AF_UNIX, SOCK_STREAM, 0);
    const mode_t old = umask(0077);                          // the socket is the owner's only
    if (ls < 0 || bind(ls, reinterpret_cast<sockaddr *>(&addr), sizeof addr) != 0 || listen(ls, 64) != 0) {
        fprintf(stderr, "yama_b200d: cannot listen on %s: %s\n", path.c_str(), strerror(errno));
        return 1;
    }
    umask(old);
    g_sock = path;
    signal(SIGTERM, on_signal);
    signal(SIGINT, on_signal);
    fprintf(stderr, "yama_b200d: serving %d device(s) on %s (idle limit %d s)\n", yb_device_count(S.ctx), path.c_str(), idle_s);
    struct Client { int fd; bool greeted; std::vector<int32_t> scores; };
    std::vector<Client> clients;
    std::vector<pollfd> fds;
    for (;;) {
        fds.assign(1, pollfd{ls, POLLIN, 0});
        for (auto &c : clients) fds.push_back(pollfd{c.fd, POLLIN, 0});
        int r = poll(fds.data(), fds.size(), (clients.empty() && idle_s > 0) ? idle_s * 1000 : -1);
        if (r < 0 && errno == EINTR) continue;
        if (r < 0) break;
        if (r == 0) { if (clients.empty()) break; else continue; }      // idle with nobody connected: leave
        if (fds[0].revents & POLLIN) {
            int fd = accept(ls, nullptr, nullptr);
            if (fd >= 0 && !peer_is_me(fd)) { close(fd); fd = -1; }          // only this user's processes are served
            if (fd >= 0) clients.push_back(Client{fd, false, {}});
        }
        for (size_t k = 1; k < fds.size(); ++k) {
            if (!(fds[k].revents & (POLLIN | POLLHUP | POLLERR))) continue;
            Client &c = clients[k - 1];
            bool ok;
            if (!c.greeted) { ok = S.hello(c.fd, c.scores); c.greeted = ok; }
            else ok = S.request(c.fd, c.scores);
            if (!ok) { close(c.fd); c.fd = -1; }
        }
        for (size_t k = clients.size(); k-- > 0;)
            if (clients[k].fd < 0) clients.erase(clients.begin() + (long)k);
    }
    fprintf(stderr, "yama_b200d: idle, leaving after %llu batches / %llu pairs\n", (unsigned long long)S.batches, (unsigned long long)S.pairs);
    unlink(path.c_str());
    yb_destroy(S.ctx);
    return 0;
}
