// yb_wire.h -- the byte protocol between the drop-in (yama_dropin.cpp, client) and yama_b200d (yama_served.cpp),
// the resident process that keeps the CUDA context, the streams and the staging buffers of libyama_b200.so warm
// across the many short multiz / multic invocations of a tba or roast run.  Same machine, same ABI: plain structs
// over a unix stream socket, no byte swapping.  The reference has no counterpart (it is a single process on a CPU).
//
//   client                                   server
//   Hello + score tables (128*128+17 ints) ->
//   BatchReq + n Job + arena bytes          ->  yb_run_batch
//                                           <-  BatchResp + n Res + packed scripts + error text
//   ScoreReq + nrows*text_size bytes        ->  yb_score_blocks
//                                           <-  ScoreResp + error text
//   (close)
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

#include <sys/socket.h>
#include <sys/stat.h>
#include <unistd.h>

namespace ybwire {

constexpr uint32_t MAGIC_HELLO = 0x59425331u;   // "YBS1"
constexpr uint32_t MAGIC_BATCH = 0x59424254u;   // "YBBT"
constexpr uint32_t MAGIC_SCORE = 0x59425343u;   // "YBSC"
constexpr uint32_t MAGIC_RESP = 0x59425250u;    // "YBRP"
constexpr uint32_t VERSION = 1;
constexpr int SCORE_INTS = 128 * 128 + 16 + 1;  // ss, gop, gap_extend

struct Hello { uint32_t magic, version; };
struct BatchReq { uint32_t magic, pad; uint64_t n, arenaBytes; };
struct Job { int32_t K, M, L, N; uint64_t offA, offB, offLB, offRB; };     // offsets into the arena that follows
struct BatchResp {
    uint32_t magic; int32_t rc;
    uint64_t n, scriptBytes;
    double kernel_ms, total_ms;
    int64_t cells;
    uint32_t errLen; int32_t devices;
};
struct Res { int32_t status, m_new; uint64_t scriptOff; };                 // scriptOff into the script bytes that follow
struct ScoreReq { uint32_t magic; int32_t nrows, text_size, start, size, pad; };
struct ScoreResp { uint32_t magic; int32_t rc; double score; uint32_t errLen, pad; };

// Where the default socket, its lock and the server log live: a directory only this user can enter --
// $XDG_RUNTIME_DIR when it is one, else /tmp/yama_b200-<uid> (created 0700; refused if it is a symlink, somebody
// else's, or open to others).  Empty on failure: the caller then needs an explicit socket path.
inline std::string private_dir() {
    auto mine = [](const char *d) {
        struct stat sb;
        return lstat(d, &sb) == 0 && S_ISDIR(sb.st_mode) && sb.st_uid == getuid() && (sb.st_mode & 077) == 0;
    };
    if (const char *x = getenv("XDG_RUNTIME_DIR")) if (*x && mine(x)) return x;
    char b[128];
    snprintf(b, sizeof b, "/tmp/yama_b200-%u", (unsigned)getuid());
    mkdir(b, 0700);                                  // (fails if it exists: judged by lstat either way)
    return mine(b) ? std::string(b) : std::string();
}
inline std::string default_socket() {
    const std::string d = private_dir();
    return d.empty() ? std::string() : d + "/yama_b200.sock";
}
// the process at the other end of a connected unix socket runs as this user
inline bool peer_is_me(int fd) {
    struct ucred cr;
    socklen_t len = sizeof cr;
    return getsockopt(fd, SOL_SOCKET, SO_PEERCRED, &cr, &len) == 0 && cr.uid == getuid();
}

}  // namespace ybwire
