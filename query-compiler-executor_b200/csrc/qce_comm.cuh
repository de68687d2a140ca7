// qce_comm.cuh -- the ranks of one node, without torch / MPI / NCCL.
//
// The engine is SPMD: one process per GPU runs the same host operator layer
// (src/*.c) over its share of the data; everything the ranks must agree on is a
// handful of small host vectors per operator (histograms, counts, checksums).
// Those travel through one POSIX shared-memory segment: a sense-reversing
// barrier and per-rank slots that make all-gather a memcpy.  The bulk data never
// passes through here -- it moves GPU to GPU over NVLink, stored by the pushing
// kernels straight into the peers' CUDA-IPC-mapped windows (k_exchange.cuh).
//
// Two ways in:
//   fork   qce_comm_fork(n): called by the C host layer before any CUDA call when
//          QCE_GPUS=n; the parent becomes rank 0, the children ranks 1..n-1 (this is how
//          the unchanged reference main runs sharded);
//   named  qce_comm_attach(name, rank, world): independent processes (torchrun's
//          ranks in bench.py) meet in /dev/shm/<name>.
//
// Every wait is bounded (QCE_COMM_TIMEOUT_S, default 120 s) and every rank that
// fails raises a shared abort flag, so one rank's error ends all of them instead
// of leaving the others spinning on a barrier.
#pragma once
#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/prctl.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <string>
#include <vector>

namespace qcecomm {

constexpr uint32_t kMaxRanks = 16;
constexpr size_t kSlotBytes = 1u << 20; // per-rank all-gather slot (histograms are 2-16 KB)
constexpr uint32_t kMagic = 0x51434531u;

struct alignas(64) Header {
    std::atomic<uint32_t> magic;
    std::atomic<uint32_t> world;
    std::atomic<uint64_t> token;
    alignas(64) std::atomic<uint32_t> arrive;
    alignas(64) std::atomic<uint32_t> gen;
    alignas(64) std::atomic<uint32_t> abort_flag;
    alignas(64) std::atomic<uint32_t> attached;
};

struct State {
    Header *h = nullptr;
    unsigned char *slots = nullptr;
    size_t map_bytes = 0;
    uint32_t rank = 0, world = 1;
    bool forked_child = false;
    std::vector<pid_t> children;
    std::string shm_name; // named segments are unlinked by rank 0
    double timeout_s = 120.0;
    uint64_t n_barriers = 0, n_allgathers = 0;
    char err[256] = "";
};
inline State &st()
{
    static State s;
    return s;
}

inline double now_s()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
inline size_t segment_bytes(uint32_t world) { return 4096 + (size_t)world * kSlotBytes; }

inline int set_err(const char *msg)
{
    snprintf(st().err, sizeof st().err, "%s", msg);
    if (st().h) st().h->abort_flag.store(1, std::memory_order_release);
    return -1;
}
inline void init_header(Header *h, uint32_t world, uint64_t token)
{
    h->world.store(world);
    h->token.store(token);
    h->arrive.store(0);
    h->gen.store(0);
    h->abort_flag.store(0);
    h->attached.store(0);
    h->magic.store(kMagic, std::memory_order_release);
}
inline void read_env()
{
    const char *t = getenv("QCE_COMM_TIMEOUT_S");
    if (t && atof(t) > 0) st().timeout_s = atof(t);
}

// Raise the abort flag: every peer's next (or current) wait fails.
inline void abort_all()
{
    if (st().h) st().h->abort_flag.store(1, std::memory_order_release);
}

inline int barrier()
{
    State &s = st();
    if (s.world <= 1) return 0;
    Header *h = s.h;
    s.n_barriers++;
    if (h->abort_flag.load(std::memory_order_acquire)) return set_err("a peer rank failed (abort flag raised)");
    const uint32_t g = h->gen.load(std::memory_order_acquire);
    if (h->arrive.fetch_add(1, std::memory_order_acq_rel) + 1 == s.world) {
        h->arrive.store(0, std::memory_order_relaxed);
        h->gen.store(g + 1, std::memory_order_release);
        return 0;
    }
    uint64_t spins = 0;
    double t0 = 0;
    while (h->gen.load(std::memory_order_acquire) == g) {
        if (h->abort_flag.load(std::memory_order_acquire)) return set_err("a peer rank failed (abort flag raised)");
        if (++spins < 2000) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
            continue;
        }
        sched_yield();
        if ((spins & 1023) == 0) {
            const double t = now_s();
            if (t0 == 0) t0 = t;
            else if (t - t0 > s.timeout_s) return set_err("barrier timed out: a peer rank did not arrive");
        }
    }
    return 0;
}

// out[r * bytes ...] = rank r's `mine`; bytes <= kSlotBytes per round (larger payloads loop).
inline int allgather(const void *mine, size_t bytes, void *out)
{
    State &s = st();
    if (s.world <= 1) {
        if (out != mine) memcpy(out, mine, bytes);
        return 0;
    }
    s.n_allgathers++;
    for (size_t at = 0; at < bytes || (bytes == 0 && at == 0); at += kSlotBytes) {
        const size_t len = bytes - at < kSlotBytes ? bytes - at : kSlotBytes;
        memcpy(s.slots + (size_t)s.rank * kSlotBytes, (const unsigned char *)mine + at, len);
        if (barrier() != 0) return -1;
        for (uint32_t r = 0; r < s.world; r++)
            memcpy((unsigned char *)out + (size_t)r * bytes + at, s.slots + (size_t)r * kSlotBytes, len);
        if (barrier() != 0) return -1;
        if (bytes == 0) break;
    }
    return 0;
}
inline int allreduce_sum(uint64_t *v, size_t n)
{
    State &s = st();
    if (s.world <= 1) return 0;
    std::vector<uint64_t> all((size_t)s.world * n);
    if (allgather(v, n * sizeof(uint64_t), all.data()) != 0) return -1;
    for (size_t i = 0; i < n; i++) {
        uint64_t a = 0;
        for (uint32_t r = 0; r < s.world; r++) a += all[(size_t)r * n + i];
        v[i] = a;
    }
    return 0;
}
inline int allreduce_max(uint64_t *v, size_t n)
{
    State &s = st();
    if (s.world <= 1) return 0;
    std::vector<uint64_t> all((size_t)s.world * n);
    if (allgather(v, n * sizeof(uint64_t), all.data()) != 0) return -1;
    for (size_t i = 0; i < n; i++) {
        uint64_t a = 0;
        for (uint32_t r = 0; r < s.world; r++) a = all[(size_t)r * n + i] > a ? all[(size_t)r * n + i] : a;
        v[i] = a;
    }
    return 0;
}
// Variable-length blobs to rank 0 (per-query stdout of the queries a rank ran alone):
// lens all-gathered, then the payloads in slot-sized rounds.  out[r] valid on rank 0 only.
inline int gatherv_root(const void *mine, size_t bytes, std::vector<std::string> *out)
{
    State &s = st();
    std::vector<uint64_t> lens(s.world);
    uint64_t me = bytes;
    if (allgather(&me, sizeof me, lens.data()) != 0) return -1;
    uint64_t longest = 0;
    for (auto l : lens) longest = l > longest ? l : longest;
    if (out) { out->assign(s.world, std::string()); }
    if (s.world <= 1) {
        if (out) (*out)[0].assign((const char *)mine, bytes);
        return 0;
    }
    for (uint64_t at = 0; at < longest; at += kSlotBytes) {
        if (at < bytes) {
            const size_t len = bytes - at < kSlotBytes ? bytes - at : kSlotBytes;
            memcpy(s.slots + (size_t)s.rank * kSlotBytes, (const unsigned char *)mine + at, len);
        }
        if (barrier() != 0) return -1;
        if (s.rank == 0 && out)
            for (uint32_t r = 0; r < s.world; r++)
                if (at < lens[r]) {
                    const size_t len = lens[r] - at < kSlotBytes ? lens[r] - at : kSlotBytes;
                    (*out)[r].append((const char *)(s.slots + (size_t)r * kSlotBytes), len);
                }
        if (barrier() != 0) return -1;
    }
    return 0;
}

inline int map_segment(int fd, uint32_t world)
{
    State &s = st();
    s.map_bytes = segment_bytes(world);
    void *p = mmap(NULL, s.map_bytes, PROT_READ | PROT_WRITE, MAP_SHARED | (fd < 0 ? MAP_ANONYMOUS : 0), fd, 0);
    if (p == MAP_FAILED) return set_err("mmap of the rank segment failed");
    s.h = (Header *)p;
    s.slots = (unsigned char *)p + 4096;
    return 0;
}

// fork mode: parent = rank 0.  Must run before the process touches CUDA.
inline int fork_ranks(uint32_t world)
{
    State &s = st();
    if (s.h) return set_err("ranks already exist");
    if (world < 1 || world > kMaxRanks) return set_err("world size must be 1..16");
    read_env();
    s.world = world;
    s.rank = 0;
    if (world == 1) return 0;
    if (map_segment(-1, world) != 0) return -1;
    init_header(s.h, world, (uint64_t)getpid());
    fflush(stdout);
    fflush(stderr);
    for (uint32_t r = 1; r < world; r++) {
        pid_t p = fork();
        if (p < 0) {
            abort_all();
            return set_err("fork failed");
        }
        if (p == 0) {
            prctl(PR_SET_PDEATHSIG, SIGKILL);
            s.rank = r;
            s.forked_child = true;
            s.children.clear();
            return 0;
        }
        s.children.push_back(p);
    }
    return 0;
}
// rank 0 of fork mode: collect the children; returns the number that failed
inline int join_children()
{
    State &s = st();
    int bad = 0;
    for (pid_t p : s.children) {
        int status = 0;
        if (waitpid(p, &status, 0) < 0 || !WIFEXITED(status) || WEXITSTATUS(status) != 0) bad++;
    }
    s.children.clear();
    return bad;
}

inline int attach_named(const char *name, uint32_t rank, uint32_t world, uint64_t token)
{
    State &s = st();
    if (s.h) return set_err("ranks already exist");
    if (world < 1 || world > kMaxRanks || rank >= world) return set_err("bad rank / world size");
    read_env();
    s.world = world;
    s.rank = rank;
    if (world == 1) return 0;
    std::string nm = std::string(name[0] == '/' ? "" : "/") + name;
    int fd = -1;
    if (rank == 0) {
        shm_unlink(nm.c_str());
        fd = shm_open(nm.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0) return set_err("shm_open (create) failed");
        if (ftruncate(fd, (off_t)segment_bytes(world)) != 0) { close(fd); return set_err("ftruncate failed"); }
        if (map_segment(fd, world) != 0) { close(fd); return -1; }
        close(fd);
        init_header(s.h, world, token);
        s.shm_name = nm;
    } else {
        const double t0 = now_s();
        for (;;) {
            fd = shm_open(nm.c_str(), O_RDWR, 0600);
            if (fd >= 0) {
                struct stat sb;
                if (fstat(fd, &sb) == 0 && (size_t)sb.st_size >= segment_bytes(world)) {
                    if (map_segment(fd, world) != 0) { close(fd); return -1; }
                    close(fd);
                    // a stale segment of an earlier run carries another token: wait for rank 0 to replace it
                    if (s.h->magic.load(std::memory_order_acquire) == kMagic && s.h->token.load() == token &&
                        s.h->world.load() == world)
                        break;
                    munmap((void *)s.h, s.map_bytes);
                    s.h = nullptr;
                } else
                    close(fd);
            }
            if (now_s() - t0 > s.timeout_s) return set_err("rank 0's segment did not appear");
            usleep(1000);
        }
    }
    s.h->attached.fetch_add(1);
    const double t0 = now_s();
    while (s.h->attached.load() < world) {
        if (now_s() - t0 > s.timeout_s) return set_err("not every rank attached");
        usleep(200);
    }
    if (barrier() != 0) return -1;
    if (rank == 0) shm_unlink(nm.c_str()); // every rank holds a mapping now
    return 0;
}

inline void detach()
{
    State &s = st();
    if (s.h) munmap((void *)s.h, s.map_bytes);
    s.h = nullptr;
    s.slots = nullptr;
    s.world = 1;
    s.rank = 0;
}

} // namespace qcecomm
