// ipc_probe.cu -- does CUDA IPC work between processes that share ONE device (the test box has one GPU),
// and between devices?  Forks `world` ranks BEFORE any CUDA call; rank r uses device r % ndev.
// Also times: pageable-mmap H2D vs threaded pinned staging, and remote 8-byte gathers through the mapping.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("rank %d: %s failed: %s\n", rank, #x, cudaGetErrorString(e)); fflush(stdout); _exit(3); } } while (0)
struct Shm { std::atomic<int> arrived; std::atomic<int> gen; cudaIpcMemHandle_t h[8]; };
static void barrier(Shm *s, int world) {
    int g = s->gen.load();
    if (s->arrived.fetch_add(1) + 1 == world) { s->arrived.store(0); s->gen.fetch_add(1); }
    else while (s->gen.load() == g) { }
}
__global__ void k_fill(unsigned long long *p, size_t n, unsigned long long v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v + i;
}
__global__ void k_gather(const unsigned long long *__restrict__ col, const unsigned *__restrict__ ids, size_t n, unsigned long long *out) {
    unsigned long long acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc += col[ids[i]];
    atomicAdd(out, acc);
}
__global__ void k_ids(unsigned *ids, size_t n, unsigned m) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long x = i * 0x9E3779B97F4A7C15ull; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        ids[i] = (unsigned)(x % m);
    }
}
int main(int argc, char **argv) {
    int world = argc > 1 ? atoi(argv[1]) : 2;
    Shm *s = (Shm *)mmap(NULL, sizeof(Shm), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    new (s) Shm(); s->arrived = 0; s->gen = 0;
    int rank = 0;
    for (int r = 1; r < world; r++) { pid_t p = fork(); if (p == 0) { rank = r; break; } }
    int ndev = 0; CK(cudaGetDeviceCount(&ndev));
    CK(cudaSetDevice(rank % ndev));
    const size_t n = 64u << 20; // 64 M u64 = 512 MB
    unsigned long long *buf; CK(cudaMalloc(&buf, n * 8));
    k_fill<<<1024, 256>>>(buf, n, (unsigned long long)rank << 40); CK(cudaDeviceSynchronize());
    CK(cudaIpcGetMemHandle(&s->h[rank], buf));
    barrier(s, world);
    int peer = (rank + 1) % world;
    unsigned long long *pb; CK(cudaIpcOpenMemHandle((void **)&pb, s->h[peer], cudaIpcMemLazyEnablePeerAccess));
    unsigned long long probe[2]; CK(cudaMemcpy(probe, pb + 5, 8, cudaMemcpyDeviceToHost));
    printf("rank %d (dev %d of %d): peer %d word5 = %llx (want %llx) %s\n", rank, rank % ndev, ndev, peer, probe[0], ((unsigned long long)peer << 40) + 5, probe[0] == ((unsigned long long)peer << 40) + 5 ? "OK" : "BAD");
    // remote random gathers vs local
    const size_t m = 32u << 20; unsigned *ids; CK(cudaMalloc(&ids, m * 4)); unsigned long long *out; CK(cudaMalloc(&out, 8));
    k_ids<<<1024, 256>>>(ids, m, (unsigned)n); CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 2; which++) {
        const unsigned long long *src = which ? pb : buf;
        k_gather<<<148 * 8, 256>>>(src, ids, m, out); CK(cudaDeviceSynchronize());
        barrier(s, world);
        cudaEventRecord(e0); k_gather<<<148 * 8, 256>>>(src, ids, m, out); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("rank %d: %s random 8B gathers: %.3f ms for %zu M -> %.1f G gathers/s\n", rank, which ? "PEER " : "LOCAL", ms, m >> 20, m / ms / 1e6);
        barrier(s, world);
    }
    // peer streaming copy
    unsigned long long *dst; CK(cudaMalloc(&dst, n * 8));
    CK(cudaMemcpy(dst, pb, n * 8, cudaMemcpyDeviceToDevice)); barrier(s, world);
    cudaEventRecord(e0); CK(cudaMemcpyAsync(dst, pb, n * 8, cudaMemcpyDeviceToDevice)); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    { float ms; cudaEventElapsedTime(&ms, e0, e1); printf("rank %d: peer->local copy 512 MB: %.3f ms = %.1f GB/s\n", rank, ms, n * 8 / ms / 1e6); }
    barrier(s, world);
    if (rank == 0) {
        // H2D: pageable (malloc, touched) vs threaded pinned staging
        const size_t bytes = 1ull << 30; char *src = (char *)malloc(bytes); memset(src, 1, bytes);
        char *d; CK(cudaMalloc(&d, bytes));
        auto t0 = std::chrono::steady_clock::now(); CK(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
        double s0 = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("pageable H2D 1 GB: %.3f s = %.1f GB/s\n", s0, bytes / s0 / 1e9);
        for (int nthr : {2, 4, 8}) {
            const size_t CH = 8u << 20; const int NB = nthr * 2; char *pin; CK(cudaMallocHost(&pin, CH * NB));
            std::vector<cudaStream_t> st(nthr); std::vector<cudaEvent_t> ev(NB);
            for (auto &x : st) cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking);
            for (auto &x : ev) cudaEventCreateWithFlags(&x, cudaEventDisableTiming);
            std::atomic<size_t> next{0}; const size_t nch = bytes / CH;
            t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int t = 0; t < nthr; t++) th.emplace_back([&, t] {
                cudaSetDevice(0); int flip = 0; size_t c;
                while ((c = next.fetch_add(1)) < nch) {
                    int b = t * 2 + flip; flip ^= 1;
                    cudaEventSynchronize(ev[b]);
                    memcpy(pin + b * CH, src + c * CH, CH);
                    cudaMemcpyAsync(d + c * CH, pin + b * CH, CH, cudaMemcpyHostToDevice, st[t]);
                    cudaEventRecord(ev[b], st[t]);
                }
                cudaStreamSynchronize(st[t]);
            });
            for (auto &x : th) x.join();
            double s1 = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("pinned staging, %d threads: %.3f s = %.1f GB/s\n", nthr, s1, bytes / s1 / 1e9);
            cudaFreeHost(pin);
        }
        while (wait(NULL) > 0) { }
    }
    fflush(stdout);
    _exit(0);
}
