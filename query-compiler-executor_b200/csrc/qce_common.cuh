// qce_common.cuh -- shared device helpers for the sm_100a kernels.
//
// Everything on this path is unsigned 64-bit integer streaming work that is
// bound by HBM bandwidth: no tensor cores, no floating point.
// The helpers below are the three things every kernel needs: streaming
// (no-L1-allocate) vector loads/stores, warp/block reductions and scans, and
// the packed tuple accessors.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#define QCE_FULL_MASK 0xffffffffu

// ---------------------------------------------------------------- loads/stores
// Streaming 128-bit load of two adjacent uint64 (read-only path, do not
// allocate in L1: each base-column byte is touched once per scan).
__device__ __forceinline__ void ld_stream_u64x2(const u64 *p, u64 &a, u64 &b)
{
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];"
                 : "=l"(a), "=l"(b)
                 : "l"(p));
}
__device__ __forceinline__ u64 ld_stream_u64(const u64 *p)
{
    u64 a;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(a) : "l"(p));
    return a;
}
__device__ __forceinline__ u32 ld_stream_u32(const u32 *p)
{
    u32 a;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(a) : "l"(p));
    return a;
}
__device__ __forceinline__ uint4 ld_stream_u32x4(const u32 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_u64x2(u64 *p, u64 a, u64 b)
{
    asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b)
                 : "memory");
}
__device__ __forceinline__ void st_stream_u64(u64 *p, u64 a)
{
    asm volatile("st.global.L1::no_allocate.u64 [%0], %1;" ::"l"(p), "l"(a) : "memory");
}
__device__ __forceinline__ void st_stream_u32(u32 *p, u32 a)
{
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(a) : "memory");
}
// Look-back status words: single-word publish/poll at GPU scope.
__device__ __forceinline__ u32 ld_relaxed_gpu_u32(const u32 *p)
{
    u32 a;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(a) : "l"(p) : "memory");
    return a;
}
__device__ __forceinline__ void st_relaxed_gpu_u32(u32 *p, u32 a)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(a) : "memory");
}

__device__ __forceinline__ u64 ld_relaxed_gpu_u64(const u64 *p)
{
    u64 a;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    return a;
}
__device__ __forceinline__ void st_relaxed_gpu_u64(u64 *p, u64 a)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(a) : "memory");
}

__device__ __forceinline__ u32 lanemask_lt()
{
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ---------------------------------------------------------------- predicates
// The reference's three filter operators, unsigned (src/filter.c:9-26,42-56).
enum { QCE_OP_EQ = 0, QCE_OP_GT = 1, QCE_OP_LT = 2 };
template <int OP> __device__ __forceinline__ bool qce_pred(u64 v, u64 c)
{
    if (OP == QCE_OP_EQ) return v == c;
    if (OP == QCE_OP_GT) return v > c;
    return v < c;
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ u64 warp_sum_u64(u64 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(QCE_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ u32 warp_sum_u32(u32 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(QCE_FULL_MASK, v, o);
    return v;
}
// Inclusive warp scan.
template <typename T> __device__ __forceinline__ T warp_scan_incl(T v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T n = __shfl_up_sync(QCE_FULL_MASK, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// Block-wide sum, result valid in every thread.  `scratch` >= 33 elements.
template <typename T, int THREADS> __device__ __forceinline__ T block_sum(T v, T *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(QCE_FULL_MASK, v, o);
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        T w = (lane < THREADS / 32) ? scratch[lane] : (T)0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(QCE_FULL_MASK, w, o);
        if (lane == 0) scratch[32] = w;
    }
    __syncthreads();
    T r = scratch[32];
    __syncthreads();
    return r;
}

// Block-wide exclusive scan of one value per thread; returns the exclusive
// prefix, *total gets the block sum.  `scratch` >= 33 elements.
template <typename T, int THREADS>
__device__ __forceinline__ T block_scan_excl(T v, T *scratch, T *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T incl = warp_scan_incl<T>(v);
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        T w = (lane < THREADS / 32) ? scratch[lane] : (T)0;
        T wi = warp_scan_incl<T>(w);
        scratch[lane] = wi - w; // exclusive prefix of each warp
        if (lane == 31) scratch[32] = wi;
    }
    __syncthreads();
    T r = scratch[warp] + incl - v;
    *total = scratch[32];
    __syncthreads();
    return r;
}

// ---------------------------------------------------------------- base columns
// A base column as the gathering kernels see it.  Whole column in this GPU's HBM
// (one GPU, or a replicated column): rpr == 0, value = d[id].  Row-sharded over
// the ranks of the node: rank r holds rows [r * rpr, (r+1) * rpr),
// vb[r] = (r's window pointer, mapped through CUDA IPC) - r * rpr, so the value is
// vb[owner(id)][id] -- an NVLink load when the owner is a peer.
#define QCE_MAX_RANKS 16
struct ColRef {
    const u64 *d;
    u32 rpr, last;   // rows per rank, last rank
    float inv;       // 1 / rpr, rounded down
    const u64 *vb[QCE_MAX_RANKS];
    __device__ __forceinline__ u32 owner(u32 id) const
    {
        u32 r = min(__float2uint_rz(__uint2float_rz(id) * inv), last);
        r -= (id < r * rpr);
        r += (r < last) & (id >= (r + 1) * rpr);
        return r;
    }
    __device__ __forceinline__ u64 operator()(u32 id) const
    {
        if (rpr == 0) return __ldg(d + id);
        return __ldg(vb[owner(id)] + id);
    }
};

// ---------------------------------------------------------------- tuple runs
// A run of (key,rowid) tuples.  Packed: one uint64 = key << 32 | rowid (every
// key < 2^32).  Wide: separate uint64 keys[] and uint32 ids[] arrays (SoA).
struct TupleView {
    const u64 *a;   // packed words, or keys when wide
    const u32 *ids; // wide only
};
template <bool WIDE> __device__ __forceinline__ u64 tv_key(const TupleView &t, size_t i)
{
    return WIDE ? t.a[i] : (t.a[i] >> 32);
}
template <bool WIDE> __device__ __forceinline__ u32 tv_id(const TupleView &t, size_t i)
{
    return WIDE ? t.ids[i] : (u32)t.a[i];
}
