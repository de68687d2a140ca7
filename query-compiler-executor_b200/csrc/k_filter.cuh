// k_filter.cuh -- predicate scans and order-preserving row-id compaction.
//
// Replaces exec_filter_rel_no_exists (src/filter.c:37-64),
// exec_filter_rel_exists (src/filter.c:3-35) and the compare+push loop of
// scan_join (src/join.c:407-412) of the reference.
//
// Shape: two streaming passes joined by a one-CTA scan, no inter-CTA waiting.
//   pass 1  reads the column with 128-bit loads, evaluates the predicate,
//           warp-ballots it into a bit mask (n/8 bytes) and counts per tile
//   scan    exclusive scan of the per-tile counts (k_scan_excl)
//   pass 2  reads only the bit mask, ranks set bits with popc prefixes,
//           stages the surviving row ids in shared memory and writes them
//           out coalesced, in ascending position order
// HBM traffic for the base scan: 8 B/row in + n/8 B mask (written, re-read)
// + 4 B per surviving row out (row ids are uint32 on the device).
#pragma once
#include "qce_common.cuh"

#define QCE_FTILE 4096      // elements per tile (256 threads x 16)
#define QCE_FTHREADS 256
#define QCE_FWORDS 128      // mask words per tile

enum { QCE_LAYOUT_PAIR = 0, QCE_LAYOUT_NATURAL = 1 };
enum { QCE_EMIT_INDEX = 0, QCE_EMIT_SRC = 1, QCE_EMIT_SRC2 = 2, QCE_EMIT_PACKED = 3 };

// ---- pass 1, base column ----------------------------------------------------
// Mask layout PAIR: for the 64-element group G, word 2G holds the predicate of
// elements 64G+2l (bit l), word 2G+1 of elements 64G+2l+1 -- exactly what two
// ballots over a warp of 128-bit loads produce; pass 2 knows the layout.
template <int OP>
__global__ void __launch_bounds__(QCE_FTHREADS)
k_filter_mask_base(const u64 *__restrict__ col, u64 n, u64 c, u32 *__restrict__ mask,
                   u32 *__restrict__ tile_count)
{
    __shared__ u32 wcnt[QCE_FTHREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 base = (u64)blockIdx.x * QCE_FTILE;
    u32 *words = mask + (u64)blockIdx.x * QCE_FWORDS;
    u32 cnt = 0;

    if (base + QCE_FTILE <= n) {
        u64 a[8], b[8];
#pragma unroll
        for (int k = 0; k < 8; k++) ld_stream_u64x2(col + base + k * 512 + tid * 2, a[k], b[k]);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            u32 b0 = __ballot_sync(QCE_FULL_MASK, qce_pred<OP>(a[k], c));
            u32 b1 = __ballot_sync(QCE_FULL_MASK, qce_pred<OP>(b[k], c));
            if (lane == 0) {
                int g = k * 8 + warp;
                *reinterpret_cast<uint2 *>(words + 2 * g) = make_uint2(b0, b1);
            }
            cnt += __popc(b0) + __popc(b1);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            u64 e = base + k * 512 + tid * 2;
            bool pa = false, pb = false;
            if (e < n) pa = qce_pred<OP>(col[e], c);
            if (e + 1 < n) pb = qce_pred<OP>(col[e + 1], c);
            u32 b0 = __ballot_sync(QCE_FULL_MASK, pa);
            u32 b1 = __ballot_sync(QCE_FULL_MASK, pb);
            if (lane == 0) {
                int g = k * 8 + warp;
                *reinterpret_cast<uint2 *>(words + 2 * g) = make_uint2(b0, b1);
            }
            cnt += __popc(b0) + __popc(b1);
        }
    }
    if (lane == 0) wcnt[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
        u32 t = 0;
#pragma unroll
        for (int w = 0; w < QCE_FTHREADS / 32; w++) t += wcnt[w];
        tile_count[blockIdx.x] = t;
    }
}

// ---- pass 1, through a row-id column (filter refine) --------------------------
// Mask layout NATURAL: word w holds elements 32w .. 32w+31.
template <int OP>
__global__ void __launch_bounds__(QCE_FTHREADS)
k_refine_mask(const u32 *__restrict__ ids, u64 n, const __grid_constant__ ColRef col, u64 c,
              u32 *__restrict__ mask, u32 *__restrict__ tile_count)
{
    __shared__ u32 wcnt[QCE_FTHREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 base = (u64)blockIdx.x * QCE_FTILE;
    u32 *words = mask + (u64)blockIdx.x * QCE_FWORDS;
    u32 cnt = 0;
    u32 id[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        u64 i = base + k * 256 + tid;
        id[k] = (i < n) ? ids[i] : 0xffffffffu;
    }
    u64 v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = (id[k] != 0xffffffffu) ? col(id[k]) : 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        bool p = (id[k] != 0xffffffffu) && qce_pred<OP>(v[k], c);
        u32 b = __ballot_sync(QCE_FULL_MASK, p);
        if (lane == 0) words[k * 8 + warp] = b;
        cnt += __popc(b);
    }
    if (lane == 0) wcnt[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
        u32 t = 0;
#pragma unroll
        for (int w = 0; w < QCE_FTHREADS / 32; w++) t += wcnt[w];
        tile_count[blockIdx.x] = t;
    }
}

// ---- pass 1, positional key equality (scan_join, src/join.c:407-412) ---------
// idsR / idsS may be NULL: the side is then a base column (row id = position).
__global__ void __launch_bounds__(QCE_FTHREADS)
k_scanjoin_mask(const u32 *__restrict__ idsR, const __grid_constant__ ColRef colR,
                const u32 *__restrict__ idsS, const __grid_constant__ ColRef colS, u64 n,
                u32 *__restrict__ mask, u32 *__restrict__ tile_count)
{
    __shared__ u32 wcnt[QCE_FTHREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 base = (u64)blockIdx.x * QCE_FTILE;
    u32 *words = mask + (u64)blockIdx.x * QCE_FWORDS;
    u32 cnt = 0;
#pragma unroll 4
    for (int k = 0; k < 16; k++) {
        u64 i = base + k * 256 + tid;
        bool p = false;
        if (i < n) {
            u32 r = idsR ? idsR[i] : (u32)i;
            u32 s = idsS ? idsS[i] : (u32)i;
            p = colR(r) == colS(s);
        }
        u32 b = __ballot_sync(QCE_FULL_MASK, p);
        if (lane == 0) words[k * 8 + warp] = b;
        cnt += __popc(b);
    }
    if (lane == 0) wcnt[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
        u32 t = 0;
#pragma unroll
        for (int w = 0; w < QCE_FTHREADS / 32; w++) t += wcnt[w];
        tile_count[blockIdx.x] = t;
    }
}

// ---- pass 1, "differs from predecessor" over sorted packed words --------------
// Head flags for the distinct-pair pass (Hashmap dedup, src/join.c:358-367).
__global__ void __launch_bounds__(QCE_FTHREADS)
k_unique_mask(const u64 *__restrict__ w, u64 n, u32 *__restrict__ mask,
              u32 *__restrict__ tile_count)
{
    __shared__ u32 wcnt[QCE_FTHREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 base = (u64)blockIdx.x * QCE_FTILE;
    u32 *words = mask + (u64)blockIdx.x * QCE_FWORDS;
    u32 cnt = 0;
#pragma unroll 4
    for (int k = 0; k < 16; k++) {
        u64 i = base + k * 256 + tid;
        bool p = false;
        if (i < n) p = (i == 0) || (w[i] != w[i - 1]);
        u32 b = __ballot_sync(QCE_FULL_MASK, p);
        if (lane == 0) words[k * 8 + warp] = b;
        cnt += __popc(b);
    }
    if (lane == 0) wcnt[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
        u32 t = 0;
#pragma unroll
        for (int w2 = 0; w2 < QCE_FTHREADS / 32; w2++) t += wcnt[w2];
        tile_count[blockIdx.x] = t;
    }
}

// ---- exclusive scan of per-tile values by one CTA ------------------------------
// n is the number of tiles (<= a few hundred thousand); total goes to *total.
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(1024)
k_scan_excl(const Tin *__restrict__ in, Tout *__restrict__ out, u64 n, u64 *__restrict__ total)
{
    __shared__ Tout scratch[33];
    Tout running = 0;
    for (u64 base = 0; base < n; base += 4096) {
        u64 i0 = base + (u64)threadIdx.x * 4;
        Tout v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = (i0 + k < n) ? (Tout)in[i0 + k] : (Tout)0;
        Tout s = v[0] + v[1] + v[2] + v[3];
        Tout tot;
        Tout ex = block_scan_excl<Tout, 1024>(s, scratch, &tot) + running;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k < n) out[i0 + k] = ex;
            ex += v[k];
        }
        running += tot;
    }
    if (threadIdx.x == 0) *total = (u64)running;
}
// Same for a handful of tiles (small queries): one warp, no block barriers.
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(32)
k_scan_excl_warp(const Tin *__restrict__ in, Tout *__restrict__ out, u64 n, u64 *__restrict__ total)
{
    Tout running = 0;
    for (u64 base = 0; base < n; base += 128) {
        u64 i0 = base + (u64)threadIdx.x * 4;
        Tout v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = (i0 + k < n) ? (Tout)in[i0 + k] : (Tout)0;
        Tout s = v[0] + v[1] + v[2] + v[3];
        Tout incl = warp_scan_incl<Tout>(s);
        Tout ex = incl - s + running;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k < n) out[i0 + k] = ex;
            ex += v[k];
        }
        running += __shfl_sync(QCE_FULL_MASK, incl, 31);
    }
    if (threadIdx.x == 0) *total = (u64)running;
}

// ---- pass 2: mask -> compacted outputs -----------------------------------------
// MODE  QCE_EMIT_INDEX   out0[r] = position
//       QCE_EMIT_SRC     out0[r] = src0[position]                (uint32 source)
//       QCE_EMIT_SRC2    out0[r] = src0[position], out1[r] = src1[position]
//                        (a NULL source stands for the position itself)
//       QCE_EMIT_PACKED  src0 is uint64: out0[r] = hi32, out1[r] = lo32
template <int LAYOUT, int MODE>
__global__ void __launch_bounds__(QCE_FTHREADS)
k_compact(const u32 *__restrict__ mask, const u32 *__restrict__ tile_off,
          const u32 *__restrict__ tile_count, const void *__restrict__ src0,
          const u32 *__restrict__ src1, u32 *__restrict__ out0, u32 *__restrict__ out1, u32 id_base)
{
    __shared__ u32 s0[QCE_FTILE];
    __shared__ u32 s1[(MODE >= QCE_EMIT_SRC2) ? QCE_FTILE : 1];
    __shared__ u32 gt[64], goff[64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 cnt = tile_count[blockIdx.x];
    if (cnt == 0) return;
    const u64 base = (u64)blockIdx.x * QCE_FTILE;
    const u32 *words = mask + (u64)blockIdx.x * QCE_FWORDS;

    if (tid < 64) gt[tid] = __popc(words[2 * tid]) + __popc(words[2 * tid + 1]);
    __syncthreads();
    if (warp == 0) {
        u32 a = gt[2 * lane], b = gt[2 * lane + 1];
        u32 incl = warp_scan_incl<u32>(a + b);
        u32 ex = incl - (a + b);
        goff[2 * lane] = ex;
        goff[2 * lane + 1] = ex + a;
    }
    __syncthreads();
    const u32 lt = lanemask_lt();
    for (int g = warp; g < 64; g += QCE_FTHREADS / 32) {
        const u32 b0 = words[2 * g], b1 = words[2 * g + 1];
        if ((b0 | b1) == 0) continue;
        u64 ea, eb;
        u32 ra, rb;
        const bool pa = (b0 >> lane) & 1u, pb = (b1 >> lane) & 1u;
        if (LAYOUT == QCE_LAYOUT_PAIR) {
            ea = base + g * 64 + 2 * lane;
            eb = ea + 1;
            ra = goff[g] + __popc(b0 & lt) + __popc(b1 & lt);
            rb = ra + (pa ? 1u : 0u);
        } else {
            ea = base + g * 64 + lane;
            eb = ea + 32;
            ra = goff[g] + __popc(b0 & lt);
            rb = goff[g] + __popc(b0) + __popc(b1 & lt);
        }
        if (MODE == QCE_EMIT_INDEX) {
            if (pa) s0[ra] = (u32)ea + id_base;
            if (pb) s0[rb] = (u32)eb + id_base;
        } else if (MODE == QCE_EMIT_SRC) {
            const u32 *s = (const u32 *)src0;
            if (pa) s0[ra] = s[ea];
            if (pb) s0[rb] = s[eb];
        } else if (MODE == QCE_EMIT_SRC2) {
            const u32 *s = (const u32 *)src0;
            if (pa) { s0[ra] = s ? s[ea] : (u32)ea; s1[ra] = src1 ? src1[ea] : (u32)ea; }
            if (pb) { s0[rb] = s ? s[eb] : (u32)eb; s1[rb] = src1 ? src1[eb] : (u32)eb; }
        } else {
            const u64 *s = (const u64 *)src0;
            if (pa) { u64 w = s[ea]; s0[ra] = (u32)(w >> 32); s1[ra] = (u32)w; }
            if (pb) { u64 w = s[eb]; s0[rb] = (u32)(w >> 32); s1[rb] = (u32)w; }
        }
    }
    __syncthreads();
    const u64 o = tile_off[blockIdx.x];
    for (u32 p = tid; p < cnt; p += QCE_FTHREADS) {
        out0[o + p] = s0[p];
        if (MODE >= QCE_EMIT_SRC2) out1[o + p] = s1[p];
    }
}
