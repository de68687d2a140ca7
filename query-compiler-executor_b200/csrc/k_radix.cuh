// k_radix.cuh -- tuple build, digit histograms and the one-sweep LSD radix pass.
//
// Replaces allocate_relation / allocate_relation_mid_results
// (src/join.c:96-142), build_histogram / build_psum / build_reordered_array
// (src/utilities.c:20-70), iterative_sort (src/join.c:5-94) and the
// quicksort leaves (src/quicksort.c:54-64) of the reference.
//
// The reference sorts 16-byte AoS tuples MSD-first through 8 byte levels
// (4 of them whole-array no-ops when keys < 2^32) and finishes buckets of
// < 4096 tuples with an unstable randomized quicksort.  Here:
//   * a run is packed to one 8-byte word (key << 32 | rowid) whenever the
//     source column's maximum is < 2^32 (known from the load-time column
//     statistics), halving the bytes every pass moves; wider keys use SoA
//     uint64 keys + uint32 ids;
//   * only the ceil(bitlen(max key)/8) significant digits are sorted;
//   * each digit is one kernel that reads every tuple once and writes it once
//     ("one sweep"): per-tile digit counts are published to a status array and
//     the tile's global offsets come from a decoupled look-back over the
//     preceding tiles, so there is no separate per-pass counting kernel;
//   * ranks inside a tile come from warp match/popc (stable), tuples are
//     staged in shared memory in digit order and leave in coalesced runs.
// The sort is stable, so equal keys keep their input order (the reference's
// tie order is unspecified).
//
// Algorithmic HBM bytes: 8 B/tuple once (histogram) + 16 B/tuple/pass packed
// (32 B/tuple/pass for the reference's 16-byte tuples; 24 B wide SoA here).
#pragma once
#include "qce_common.cuh"

#define QCE_RADIX_BITS 8
#define QCE_RADIX_BINS 256
#define QCE_MAX_PASSES 8

struct RadixShifts {
    int npass;
    int shift[QCE_MAX_PASSES];
};

// ---- tuple build ---------------------------------------------------------------
// Packed, from a base column: out[i] = col[i] << 32 | i   (src/join.c:131-134)
__global__ void __launch_bounds__(256)
k_build_packed_base(const u64 *__restrict__ col, u64 n, u64 *__restrict__ out, u64 id_base)
{
    // id_base: first row id of the window when a rank builds only its row range
    const u64 stride = (u64)gridDim.x * 512;
    for (u64 e = ((u64)blockIdx.x * 256 + threadIdx.x) * 2; e < n; e += stride) {
        if (e + 1 < n) {
            u64 a, b;
            ld_stream_u64x2(col + e, a, b);
            st_stream_u64x2(out + e, (a << 32) | (e + id_base), (b << 32) | (e + 1 + id_base));
        } else {
            out[e] = (col[e] << 32) | (e + id_base);
        }
    }
}
// Packed, through a row-id column: out[i] = col[ids[i]] << 32 | ids[i]
// (src/join.c:107-112).  The gather is sector-bound unless ids are clustered
// (filter outputs are ascending).
__global__ void __launch_bounds__(256)
k_build_packed_ids(const u64 *__restrict__ col, const u32 *__restrict__ ids, u64 n,
                   u64 *__restrict__ out)
{
    const u64 stride = (u64)gridDim.x * 256 * 4;
    for (u64 i0 = (u64)blockIdx.x * 256 * 4 + threadIdx.x; i0 < n; i0 += stride) {
        u32 id[4];
        u64 v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) id[k] = (i0 + k * 256 < n) ? ids[i0 + k * 256] : 0;
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = (i0 + k * 256 < n) ? __ldg(col + id[k]) : 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (i0 + k * 256 < n) out[i0 + k * 256] = (v[k] << 32) | id[k];
    }
}
// Wide (keys may exceed 32 bits): SoA keys[] / ids[].
__global__ void __launch_bounds__(256)
k_build_wide(const u64 *__restrict__ col, const u32 *__restrict__ ids, u64 n,
             u64 *__restrict__ keys, u32 *__restrict__ out_ids, u32 id_base)
{
    const u64 stride = (u64)gridDim.x * 256;
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        u32 id = ids ? ids[i] : (u32)i + id_base;
        keys[i] = __ldg(col + (ids ? (u64)id : i));
        out_ids[i] = id;
    }
}
// Packed from two row-id columns: out[i] = hi[i] << 32 | (lo ? lo[i] : 0).
// Used by the bystander re-join (R' = (key=last[i], payload=edit[i]),
// S' = (key=driver[i], 0); src/join.c:431-442) and the distinct-pair pass.
__global__ void __launch_bounds__(256)
k_pack_pairs(const u32 *__restrict__ hi, const u32 *__restrict__ lo, u64 n, u64 *__restrict__ out)
{
    const u64 stride = (u64)gridDim.x * 256;
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += stride)
        out[i] = ((u64)hi[i] << 32) | (lo ? (u64)lo[i] : 0ull);
}
__global__ void __launch_bounds__(256)
k_unpack_lo(const u64 *__restrict__ in, u64 n, u32 *__restrict__ out)
{
    const u64 stride = (u64)gridDim.x * 256;
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) out[i] = (u32)in[i];
}

// ---- digit histograms of every pass in one read (build_histogram,
// src/utilities.c:20-31, hoisted out of the per-bucket loop) ------------------------
__global__ void __launch_bounds__(512)
k_radix_hist(const u64 *__restrict__ keys, u64 n, RadixShifts rs, u32 *__restrict__ ghist)
{
    __shared__ u32 sh[QCE_MAX_PASSES * QCE_RADIX_BINS];
    for (int i = threadIdx.x; i < rs.npass * QCE_RADIX_BINS; i += 512) sh[i] = 0;
    __syncthreads();
    const u64 stride = (u64)gridDim.x * 1024;
    for (u64 e = ((u64)blockIdx.x * 512 + threadIdx.x) * 2; e < n; e += stride) {
        u64 a, b = 0;
        const bool two = e + 1 < n;
        if (two) ld_stream_u64x2(keys + e, a, b);
        else a = keys[e];
#pragma unroll
        for (int p = 0; p < QCE_MAX_PASSES; p++) {
            if (p < rs.npass) {
                atomicAdd(&sh[p * QCE_RADIX_BINS + ((a >> rs.shift[p]) & 255)], 1u);
                if (two) atomicAdd(&sh[p * QCE_RADIX_BINS + ((b >> rs.shift[p]) & 255)], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < rs.npass * QCE_RADIX_BINS; i += 512)
        if (sh[i]) atomicAdd(&ghist[i], sh[i]);
}

// One-digit histogram of 32-bit keys (row ids), for the bucketed checksum.
__global__ void __launch_bounds__(512)
k_hist_u32(const u32 *__restrict__ keys, u64 n, int shift, u32 *__restrict__ ghist)
{
    __shared__ u32 sh[QCE_RADIX_BINS];
    if (threadIdx.x < QCE_RADIX_BINS) sh[threadIdx.x] = 0;
    __syncthreads();
    const u64 stride = (u64)gridDim.x * 2048;
    const u64 n4 = n & ~3ull;
    for (u64 e = ((u64)blockIdx.x * 512 + threadIdx.x) * 4; e < n4; e += stride) {
        uint4 k = ld_stream_u32x4(keys + e);
        atomicAdd(&sh[(k.x >> shift) & 255], 1u);
        atomicAdd(&sh[(k.y >> shift) & 255], 1u);
        atomicAdd(&sh[(k.z >> shift) & 255], 1u);
        atomicAdd(&sh[(k.w >> shift) & 255], 1u);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n - n4)) atomicAdd(&sh[(keys[n4 + threadIdx.x] >> shift) & 255], 1u);
    __syncthreads();
    if (threadIdx.x < QCE_RADIX_BINS && sh[threadIdx.x]) atomicAdd(&ghist[threadIdx.x], sh[threadIdx.x]);
}

// build_psum (src/utilities.c:34-48): exclusive prefix over the 256 bins of
// each pass.  One CTA of 256 threads per pass.
__global__ void __launch_bounds__(256) k_radix_bases(const u32 *__restrict__ ghist,
                                                     u32 *__restrict__ gbase)
{
    __shared__ u32 scratch[33];
    u32 v = ghist[blockIdx.x * QCE_RADIX_BINS + threadIdx.x];
    u32 tot;
    u32 ex = block_scan_excl<u32, 256>(v, scratch, &tot);
    gbase[blockIdx.x * QCE_RADIX_BINS + threadIdx.x] = ex;
}

// ---- digit functors ------------------------------------------------------------
struct DigitShift {
    int shift;
    __device__ __forceinline__ u32 operator()(u64 k) const { return (u32)(k >> shift) & 255u; }
    __device__ __forceinline__ u32 operator()(u32 k) const { return (k >> shift) & 255u; }
};
// Range partition for the multi-GPU exchange.  Splitters sit on the boundaries of
// the 256-bin histogram of the top 8 significant key bits, so the destination of
// a tuple is a table lookup on those bits (4 destinations per table word).
struct DigitSplit {
    const u32 *lut; // device memory, 256 x uint8: histogram bin -> destination rank.  Not a
                    // kernel-parameter array: lanes index it divergently, which the
                    // constant bank serialises; 256 B in global memory stay in L1.
    int shift;      // 32 + key_bits - 8 (packed word)
    __device__ __forceinline__ u32 operator()(u64 k) const
    {
        const u32 bin = (u32)(k >> shift) & 255u;
        return (__ldg(lut + (bin >> 2)) >> ((bin & 3u) * 8)) & 255u;
    }
};

// Lanes of the warp holding the same 8-bit digit.  `match.any` gives this in one
// instruction, but on sm_100a it runs on the address-divergence unit at about
// one warp instruction per 60 cycles per SM (ncu: pipe_adu 73 % busy, the pass
// capped at ~2 TB/s).  Eight ballots -- one per digit bit, intersected -- go
// through the vote path at two warps per clock instead.
__device__ __forceinline__ u32 warp_peers_8bit(u32 d)
{
    u32 peers = QCE_FULL_MASK;
#pragma unroll
    for (int b = 0; b < QCE_RADIX_BITS; b++) {
        const bool bit = (d >> b) & 1u;
        const u32 m = __ballot_sync(QCE_FULL_MASK, bit);
        peers &= bit ? m : ~m;
    }
    return peers;
}

// ---- one-sweep pass ------------------------------------------------------------
// Tile status word: top 2 bits = flag, low 30 bits = count (n < 2^30 per sort).
#define QCE_ST_PART 0x40000000u
#define QCE_ST_INCL 0x80000000u
#define QCE_ST_MASK 0x3fffffffu

template <typename KeyT> __device__ __forceinline__ KeyT ld_stream_key(const KeyT *p);
template <> __device__ __forceinline__ u64 ld_stream_key<u64>(const u64 *p) { return ld_stream_u64(p); }
template <> __device__ __forceinline__ u32 ld_stream_key<u32>(const u32 *p) { return ld_stream_u32(p); }

template <int THREADS, int ITEMS, typename KeyT = u64> struct OnesweepSmem {
    u32 warp_hist[(THREADS / 32) * QCE_RADIX_BINS]; // per-warp digit counts -> bases
    u32 tile_excl[QCE_RADIX_BINS];                  // digit start inside the sorted tile
    u32 goff[QCE_RADIX_BINS];                       // global start of digit minus tile_excl
    u32 tile_cnt[QCE_RADIX_BINS];                   // early digit counts of the tile
    u32 scratch[33];
    u32 tile_id;
    KeyT keys[THREADS * ITEMS];
};

// One tile of the pass.  FULL = every slot of the tile holds a real tuple (no
// bounds checks, no padding logic): all tiles but the last.
template <int THREADS, int ITEMS, bool HAS_VALS, bool FULL, bool EARLY, int MATCH_EVERY, typename KeyT, typename DigitOp>
__device__ __forceinline__ void
onesweep_tile(OnesweepSmem<THREADS, ITEMS, KeyT> &sm, u32 *svals, const KeyT *__restrict__ keys_in,
              KeyT *__restrict__ keys_out, const u32 *__restrict__ vals_in, u32 *__restrict__ vals_out, u32 n,
              const DigitOp &digit, const u32 *__restrict__ gbase, u32 *__restrict__ status, u32 tile)
{
    constexpr int TILE = THREADS * ITEMS;
    constexpr int WARPS = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 tbase = tile * TILE;
    const u32 nvalid = FULL ? (u32)TILE : (n - tbase);

    // ---- load: warp-striped so that (warp, item, lane) order == input order
    KeyT key[ITEMS];
    u32 val[HAS_VALS ? ITEMS : 1];
    const u32 wbase = tbase + warp * (32 * ITEMS) + lane;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const u32 idx = wbase + j * 32;
        if (FULL || idx < n) {
            key[j] = ld_stream_key<KeyT>(keys_in + idx);
            if (HAS_VALS) val[j] = ld_stream_u32(vals_in + idx);
        } else {
            key[j] = (KeyT)~(KeyT)0;
            if (HAS_VALS) val[j] = 0u;
        }
    }

    // ---- early counts (EARLY): tile digit histogram with no-return shared
    // atomics, so the tile's PARTIAL status is published before the (long)
    // ranking phase and successors' look-back finds it without spinning.
    u32 my_count = 0;
    if (EARLY) {
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const u32 d = (!FULL && (wbase + j * 32) >= n) ? 255u : digit(key[j]);
            atomicAdd(&sm.tile_cnt[d], 1u);
        }
        __syncthreads();
        if (tid < QCE_RADIX_BINS) {
            my_count = sm.tile_cnt[tid];
            // the padding of the last tile is not part of the global count
            const u32 pub = my_count - ((!FULL && tid == 255) ? (TILE - nvalid) : 0u);
            st_relaxed_gpu_u32(status + (size_t)tile * QCE_RADIX_BINS + tid,
                               (tile == 0 ? QCE_ST_INCL : QCE_ST_PART) | pub);
        }
    }

    // ---- rank inside the warp.  Lanes with the same digit are found with
    // per-bit ballots (warp_peers_8bit); the lowest of them reserves the group's slots with one shared
    // atomicAdd (program order within the warp keeps items in input order) and
    // hands the base to its peers.  Done in chunks of CH items with the three
    // long-latency steps (ballots, ATOMS, SHFL) each issued back to back, so their
    // latencies overlap instead of adding up.
    u32 rd[ITEMS]; // digit << 16 | rank inside (warp, digit)
    u32 *wh = sm.warp_hist + warp * QCE_RADIX_BINS;
    const u32 lt = lanemask_lt();
    constexpr int CH = (ITEMS % 4 == 0) ? 4 : ((ITEMS % 3 == 0) ? 3 : 1);
#pragma unroll
    for (int j0 = 0; j0 < ITEMS; j0 += CH) {
        u32 d[CH], peers[CH], old[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) {
            d[c] = (!FULL && (wbase + (j0 + c) * 32) >= n) ? 255u : digit(key[j0 + c]);
            // every MATCH_EVERY-th item goes to match.any: the ADU pipe is idle
            // otherwise, so it takes a share of the ranking off the ALU ballots
            peers[c] = (MATCH_EVERY > 0 && ((j0 + c) % (MATCH_EVERY > 0 ? MATCH_EVERY : 1)) == MATCH_EVERY - 1)
                           ? __match_any_sync(QCE_FULL_MASK, d[c])
                           : warp_peers_8bit(d[c]);
        }
#pragma unroll
        for (int c = 0; c < CH; c++) {
            old[c] = 0;
            if ((peers[c] & lt) == 0) old[c] = atomicAdd(&wh[d[c]], (u32)__popc(peers[c]));
        }
#pragma unroll
        for (int c = 0; c < CH; c++) {
            old[c] = __shfl_sync(QCE_FULL_MASK, old[c], __ffs(peers[c]) - 1);
            rd[j0 + c] = (d[c] << 16) | (old[c] + __popc(peers[c] & lt));
        }
    }
    __syncthreads(); // all warps have ranked
    // ---- per digit: exclusive scan of the warp counts
    if (tid < QCE_RADIX_BINS) {
        u32 run = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) {
            const u32 c = sm.warp_hist[w * QCE_RADIX_BINS + tid];
            sm.warp_hist[w * QCE_RADIX_BINS + tid] = run;
            run += c;
        }
        if (!EARLY) {
            my_count = run;
            const u32 pub = my_count - ((!FULL && tid == 255) ? (TILE - nvalid) : 0u);
            st_relaxed_gpu_u32(status + (size_t)tile * QCE_RADIX_BINS + tid,
                               (tile == 0 ? QCE_ST_INCL : QCE_ST_PART) | pub);
        }
    }
    // ---- digit starts inside the tile (exclusive scan over the 256 digits)
    {
        u32 tot;
        u32 ex = block_scan_excl<u32, THREADS>((tid < QCE_RADIX_BINS) ? my_count : 0u, sm.scratch, &tot);
        if (tid < QCE_RADIX_BINS) sm.tile_excl[tid] = ex;
    }
    if (tid < QCE_RADIX_BINS) {
        // ---- decoupled look-back: sum the counts of the preceding tiles
        u32 excl = 0;
        if (tile > 0) {
            const u32 pub = my_count - ((!FULL && tid == 255) ? (TILE - nvalid) : 0u);
            int p = (int)tile - 1;
            while (true) {
                const u32 v = ld_relaxed_gpu_u32(status + (size_t)p * QCE_RADIX_BINS + tid);
                if ((v & ~QCE_ST_MASK) == 0) { __nanosleep(20); continue; }
                excl += v & QCE_ST_MASK;
                if (v & QCE_ST_INCL) break;
                p--;
            }
            st_relaxed_gpu_u32(status + (size_t)tile * QCE_RADIX_BINS + tid, QCE_ST_INCL | (excl + pub));
        }
        sm.goff[tid] = gbase[tid] + excl - sm.tile_excl[tid];
    }
    __syncthreads();

    // ---- stage in shared memory in digit order
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const u32 d = rd[j] >> 16;
        const u32 pos = sm.tile_excl[d] + wh[d] + (rd[j] & 0xffffu);
        sm.keys[pos] = key[j];
        if (HAS_VALS) svals[pos] = val[j];
    }
    __syncthreads();

    // ---- coalesced write-out: consecutive threads, consecutive slots of a digit
    //      (positions >= nvalid are exactly the padding of the last tile)
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const u32 p = tid + j * THREADS;
        if (FULL || p < nvalid) {
            const KeyT k = sm.keys[p];
            const u32 g = sm.goff[digit(k)] + p;
            keys_out[g] = k;
            if (HAS_VALS) vals_out[g] = svals[p];
        }
    }
}

template <int THREADS, int ITEMS, int MIN_CTAS, bool EARLY, int MATCH_EVERY, bool HAS_VALS, typename KeyT, typename DigitOp>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
k_onesweep(const KeyT *__restrict__ keys_in, KeyT *__restrict__ keys_out,
           const u32 *__restrict__ vals_in, u32 *__restrict__ vals_out, u32 n, DigitOp digit,
           const u32 *__restrict__ gbase, u32 *__restrict__ status, u32 *__restrict__ tile_counter)
{
    static_assert(THREADS >= QCE_RADIX_BINS && THREADS * ITEMS <= 65536, "tile shape");
    constexpr int TILE = THREADS * ITEMS;
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OnesweepSmem<THREADS, ITEMS, KeyT> &sm = *reinterpret_cast<OnesweepSmem<THREADS, ITEMS, KeyT> *>(smem_raw);
    u32 *svals = reinterpret_cast<u32 *>(smem_raw + sizeof(OnesweepSmem<THREADS, ITEMS, KeyT>));

    // Tiles are claimed in launch order so that every tile a CTA may wait on in
    // the look-back is owned by a CTA that is already resident.
    if (threadIdx.x == 0) sm.tile_id = atomicAdd(tile_counter, 1u);
    for (int i = threadIdx.x; i < WARPS * QCE_RADIX_BINS; i += THREADS) sm.warp_hist[i] = 0;
    if (threadIdx.x < QCE_RADIX_BINS) sm.tile_cnt[threadIdx.x] = 0;
    __syncthreads();
    const u32 tile = sm.tile_id;
    if ((tile + 1) * (u32)TILE <= n)
        onesweep_tile<THREADS, ITEMS, HAS_VALS, true, EARLY, MATCH_EVERY, KeyT>(sm, svals, keys_in, keys_out, vals_in, vals_out, n, digit,
                                                      gbase, status, tile);
    else
        onesweep_tile<THREADS, ITEMS, HAS_VALS, false, EARLY, MATCH_EVERY, KeyT>(sm, svals, keys_in, keys_out, vals_in, vals_out, n, digit,
                                                       gbase, status, tile);
}

// 1 if keys[i-1] > keys[i] anywhere (checks the "already sorted" assumption of
// JOIN_SORT_LHS / JOIN_SORT_RHS, src/join.c:647-658).
template <bool WIDE>
__global__ void __launch_bounds__(256) k_is_unsorted(TupleView t, u64 n, u32 *__restrict__ flag)
{
    const u64 stride = (u64)gridDim.x * 256;
    bool bad = false;
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x + 1; i < n; i += stride)
        bad |= tv_key<WIDE>(t, i - 1) > tv_key<WIDE>(t, i);
    if (__any_sync(QCE_FULL_MASK, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}
