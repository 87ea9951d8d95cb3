// primitives.cuh — hand-written device-wide exclusive scan and stable LSD radix sort (key,value pairs).
// Used by the spatial-index rebuild: counting sort of centroids into grid bins, counting sort of particles into
// Voronoi cells (voronoi.h:217-231) and the Morton sort of the centroids (reorder_morton.h:44-122).
#pragma once
#include "common.cuh"

namespace orbc {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;           // two 16-byte loads per thread; small tiles: the latency of ONE tile bounds the whole scan
constexpr int kScanTile = kScanThreads * kScanItems;

// exclusive scan of `v` across the block; returns the exclusive prefix for this thread and the block total
__device__ __forceinline__ int block_exclusive_scan(int v, int &total) {
    __shared__ int warp_sums[33];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int incl = v;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const int w = lane < nw ? warp_sums[lane] : 0;
        int wi = w;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += y; }
        warp_sums[lane] = wi - w;             // exclusive warp offsets
        if (lane == 31) warp_sums[32] = wi;   // block total
    }
    __syncthreads();
    total = warp_sums[32];
    return warp_sums[wid] + incl - v;
}

// ---- single-pass scan (decoupled look-back, Merrill & Garland 2016) -------------------------------------------------------------
// One launch instead of three: a tile publishes its aggregate, looks back over its predecessors' descriptors until it meets an
// inclusive prefix, then publishes its own.  Tiles take their index from a counter in scheduling order, so a tile only ever
// waits for tiles that are already resident.  Descriptors carry the scan's epoch, so nothing has to be cleared between scans:
//   desc = epoch << 34 | state << 32 | value     state 1 = aggregate, 2 = inclusive prefix
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
// (data1: a second array of the same length scanned by the same launch -- blocks [n_tiles, 2 n_tiles) -- with its own descriptors
// and tile counter behind those of the first)
__global__ void __launch_bounds__(kScanThreads) k_scan_onepass(int *__restrict__ data, int *__restrict__ data1, int n, int n_tiles, unsigned long long *__restrict__ desc,
                                                               unsigned *__restrict__ counter, unsigned epoch) {
    __shared__ int s_tile, s_prefix;
    if (data1 && (int)blockIdx.x >= n_tiles) { data = data1; desc += n_tiles; counter += 1; }
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(counter, 1u);
    __syncthreads();
    const int tile = s_tile;
    const int base = tile * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems], s = 0;
    const bool whole = base + kScanItems <= n;                   // `data` is a cudaMalloc base pointer: 16-byte aligned groups
    if (whole) {
        const int4 a = *reinterpret_cast<const int4 *>(data + base), b = *reinterpret_cast<const int4 *>(data + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        #pragma unroll
        for (int k = 0; k < kScanItems; ++k) v[k] = base + k < n ? data[base + k] : 0;
    }
    #pragma unroll
    for (int k = 0; k < kScanItems; ++k) s += v[k];
    int total; int off = block_exclusive_scan(s, total);
    if (threadIdx.x < 32) {                                      // warp 0: publish, look back 32 predecessors at a time
        const int lane = threadIdx.x;
        const unsigned long long tag = (unsigned long long)epoch << 34;
        int prefix = 0;
        if (tile == 0) { if (lane == 0) st_release_u64(desc, tag | (2ull << 32) | (unsigned)total); }
        else {
            if (lane == 0) st_release_u64(desc + tile, tag | (1ull << 32) | (unsigned)total);
            for (int t = tile - 1; ; t -= 32) {
                const int idx = t - lane;                        // lane 0 looks at the nearest predecessor
                unsigned long long d = tag | (2ull << 32);       // before tile 0: an inclusive prefix of zero
                bool ready;
                do {
                    if (idx >= 0) d = ld_acquire_u64(desc + idx);
                    ready = (d >> 34) == epoch && ((d >> 32) & 3) != 0;
                } while (!__all_sync(0xffffffffu, ready));
                const unsigned has_prefix = __ballot_sync(0xffffffffu, ((d >> 32) & 3) == 2);
                const int first = has_prefix ? __ffs(has_prefix) - 1 : 31;
                int part = lane <= first ? (int)(unsigned)d : 0;
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                prefix += part;
                if (has_prefix) break;
            }
            if (lane == 0) st_release_u64(desc + tile, tag | (2ull << 32) | (unsigned)(prefix + total));
        }
        if (lane == 0) {
            s_prefix = prefix;
            if (tile == n_tiles - 1) { data[n] = prefix + total; *counter = 0u; }   // every tile has taken its index by now
        }
    }
    __syncthreads();
    off += s_prefix;
    int o[kScanItems];
    #pragma unroll
    for (int k = 0; k < kScanItems; ++k) { o[k] = off; off += v[k]; }
    if (whole) {
        *reinterpret_cast<int4 *>(data + base) = make_int4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<int4 *>(data + base + 4) = make_int4(o[4], o[5], o[6], o[7]);
    } else {
        #pragma unroll
        for (int k = 0; k < kScanItems; ++k) if (base + k < n) data[base + k] = o[k];
    }
}

// counts in data[0..n) -> exclusive offsets in data[0..n], data[n] = total; data1 (optional): a second array of the same length
inline int scan_exclusive(orbc_ctx *c, int *data, int n, int *data1 = nullptr) {
    if (n <= 0) return ORBC_OK;
    const int n_tiles = (n + kScanTile - 1) / kScanTile;
    // the two tile counters + descriptors (2 ints each, two sets) live in scan_tmp; a new buffer starts zeroed (epoch 0 is never used)
    const size_t need = 4 * (size_t)n_tiles + 4;
    if (c->scan_tmp_cap < need) {
        const size_t cap = need < 65540 ? 65540 : need;
        ORBC_TRY(dev_alloc(&c->scan_tmp, cap)); c->scan_tmp_cap = cap;
        ORBC_CUDA(cudaMemsetAsync(c->scan_tmp, 0, sizeof(int) * cap, c->stream));
        c->scan_epoch = 0;
    }
    if (++c->scan_epoch >= (1u << 30)) {
        ORBC_CUDA(cudaMemsetAsync(c->scan_tmp, 0, sizeof(int) * c->scan_tmp_cap, c->stream));
        c->scan_epoch = 1;
    }
    unsigned *counter = (unsigned *)c->scan_tmp;
    unsigned long long *desc = (unsigned long long *)(c->scan_tmp + 2);
    ORBC_LAUNCH(c, k_scan_onepass, data1 ? 2 * n_tiles : n_tiles, kScanThreads, 0, data, data1, n, n_tiles, desc, counter, c->scan_epoch);
    return ORBC_OK;
}

// ------------------------------------------------------------------------------------------------
// Stable LSD radix sort of (uint32 key, int value), 8 bits per pass.  One warp owns one tile of kRadixTile keys and
// walks it 32 keys at a time; __match_any_sync gives each key its rank among equal digits of the same row, a per-warp
// shared-memory counter carries the running count across rows, so equal keys keep their input order (stable).
// ------------------------------------------------------------------------------------------------
constexpr int kRadixTile = 1024;
constexpr int kRadixWarps = 4;

__global__ void __launch_bounds__(kRadixWarps * 32) k_radix_hist(const uint32_t *__restrict__ keys, int n, int shift, int n_tiles, int *__restrict__ hist) {
    __shared__ int cnt[kRadixWarps][256];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * kRadixWarps + w;
    for (int d = lane; d < 256; d += 32) cnt[w][d] = 0;
    __syncwarp();
    if (tile < n_tiles) {
        const int beg = tile * kRadixTile, end = min(n, beg + kRadixTile);
        for (int i = beg + lane; i < end; i += 32) atomicAdd(&cnt[w][(keys[i] >> shift) & 255], 1);
    }
    __syncwarp();
    if (tile < n_tiles) for (int d = lane; d < 256; d += 32) hist[d * n_tiles + tile] = cnt[w][d];
}

__global__ void __launch_bounds__(kRadixWarps * 32) k_radix_scatter(const uint32_t *__restrict__ keys_in, const int *__restrict__ vals_in,
                                                                     uint32_t *__restrict__ keys_out, int *__restrict__ vals_out,
                                                                     int n, int shift, int n_tiles, const int *__restrict__ hist) {
    __shared__ int off[kRadixWarps][256];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * kRadixWarps + w;
    if (tile >= n_tiles) return;
    for (int d = lane; d < 256; d += 32) off[w][d] = hist[d * n_tiles + tile];
    __syncwarp();
    const int beg = tile * kRadixTile, end = min(n, beg + kRadixTile);
    for (int row = beg; row < end; row += 32) {
        const int i = row + lane;
        const bool live = i < end;
        const uint32_t key = live ? keys_in[i] : 0xffffffffu;
        const int val = live ? vals_in[i] : 0;
        const unsigned digit = live ? (key >> shift) & 255u : 256u + lane; // dead lanes never match a live digit
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        int base = 0;
        if (live) base = off[w][digit];
        __syncwarp();
        if (live && rank == 0) off[w][digit] = base + __popc(peers);
        __syncwarp();
        if (live) { keys_out[base + rank] = key; vals_out[base + rank] = val; }
    }
}

// sorts (keys, vals) ascending by key, stable; result ends up back in keys/vals (4 passes = even number of swaps)
inline int radix_sort_pairs(orbc_ctx *c, uint32_t *keys, int *vals, uint32_t *keys_tmp, int *vals_tmp, int n) {
    if (n <= 1) return ORBC_OK;
    const int n_tiles = (n + kRadixTile - 1) / kRadixTile;
    const size_t hist_n = (size_t)256 * n_tiles;
    if (c->radix_hist_cap < hist_n + 1) { ORBC_TRY(dev_alloc(&c->radix_hist, hist_n + 1)); c->radix_hist_cap = hist_n + 1; }
    const int blocks = (n_tiles + kRadixWarps - 1) / kRadixWarps;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 8 * pass;
        ORBC_LAUNCH(c, k_radix_hist, blocks, kRadixWarps * 32, 0, keys, n, shift, n_tiles, c->radix_hist);
        ORBC_TRY(scan_exclusive(c, c->radix_hist, (int)hist_n));
        ORBC_LAUNCH(c, k_radix_scatter, blocks, kRadixWarps * 32, 0, keys, vals, keys_tmp, vals_tmp, n, shift, n_tiles, c->radix_hist);
        uint32_t *tk = keys; keys = keys_tmp; keys_tmp = tk;
        int *tv = vals; vals = vals_tmp; vals_tmp = tv;
    }
    return ORBC_OK;
}

} // namespace orbc
