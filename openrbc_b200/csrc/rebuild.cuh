// rebuild.cuh — the spatial index on the device: Voronoi centroids, Morton order of the centroids, the uniform grid that
// stands in for the reference's k-d tree, per-cell centroid stencils (r < 9 / 8 / 6), nearest-centroid partition of the
// particles, stable counting sort into cells and the gather-reorder.
//
// Replaces (results identical, integer outputs bit-exact): voronoi.h:77-86,105-117,123-140,153-237, kdtree.h:116-285,
// reorder.h:73-149, reorder_morton.h:25-122, container.h:39-58, cleanup.h:29-91.
//
// All floating-point expressions that decide an integer outcome use the _rn intrinsics in the reference's operation
// order (no FMA contraction), so centroids, Morton keys, nearest-centroid choices and stencil sets match the
// reference's strict build bit for bit.
#pragma once
#include "common.cuh"
#include "primitives.cuh"

namespace orbc {

struct GridDev {
    float lox, loy, loz, inv_h, h;
    int dx, dy, dz;
    const int *bin_start;
    const int *bin_items;
    const float4 *sorted;     // centroids in bin order: (x, y, z, cell id as int bits)
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ void grid_bin(const GridDev &g, float4 p, int &bx, int &by, int &bz) {
    bx = clampi((int)floorf((p.x - g.lox) * g.inv_h), 0, g.dx - 1);
    by = clampi((int)floorf((p.y - g.loy) * g.inv_h), 0, g.dy - 1);
    bz = clampi((int)floorf((p.z - g.loz) * g.inv_h), 0, g.dz - 1);
}

// Static partition of the Voronoi cells over the ranks of a decomposed run: rank g owns cells [beg[g], beg[g + 1]),
// beg[g] = g * n_cells / world — the reference's own thread partition (util_numa.h:41-42).  world == 1: everything.
struct CellOwners {
    int beg[kMaxWorld + 1];
    int world;
    __host__ __device__ int owner(int c) const {
        int g = 0;
        #pragma unroll
        for (int r = 1; r < kMaxWorld; ++r) g += (r < world && c >= beg[r]);
        return g;
    }
};
struct CentroidOut { float4 *dst[kMaxWorld]; int world; };

// ---- voronoi.h:123-140 — centroid = fp32 sequential sum over the cell's slots, times 1/count --------------------------
// cells [cb, ce) (the owned ones); the result goes to every rank's copy of the centroid array
// `map` (VoronoiDiagram::init, voronoi.h:69: the members of a cell through VCellList::cells while the container is not reordered) or null
__global__ void k_centroid_update(const int *__restrict__ cell_start, const float4 *__restrict__ x, int cb, int ce, CentroidOut out, const int *__restrict__ map) {
    const int i = cb + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ce) return;
    const int b = cell_start[i], e = cell_start[i + 1];
    float cx = 0.f, cy = 0.f, cz = 0.f;
    for (int j = b; j < e; ++j) {
        const float4 p = x[map ? map[j] : j];
        cx = __fadd_rn(cx, p.x); cy = __fadd_rn(cy, p.y); cz = __fadd_rn(cz, p.z);
    }
    const float s = __fdiv_rn(1.0f, __int2float_rn(e - b));   // empty cell: inf -> NaN centroid, as in the reference
    const float4 r = make_float4(__fmul_rn(cx, s), __fmul_rn(cy, s), __fmul_rn(cz, s), 0.f);
    for (int g = 0; g < out.world; ++g) out.dst[g][i] = r;
}

// ---- reorder_morton.h:25-42 ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bit_space3(uint32_t x) {
    x = (x | (x << 12)) & 0X00FC003FU;
    x = (x | (x << 6)) & 0X381C0E07U;
    x = (x | (x << 4)) & 0X190C8643U;
    x = (x | (x << 2)) & 0X49249249U;
    return x;
}
__device__ __forceinline__ uint32_t morton_axis(float x) {
    // static_cast<unsigned>(2 * x + bsize): 2*x in fp32, the sum in fp64 (bsize = 2000.0 is a double, runtime_parameter.h:46)
    return (uint32_t)__double2uint_rz(__dadd_rn((double)__fmul_rn(2.0f, x), 2000.0));
}
__device__ __forceinline__ uint32_t morton_encode(float4 p) {
    return bit_space3(morton_axis(p.x)) | (bit_space3(morton_axis(p.y)) << 1) | (bit_space3(morton_axis(p.z)) << 2);
}
__global__ void k_morton_keys(const float4 *__restrict__ pts, int n, uint32_t *__restrict__ keys, int *__restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = morton_encode(pts[i]);
    if (idx) idx[i] = i;
}
__global__ void k_permute_centroids(const float4 *__restrict__ src, const int *__restrict__ perm, int n, float4 *__restrict__ dst, int *__restrict__ inv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int o = perm[i];
    dst[i] = src[o];
    inv[o] = i;
}
__global__ void k_remap_cellid(int *__restrict__ cellid, const int *__restrict__ range, const int *__restrict__ inv) {
    const int i = range[0] + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= range[1]) return;
    const int c = cellid[i];
    if (c >= 0) cellid[i] = inv[c];
}
__global__ void k_fill_cellid(const int *__restrict__ cell_start, int n_cells, int *__restrict__ cellid) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    for (int j = cell_start[c]; j < cell_start[c + 1]; ++j) cellid[j] = c;
}
__global__ void k_fill_int(int *__restrict__ a, size_t n, int v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

// ---- uniform grid over the centroids (counting sort into bins) --------------------------------------------------------------
__global__ void k_bin_count(const float4 *__restrict__ centroid, int n, GridDev g, int *__restrict__ bin_cnt, int *__restrict__ bin_of, int *__restrict__ bin_slot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = centroid[i];
    if (!(p.x == p.x) || !(p.y == p.y) || !(p.z == p.z)) { bin_of[i] = -1; return; }   // NaN centroid of an empty cell
    int bx, by, bz; grid_bin(g, p, bx, by, bz);
    const int b = (bz * g.dy + by) * g.dx + bx;
    bin_of[i] = b;
    bin_slot[i] = atomicAdd(&bin_cnt[b], 1);
}
__global__ void k_bin_fill(int n, const int *__restrict__ bin_start, const int *__restrict__ bin_of, const int *__restrict__ bin_slot, int *__restrict__ bin_items,
                           const float4 *__restrict__ centroid, float4 *__restrict__ sorted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = bin_of[i];
    if (b < 0) return;
    const int k = bin_start[b] + bin_slot[i];
    bin_items[k] = i;
    const float4 p = centroid[i];
    sorted[k] = make_float4(p.x, p.y, p.z, __int_as_float(i));
}

// ---- per-cell centroid stencils: {c2 : |c2 - c1|^2 < 81}, classed by < 36 / < 64 / < 81, ordered (class, id) -------------
// voronoi.h:105-117 (get_stencil_whole / refine_stencil) and kdtree.h:263-285 (find_within); the sets are identical, the
// reference's order is tree-traversal order.  One warp per cell; the 27 surrounding bins are 9 x-contiguous runs.
constexpr int kStencilWarps = 8;
// Decomposed runs also derive, per cell, the set of OTHER ranks that own a member of its r<9 stencil (dest_mask: who needs
// this cell's particles as halo) and mark the cells this rank reads (need[c2] = need_epoch for the stencil members of owned cells).
struct HaloOut { unsigned char *dest_mask; int *need; int need_epoch; CellOwners own; int rank; };

// WIDE stencils.  Centroids creep (~0.003 per rebuild), so the r<9 stencil of a cell hardly ever changes — but it must be exact at
// every rebuild.  The full search (27 grid bins, ~80 candidates per cell) therefore also records every cell closer than 9 + kWideMargin
// together with the centroids it saw (cen_ref); the rebuilds that follow re-classify only those ~25 recorded neighbours with
// the exact squared distances (k_stencil_refresh) — the same test on the same operands, so the same sets — as long as no centroid
// has moved further than kWideMargin / 2 from its recorded position; the few that have (the movers, below) are searched in full.
// A Morton renumbering of the cells always forces the full search of all cells.
constexpr float kWideMargin = 1.0f;
struct WideOut { int *wide; int *wide_cnt; float4 *cen_ref; };    // wide == nullptr: not recorded

// classification key of a candidate: class << 28 | id, class 0: d2 < 36, 1: < 64, 2: < 81, 3: only inside the wide radius
__device__ __forceinline__ int stencil_key(float d2, int id) { return ((d2 < 36.0f ? 0 : (d2 < 64.0f ? 1 : (d2 < 81.0f ? 2 : 3))) << 28) | id; }

// common tail: the keys of one cell (in shared memory, `count` of them, any order) -> ordered (class, id) stencil row, counts,
// halo bookkeeping, and (full search only) the wide row
__device__ __forceinline__ void stencil_emit(const int *s_key, int count, int c, int lane, int *__restrict__ stencil, int *__restrict__ stencil_cnt,
                                             int *__restrict__ flags, const HaloOut &halo, const WideOut &wide) {
    int n6 = 0, n8 = 0, n9 = 0;
    unsigned owners = 0;
    const bool mine = halo.own.world > 1 && halo.own.owner(c) == halo.rank;
    for (int e = lane; e < count; e += 32) {
        const int key = s_key[e];
        const int cls = key >> 28, id = key & 0x0fffffff;
        int rank = 0, rank_id = 0;
        for (int k = 0; k < count; ++k) { const int o = s_key[k]; rank += (o < key); rank_id += ((o & 0x0fffffff) < id); }
        if (wide.wide) wide.wide[(size_t)c * kStencilStride + rank_id] = id;       // the wide row in ascending id (k_stencil_refresh ranks by lane order)
        if (cls <= 2) {
            stencil[(size_t)c * kStencilStride + rank] = id;
            n6 += cls == 0; n8 += cls <= 1; ++n9;
            if (halo.own.world > 1) {
                owners |= 1u << halo.own.owner(id);
                if (mine) halo.need[id] = halo.need_epoch;
            }
        }
    }
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) { n6 += __shfl_xor_sync(0xffffffffu, n6, d); n8 += __shfl_xor_sync(0xffffffffu, n8, d); n9 += __shfl_xor_sync(0xffffffffu, n9, d); }
    if (lane == 0) {
        stencil_cnt[c] = n6 | (n8 << 8) | (n9 << 16);
        if (wide.wide) wide.wide_cnt[c] = count;
        if (n6 > 32) atomicExch(&flags[0], c + 1);               // the lipid kernels keep the r<6 cells of a cell in 32 slots
    }
    if (halo.own.world > 1) {
        owners = __reduce_or_sync(0xffffffffu, owners);
        if (lane == 0) { halo.dest_mask[c] = (unsigned char)(owners & ~(1u << halo.own.owner(c))); if (mine) halo.need[c] = halo.need_epoch; }
    }
}

// the candidates of one cell: the 27 grid bins around q, every centroid closer than sqrt(lim) as a key in s_key (any order); returns
// their number (which may exceed the row: the caller flags that).  One warp.
__device__ __forceinline__ int stencil_search(const float4 q, const GridDev &g, float lim, int *__restrict__ s_key, int lane) {
    int count = 0;
    if (!(q.x == q.x && q.y == q.y && q.z == q.z)) return 0;      // an empty cell
    int bx, by, bz; grid_bin(g, q, bx, by, bz);
    const int x0 = max(bx - 1, 0), x1 = min(bx + 1, g.dx - 1);
    // lanes 0..8 fetch the nine x-runs together, a warp scan turns them into one flat candidate list
    int beg = 0, cnt = 0;
    if (lane < 9) {
        const int zz = bz - 1 + lane / 3, yy = by - 1 + lane % 3;
        if (zz >= 0 && zz < g.dz && yy >= 0 && yy < g.dy) {
            const int row = (zz * g.dy + yy) * g.dx;
            beg = g.bin_start[row + x0];
            cnt = g.bin_start[row + x1 + 1] - beg;
        }
    }
    int incl = cnt;
    #pragma unroll
    for (int d = 1; d < 16; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
    const int total = __shfl_sync(0xffffffffu, incl, 8);
    int rb[9], re[9];                                             // run r covers flat slots [re[r] - cnt_r, re[r])
    #pragma unroll
    for (int r = 0; r < 9; ++r) { rb[r] = __shfl_sync(0xffffffffu, beg, r); re[r] = __shfl_sync(0xffffffffu, incl, r); }
    for (int s0 = 0; s0 < total; s0 += 32) {
        const int s = s0 + lane;
        int key = -1;
        if (s < total) {
            int idx = rb[0] + s;
            #pragma unroll
            for (int r = 1; r < 9; ++r) if (s >= re[r - 1]) idx = rb[r] + (s - re[r - 1]);
            const float4 p = g.sorted[idx];
            const float d2 = dist2_rn(p, q);                      // normsq(pts_[i] - q), kdtree.h:274
            if (d2 < lim) key = stencil_key(d2, __float_as_int(p.w));
        }
        const unsigned m = __ballot_sync(0xffffffffu, key >= 0);
        const int pos = count + __popc(m & ((1u << lane) - 1u));
        if (key >= 0 && pos < kStencilStride) s_key[pos] = key;
        count += __popc(m);
    }
    return count;
}

// the full search.  `gate`: run only if *gate == 0 (the refresh could not be trusted); grid-stride, so that a gated launch is cheap.
// With wide rows it also notes where EVERY centroid stood (cen_ref): the displacements are measured from this moment on.
__global__ void __launch_bounds__(kStencilWarps * 32) k_stencil_build(const float4 *__restrict__ centroid, int n_cells, int c_beg, int c_end, GridDev g,
                                                                       int *__restrict__ stencil, int *__restrict__ stencil_cnt, int *__restrict__ flags, HaloOut halo,
                                                                       WideOut wide, const int *__restrict__ gate, unsigned long long *__restrict__ counters) {
    if (gate && *gate != 0) return;
    if (gate && blockIdx.x == 0 && threadIdx.x == 0) counters[7] += 1ull << 40;             // statistics: refreshes that were not trusted (high part)
    __shared__ int s_key[kStencilWarps][kStencilStride];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float lim = wide.wide ? (9.0f + kWideMargin) * (9.0f + kWideMargin) : 81.0f;
    if (wide.wide)                                               // (the cells of other ranks; the searched ones are noted below)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += gridDim.x * blockDim.x)
            if (i < c_beg || i >= c_end) wide.cen_ref[i] = centroid[i];
    for (int c = c_beg + blockIdx.x * kStencilWarps + w; c < c_end; c += gridDim.x * kStencilWarps) {
        const float4 q = centroid[c];
        int count = stencil_search(q, g, lim, s_key[w], lane);
        if (count > kStencilStride) { if (lane == 0) atomicExch(&flags[0], c + 1); count = kStencilStride; }
        __syncwarp();
        if (wide.wide && lane == 0) wide.cen_ref[c] = q;
        stencil_emit(s_key[w], count, c, lane, stencil, stencil_cnt, flags, halo, wide);
        __syncwarp();
    }
}

// Has a centroid stayed within kWideMargin / 2 of where it stood when the wide rows were recorded?  (Two cells that both have can
// not have come closer than 9 without the one being in the other's wide row.)  An empty cell stays a NaN on both sides.
__device__ __forceinline__ bool centroid_settled(const float4 a, const float4 b) {
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    const float lim = 0.5f * kWideMargin - 1e-3f;
    return (dx * dx + dy * dy + dz * dz <= lim * lim) || (!(a.x == a.x) && !(b.x == b.x));
}
// The MOVERS: the cells that have not.  ~70 of the 188 549 cells of the RBC per rebuild: small cells whose centroid jumps by 0.5 .. 1
// when a member arrives or leaves.  They are listed (k_centroid_disp), searched in full (k_stencil_movers) and stay movers
// until the next full search of all cells; only a list that overflows condemns the refresh as a whole (*ok = 0).
struct Movers { int *list; int *count; int cap; };
__global__ void k_centroid_disp(const float4 *__restrict__ centroid, const float4 *__restrict__ cen_ref, int n, int *__restrict__ ok, Movers mv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (centroid_settled(centroid[i], cen_ref[i])) return;
    const int slot = atomicAdd(mv.count, 1);
    if (slot < mv.cap) mv.list[slot] = i; else atomicExch(ok, 0);
}

// After the refresh: every mover (of any rank: all ranks hold all centroids) is searched in full; the movers of this rank's cells
// get their exact stencil row from it.  The relation is symmetric (the same squared distance on both sides), so every cell c2 of this
// rank in a mover's stencil must have tested the mover in its refresh: it has if the mover is in its wide row -- or c2 is a mover
// itself and searches in full.  If neither holds (the mover came from beyond 9 + kWideMargin: a few cells per rebuild on the RBC) c2
// goes on a second list, `patch`, and a second launch of this kernel (VERIFY = false) searches those in full as well.  Only lists
// that overflow raise *ok = 0, and the full search of all cells, which stands by behind that flag, runs.
template <bool VERIFY>
__global__ void __launch_bounds__(kStencilWarps * 32) k_stencil_movers(const float4 *__restrict__ centroid, const float4 *__restrict__ cen_ref, int c_beg, int c_end, GridDev g,
                                                                        int *__restrict__ stencil, int *__restrict__ stencil_cnt, int *__restrict__ flags, HaloOut halo,
                                                                        const int *__restrict__ wide_rows, const int *__restrict__ wide_cnt, Movers mv, Movers patch,
                                                                        int *__restrict__ ok, unsigned long long *__restrict__ counters) {
    if (*ok == 0) return;
    __shared__ int s_key[kStencilWarps][kStencilStride];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = min(*mv.count, mv.cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[7] += (unsigned long long)n;      // statistics: cells searched in full (low 40 bits)
    for (int k = blockIdx.x * kStencilWarps + w; k < n; k += gridDim.x * kStencilWarps) {
        const int m = mv.list[k];
        int count = stencil_search(centroid[m], g, 81.0f, s_key[w], lane);
        if (count > kStencilStride) { if (lane == 0) atomicExch(&flags[0], m + 1); count = kStencilStride; }
        __syncwarp();
        if (m >= c_beg && m < c_end) stencil_emit(s_key[w], count, m, lane, stencil, stencil_cnt, flags, halo, WideOut{nullptr, nullptr, nullptr});
        if (VERIFY) {
            for (int e = lane; e < count; e += 32) {
                const int c2 = s_key[w][e] & 0x0fffffff;
                if (c2 == m || c2 < c_beg || c2 >= c_end) continue;
                const int *row = wide_rows + (size_t)c2 * kStencilStride;      // ascending ids
                const int len = min(wide_cnt[c2], kStencilStride);
                int lo = 0, hi = len;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (row[mid] < m) lo = mid + 1; else hi = mid; }
                if (!(lo < len && row[lo] == m) && centroid_settled(centroid[c2], cen_ref[c2])) {
                    const int slot = atomicAdd(patch.count, 1);
                    if (slot < patch.cap) patch.list[slot] = c2; else atomicExch(ok, 0);
                }
            }
        }
        __syncwarp();
    }
}

// the refresh: the recorded neighbours of every cell re-classified with the exact squared distances.  `gate`: run only if *gate != 0.
// The wide row is stored in ascending id, so inside one class lane order IS id order: the slot of an entry in the (class, id) ordered
// stencil row is (entries of lower classes) + (entries of its class in lower lanes) — a few ballots instead of a sort.
__global__ void __launch_bounds__(kStencilWarps * 32) k_stencil_refresh(const float4 *__restrict__ centroid, int c_beg, int c_end, const int *__restrict__ wide_rows,
                                                                         const int *__restrict__ wide_cnt, int *__restrict__ stencil, int *__restrict__ stencil_cnt,
                                                                         int *__restrict__ flags, HaloOut halo, const int *__restrict__ gate) {
    if (gate && *gate == 0) return;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = c_beg + blockIdx.x * kStencilWarps + w;
    if (c >= c_end) return;
    const float4 q = centroid[c];
    const int n = min(wide_cnt[c], kStencilStride);
    // two entries per lane (rows hold at most 64): e = lane and e = lane + 32
    int id[2], cls[2];
    #pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int e = lane + 32 * h;
        id[h] = 0; cls[h] = 3;
        if (e < n) {
            id[h] = wide_rows[(size_t)c * kStencilStride + e];
            const float d2 = dist2_rn(centroid[id[h]], q);
            cls[h] = d2 < 36.0f ? 0 : (d2 < 64.0f ? 1 : (d2 < 81.0f ? 2 : 3));    // (a NaN distance: class 3, not a member)
        }
    }
    unsigned m[2][3];
    #pragma unroll
    for (int h = 0; h < 2; ++h)
        #pragma unroll
        for (int k = 0; k < 3; ++k) m[h][k] = __ballot_sync(0xffffffffu, cls[h] == k);
    const int n0 = __popc(m[0][0]) + __popc(m[1][0]), n1 = __popc(m[0][1]) + __popc(m[1][1]), n2 = __popc(m[0][2]) + __popc(m[1][2]);
    const unsigned lt = (1u << lane) - 1u;
    unsigned owners = 0;
    const bool mine = halo.own.world > 1 && halo.own.owner(c) == halo.rank;
    #pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (cls[h] > 2) continue;
        const int k = cls[h];
        const int below = k == 0 ? 0 : (k == 1 ? n0 : n0 + n1);
        const unsigned same0 = k == 0 ? m[0][0] : (k == 1 ? m[0][1] : m[0][2]), same1 = k == 0 ? m[1][0] : (k == 1 ? m[1][1] : m[1][2]);
        const int slot = below + (h == 0 ? __popc(same0 & lt) : __popc(same0) + __popc(same1 & lt));
        stencil[(size_t)c * kStencilStride + slot] = id[h];
        if (halo.own.world > 1) {
            owners |= 1u << halo.own.owner(id[h]);
            if (mine) halo.need[id[h]] = halo.need_epoch;
        }
    }
    if (lane == 0) {
        stencil_cnt[c] = n0 | ((n0 + n1) << 8) | ((n0 + n1 + n2) << 16);
        if (n0 > 32) atomicExch(&flags[0], c + 1);
    }
    if (halo.own.world > 1) {
        owners = __reduce_or_sync(0xffffffffu, owners);
        if (lane == 0) { halo.dest_mask[c] = (unsigned char)(owners & ~(1u << halo.own.owner(c))); if (mine) halo.need[c] = halo.need_epoch; }
    }
}

// ---- nearest centroid (voronoi.h:179-216 + kdtree.h:206-236) ------------------------------------------------------------------
// exact fallback: expanding shells of grid bins around the particle until the best distance is covered
__device__ int nearest_by_grid(float4 p, const GridDev &g, const float4 *__restrict__ centroid, float &best_out) {
    int bx, by, bz; grid_bin(g, p, bx, by, bz);
    const int maxring = max(g.dx, max(g.dy, g.dz));
    float best = INFINITY; int bi = -1;
    for (int k = 0; k <= maxring; ++k) {
        for (int zz = bz - k; zz <= bz + k; ++zz) {
            if (zz < 0 || zz >= g.dz) continue;
            for (int yy = by - k; yy <= by + k; ++yy) {
                if (yy < 0 || yy >= g.dy) continue;
                const bool face = (abs(zz - bz) == k) || (abs(yy - by) == k);
                const int step = face ? 1 : 2 * k;                      // interior rows: only the two end bins belong to shell k
                for (int xx = bx - k; xx <= bx + k; xx += (step > 0 ? step : 1)) {
                    if (xx < 0 || xx >= g.dx) continue;
                    const int b = (zz * g.dy + yy) * g.dx + xx;
                    for (int q = g.bin_start[b]; q < g.bin_start[b + 1]; ++q) {
                        const int j = g.bin_items[q];
                        const float d2 = dist2_rn(p, centroid[j]);
                        if (d2 < best || (d2 == best && j < bi)) { best = d2; bi = j; }
                    }
                }
            }
        }
        const float reach = (float)k * g.h;       // every bin outside shells 0..k is at least k*h away from p
        if (bi >= 0 && best <= reach * reach) break;
    }
    best_out = best;
    return bi;
}

// One thread per particle.  Fast path: the particle's previous cell g and g's r<9 centroid stencil; it is exact whenever
// d(best) + d(g) < 9 (every centroid at least that close to the particle is then inside the stencil).  Otherwise the grid
// search above.  Ties in the squared distance go to the lower cell id.
// (the per-container arguments and the ones both containers share: k_assign_nearest2 runs both containers in one launch)
struct AssignArgs { const float4 *x; const int *cellid; const int *range; int *aff; int *li; int *cell_cnt; const int *keep; };
struct AssignCommon { const float4 *centroid; int n_cells; const int *stencil; const int *stencil_cnt; GridDev g; unsigned long long *counters; int *flags; };
__device__ __forceinline__ void assign_nearest_body(const unsigned bid, const AssignArgs &a, const AssignCommon &s) {
    const float4 *__restrict__ x = a.x; const int *__restrict__ cellid = a.cellid; const int *__restrict__ range = a.range;
    int *__restrict__ aff = a.aff; int *__restrict__ li = a.li; int *__restrict__ cell_cnt = a.cell_cnt; const int *__restrict__ keep = a.keep;
    const float4 *__restrict__ centroid = s.centroid; const int n_cells = s.n_cells; const int *__restrict__ stencil = s.stencil;
    const int *__restrict__ stencil_cnt = s.stencil_cnt; const GridDev &g = s.g; unsigned long long *__restrict__ counters = s.counters; int *__restrict__ flags = s.flags;
    const int i = range[0] + bid * blockDim.x + threadIdx.x;
    const bool live = i < range[1];
    const bool gone = live && keep && !keep[i];                  // stray lipid being deleted (cleanup.h:29-91): it joins no cell
    int bi = -1;
    if (live && !gone) {
    const float4 p = x[i];
    const int guess = cellid ? cellid[i] : -1;
    float best = INFINITY; bool ok = false;
    if (guess >= 0 && guess < n_cells) {
        // candidates = the r<6 stencil of the previous cell; exact whenever d(best) + d(guess) < 6 (every centroid at least
        // as close to the particle as `best` is then closer than 6 to the guess, i.e. inside that stencil).  Otherwise the
        // same argument with the r<9 stencil, then the grid.
        const int cnt = stencil_cnt[guess];
        const int n6 = cnt & 255, n9 = cnt >> 16;
        const int *st = stencil + (size_t)guess * kStencilStride;
        for (int k = 0; k < n6; ++k) {
            const int j = st[k];
            const float d2 = dist2_rn(p, centroid[j]);
            if (d2 < best || (d2 == best && j < bi)) { best = d2; bi = j; }
        }
        const float dg = sqrtf(dist2_rn(p, centroid[guess]));
        if (bi >= 0) ok = sqrtf(best) + dg < 6.0f - 1e-3f;
        if (!ok) {
            for (int k = n6; k < n9; ++k) {
                const int j = st[k];
                const float d2 = dist2_rn(p, centroid[j]);
                if (d2 < best || (d2 == best && j < bi)) { best = d2; bi = j; }
            }
            if (bi >= 0) ok = sqrtf(best) + dg < 9.0f - 1e-3f;
        }
    }
    if (!ok) {
        bi = nearest_by_grid(p, g, centroid, best);
        atomicAdd(&counters[0], 1ULL);
    }
    if (bi < 0) { atomicExch(&flags[1], (int)(i & 0x7fffffff) + 1); bi = 0; }   // NaN position: keep the structure consistent, report
    }
    if (gone) aff[i] = -1;
    // arrival slots: the lanes of a warp that chose the same cell (neighbours in storage order mostly do) take ONE atomic between
    // them — the arrival order is arbitrary anyway, k_rank_and_move sorts every cell's list by old index
    const unsigned peers = __match_any_sync(0xffffffffu, bi);
    if (bi >= 0) {
        const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
        int base = 0;
        if (lane == leader) base = atomicAdd(&cell_cnt[bi], __popc(peers));
        base = __shfl_sync(peers, base, leader);
        aff[i] = bi;
        li[i] = base + __popc(peers & ((1u << lane) - 1u));
    }
}

__global__ void k_assign_nearest(AssignArgs a, AssignCommon s) { assign_nearest_body(blockIdx.x, a, s); }
// both containers: blocks [0, blocks0) take a0's particles, the others a1's (a block never mixes the two)
__global__ void k_assign_nearest2(AssignArgs a0, AssignArgs a1, unsigned blocks0, AssignCommon s) {
    if (blockIdx.x < blocks0) assign_nearest_body(blockIdx.x, a0, s); else assign_nearest_body(blockIdx.x - blocks0, a1, s);
}

// cells_tmp[cell_start[aff] + arrival slot] = i   (voronoi.h:228-231 with an arbitrary arrival order ...)
// (decomposed: this rank's members of cell c take the slots [cell_start[c] + off_me[c], + cnt_me[c]) of the arrival list)
struct ScatterArgs { const int *aff; const int *li; const int *range; const int *cell_start; const int *off_me; int *cells_tmp; };
__device__ __forceinline__ void cell_scatter_body(const unsigned bid, const ScatterArgs &a) {
    const int i = a.range[0] + bid * blockDim.x + threadIdx.x;
    if (i >= a.range[1]) return;
    const int c = a.aff[i];
    if (c < 0) return;
    a.cells_tmp[a.cell_start[c] + (a.off_me ? a.off_me[c] : 0) + a.li[i]] = i;
}
__global__ void k_cell_scatter(ScatterArgs a) { cell_scatter_body(blockIdx.x, a); }
__global__ void k_cell_scatter2(ScatterArgs a0, ScatterArgs a1, unsigned blocks0) {
    if (blockIdx.x < blocks0) cell_scatter_body(blockIdx.x, a0); else cell_scatter_body(blockIdx.x - blocks0, a1);
}
// Sorting every cell's arrival list and the gather-reorder in one pass, one thread per particle: the slot of particle i inside its new cell is
// the number of members with a lower index (= arrival order of the reference at one thread, voronoi.h:214-215), found by
// scanning the cell's arrival list (a few dozen L1-resident ints shared by neighbouring threads); the particle then moves
// itself: new[cell_start + rank] = old[i]  (reorder.h:73-149 as a scatter instead of a gather; `cells` still records the
// permutation, VCellList::cells).
// On a decomposed run the arrival list holds this rank's members only (cnt_me of them per cell), members that come from lower
// ranks precede them in the cell (off_me; ranks own ascending slot ranges, so this is still ascending old index), and the
// destination is the array of whichever rank owns the new cell — a direct store into that GPU's memory over NVLink.  A
// protein also announces its new slot to every rank's tag -> index map (container.h:39-58).
struct MoveDst {
    float4 *x[kMaxWorld], *nn[kMaxWorld], *v[kMaxWorld], *o[kMaxWorld];
    int *cellid[kMaxWorld], *tag2idx[kMaxWorld];
    CellOwners own;
};
struct MoveArgs {
    const int *aff, *range, *cell_start, *cnt_me, *off_me, *cells_tmp; int *cells;
    const float4 *x0, *n0, *v0, *o0;
    MoveDst d; int announce_tags;
};
__device__ __forceinline__ void rank_and_move_body(const unsigned bid, const MoveArgs &a) {
    const int *__restrict__ aff = a.aff; const int *__restrict__ range = a.range; const int *__restrict__ cell_start = a.cell_start;
    const int *__restrict__ cnt_me = a.cnt_me; const int *__restrict__ off_me = a.off_me; const int *__restrict__ cells_tmp = a.cells_tmp; int *__restrict__ cells = a.cells;
    const float4 *__restrict__ x0 = a.x0; const float4 *__restrict__ n0 = a.n0; const float4 *__restrict__ v0 = a.v0; const float4 *__restrict__ o0 = a.o0;
    const MoveDst &d = a.d; const int announce_tags = a.announce_tags;
    const int i = range[0] + bid * blockDim.x + threadIdx.x;
    if (i >= range[1]) return;
    const int c = aff[i];
    if (c < 0) return;
    const int b = cell_start[c] + (off_me ? off_me[c] : 0), e = cnt_me ? b + cnt_me[c] : cell_start[c + 1];
    int rank = 0;
    for (int k = b; k < e; ++k) rank += (cells_tmp[k] < i);
    const int j = b + rank;
    const int g = d.own.owner(c);
    cells[j] = i;
    const float4 nn = n0[i];
    d.x[g][j] = x0[i]; d.nn[g][j] = nn; d.v[g][j] = v0[i]; d.o[g][j] = o0[i];
    d.cellid[g][j] = c;
    if (announce_tags) {
        const int tag = __float_as_int(nn.w);
        for (int r = 0; r < d.own.world; ++r) d.tag2idx[r][tag] = j;
    }
}

__global__ void k_rank_and_move(MoveArgs a) { rank_and_move_body(blockIdx.x, a); }
__global__ void k_rank_and_move2(MoveArgs a0, MoveArgs a1, unsigned blocks0) {
    if (blockIdx.x < blocks0) rank_and_move_body(blockIdx.x, a0); else rank_and_move_body(blockIdx.x - blocks0, a1);
}

// the sorting half of k_rank_and_move alone: cells[] = the members of every cell in ascending index (VCellList::cells after
// partition(), voronoi.h:228-231), the container stays where it is (VoronoiDiagram::init reorders only at some iterations)
__global__ void k_rank_only(const int *__restrict__ aff, const int *__restrict__ range, const int *__restrict__ cell_start, const int *__restrict__ cells_tmp, int *__restrict__ cells) {
    const int i = range[0] + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= range[1]) return;
    const int c = aff[i];
    if (c < 0) return;
    const int b = cell_start[c], e = cell_start[c + 1];
    int rank = 0;
    for (int k = b; k < e; ++k) rank += (cells_tmp[k] < i);
    cells[b + rank] = i;
}

// VoronoiDiagram::init, voronoi.h:57-62: the initial guess is every delta-th particle, starting at delta / 2, wrapping around
__global__ void k_init_centroids(const float4 *__restrict__ x, size_t n, int n_cells, size_t delta, float4 *__restrict__ centroid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells) return;
    const float4 p = x[(delta / 2 + (size_t)i * delta) % n];
    centroid[i] = make_float4(p.x, p.y, p.z, 0.f);
}
// bounding box of the positions as order-preserving integer keys: bb[0..2] = min, bb[3..5] = max
__device__ __forceinline__ int float_key(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__global__ void k_bbox(const float4 *__restrict__ x, size_t n, int *__restrict__ bb) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    if (i < n) {
        const float4 p = x[i];
        if (p.x == p.x && p.y == p.y && p.z == p.z) { lo[0] = hi[0] = float_key(p.x); lo[1] = hi[1] = float_key(p.y); lo[2] = hi[2] = float_key(p.z); }
    }
    #pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int l = __reduce_min_sync(0xffffffffu, lo[d]), h = __reduce_max_sync(0xffffffffu, hi[d]);
        if ((threadIdx.x & 31) == 0) { atomicMin(bb + d, l); atomicMax(bb + 3 + d, h); }
    }
}

// The checks of an upload, on the device (the ids are there anyway; on the host the sweeps cost as much as the copy).  All results
// are maxima over zero-initialised words: chk[0] = n - (first slot with a type outside [0, n_type)), chk[1] = mask of the types present,
// chk[2] = the largest -tag, chk[3] = the largest tag.
__global__ void k_check_ids(const int *__restrict__ type, const int *__restrict__ tag, size_t n, int n_type, int *__restrict__ chk) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int bad = 0, neg = 0, top = 0; unsigned mask = 0;
    if (i < n) {
        if (type) { const unsigned t = (unsigned)type[i]; if (t >= (unsigned)n_type) bad = (int)(n - i); mask = 1u << (t & 31u); }
        if (tag) { const int g = tag[i]; neg = g < 0 ? (g == (int)0x80000000 ? 0x7fffffff : -g) : 0; top = max(g, 0); }
    }
    bad = __reduce_max_sync(0xffffffffu, bad); mask = __reduce_or_sync(0xffffffffu, mask);
    neg = __reduce_max_sync(0xffffffffu, neg); top = __reduce_max_sync(0xffffffffu, top);
    if ((threadIdx.x & 31) == 0) {
        // the words only grow: an atomic is needed only if this warp would still raise what it reads (17 000 warps on four words otherwise)
        if (bad > __ldcg(&chk[0])) atomicMax(&chk[0], bad);
        if (mask & ~(unsigned)__ldcg(&chk[1])) atomicOr((unsigned *)&chk[1], mask);
        if (neg > __ldcg(&chk[2])) atomicMax(&chk[2], neg);
        if (top > __ldcg(&chk[3])) atomicMax(&chk[3], top);
    }
}
// chk[4] = n - (first bond with a type outside [0, 4) or a tag outside the tag -> index map)
__global__ void k_check_bonds(const int *__restrict__ tij, size_t n, size_t map_size, int *__restrict__ chk) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int bad = 0;
    if (b < n) {
        const unsigned t = (unsigned)tij[3 * b], i = (unsigned)tij[3 * b + 1], j = (unsigned)tij[3 * b + 2];
        if (t >= 4u || (size_t)i >= map_size || (size_t)j >= map_size) bad = (int)(n - b);
    }
    bad = __reduce_max_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicMax(&chk[4], bad);
}

// container.h:39-58
__global__ void k_build_tag2idx(const float4 *__restrict__ nn, size_t n, int *__restrict__ map, size_t map_size, int *__restrict__ flags) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int tag = __float_as_int(nn[i].w);
    if (tag >= 0 && (size_t)tag < map_size) map[tag] = (int)i; else atomicExch(&flags[2], (int)i + 1);
}

// ---- cleanup.h:29-60: per cell, median squared distance to the centroid; keep = dr2 < median * tol^2 ------------------------
__global__ void k_stray_mask(const int *__restrict__ cell_start, int c_beg, int c_end, const float4 *__restrict__ centroid, const float4 *__restrict__ x,
                             float tol, int *__restrict__ keep) {
    const int c = c_beg + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (c >= c_end) return;
    const int b = cell_start[c], m = cell_start[c + 1] - b;
    if (m <= 0) return;
    const float4 q = centroid[c];
    // the element of rank m/2 in the sorted list (std::sort + dr2[size/2]); rank by (value, slot)
    float med = 0.f; int have = 0;
    for (int e = lane; e < m; e += 32) {
        const float d = dist2_rn(x[b + e], q);
        int rank = 0;
        for (int k = 0; k < m; ++k) { const float dk = dist2_rn(x[b + k], q); rank += (dk < d) || (dk == d && k < e); }
        if (rank == m / 2) { med = d; have = 1; }
    }
    const unsigned who = __ballot_sync(0xffffffffu, have);
    med = __shfl_sync(0xffffffffu, med, __ffs(who) - 1);
    const float threshold = __fmul_rn(__fmul_rn(med, tol), tol);
    for (int e = lane; e < m; e += 32) keep[b + e] = dist2_rn(x[b + e], q) < threshold ? 1 : 0;
}
// decomposed delete_lipid: lipids of the owned slots that do NOT survive (acc[0] += count); integrate.cuh's block_add_double is not
// visible here, a warp reduction + one atomic per warp does
__global__ void k_count_strays(const int *__restrict__ keep, const int *__restrict__ range, double *acc) {
    const int i = range[0] + blockIdx.x * blockDim.x + threadIdx.x;
    int gone = (i < range[1] && !keep[i]) ? 1 : 0;
    gone = __reduce_add_sync(0xffffffffu, gone);
    if ((threadIdx.x & 31) == 0 && gone) atomicAdd(acc, (double)gone);
}
__global__ void k_compact(const int *__restrict__ keep, const int *__restrict__ newpos, size_t n,
                          const float4 *__restrict__ x0, const float4 *__restrict__ n0, const float4 *__restrict__ v0, const float4 *__restrict__ o0, const int *__restrict__ c0,
                          float4 *__restrict__ x1, float4 *__restrict__ n1, float4 *__restrict__ v1, float4 *__restrict__ o1, int *__restrict__ c1) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const int j = newpos[i];
    x1[j] = x0[i]; n1[j] = n0[i]; v1[j] = v0[i]; o1[j] = o0[i]; c1[j] = c0[i];
}

} // namespace orbc
