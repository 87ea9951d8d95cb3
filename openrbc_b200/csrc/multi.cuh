// multi.cuh — spatial decomposition of ONE cell over the GPUs of a box.
//
// The reference is a single shared-memory process; its own thread decomposition is the template: contiguous ranges of the
// Morton-ordered Voronoi cells per worker (util_numa.h:41-42), cell pairs that straddle two ranges evaluated by BOTH owners
// one-sidedly (compute_pairwise_fused.h:264-275,287-295,303-314), so only positions and directors ever cross a range
// boundary, never forces.  Here a worker is a GPU:
//
//   * every rank keeps the containers in the same GLOBAL index space (particles sorted by cell) and owns the slots of its
//     cells; per step it computes forces for and integrates only those;
//   * halo exchange = the integrator also stores the new x, n of an owned particle straight into the arrays of the ranks
//     that own a cell of its cell's r<9 centroid stencil (or a bonded partner), at the same global slot, through
//     peer-mapped pointers over NVLink — no packing, no staging, no collective call (integrate.cuh PushArgs);
//   * rebuild (every freq_voronoi steps) = owners publish their centroids to every rank, every rank builds the same grid /
//     Morton order, particles are re-assigned by their owners, per-rank arrival counts are exchanged so that all ranks derive
//     the same global cell_start, and every particle is moved by ONE store into its new global slot on the new owner
//     (migration and reorder are the same kernel, rebuild.cuh k_rank_and_move);
//   * ranks synchronise through epoch flags in peer memory (k_mg_barrier): st.release.sys to every peer, ld.acquire.sys
//     spin on the local slots; stream order does the rest.  No host synchronisation inside a run.
//
// Results do not depend on the number of ranks: cells, slots, stencils and the per-particle summation order of the gathered
// forces are those of the single-GPU run (ranks own ascending slot ranges, so "members from lower ranks first" is still
// ascending old index, the reference's arrival order at one thread).
#pragma once
#include <unistd.h>

#include "common.cuh"
#include "primitives.cuh"
#include "rebuild.cuh"
#include "pair.cuh"
#include "integrate.cuh"

namespace orbc {

// ---- barrier ------------------------------------------------------------------------------------------------------------------
struct BarrierArgs {
    unsigned *peer_flags[kMaxWorld];     // flags array of every rank (peer-mapped)
    const unsigned *mine;                // this rank's flags array
    int rank, world;
    unsigned epoch;
    int *err;                            // d_flags: [3] = 1 + rank that never arrived
};
__global__ void k_mg_barrier(BarrierArgs b) {
    const int r = threadIdx.x;
    if (r >= b.world || r == b.rank) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(b.peer_flags[r] + b.rank), "r"(b.epoch) : "memory");
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(b.mine + r) : "memory");
        if ((int)(v - b.epoch) >= 0) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 20000000000ULL) { atomicExch(b.err + 3, r + 1); break; }     // 20 s: a peer died; report instead of hanging
        __nanosleep(100);
    }
}

// owned slot range of one or both containers from the global cell_start (device side: the host never waits for it)
struct RangeArgs { const int *cell_start; int *range2; int cap; };
__global__ void k_set_range(RangeArgs a0, RangeArgs a1, int count, int cb, int ce, int *__restrict__ flags) {
    if ((int)threadIdx.x >= count) return;
    const RangeArgs &a = threadIdx.x == 0 ? a0 : a1;
    const int b = a.cell_start[cb], e = a.cell_start[ce];
    a.range2[0] = b; a.range2[1] = e;
    if (e - b > a.cap) atomicExch(flags + 3, -(e - b));
}
__global__ void k_set_range_const(int *__restrict__ range2, int b, int e) { range2[0] = b; range2[1] = e; }

// arrival counts of this rank -> row `rank` of every peer's table.  A rank's particles only ever land in its own cells and
// their neighbours, so the row is zero almost everywhere: only entries that are non-zero now or were non-zero at the last
// exchange (prev, so the peers' copies get cleared) cross the link.
// (a0 for the blocks [0, blocks0), a1 for the others: both containers in one launch; the host passes the same set twice for one)
struct ShareArgs { const int *cnt_me; int *prev; int *dst[kMaxWorld]; };
__global__ void k_share_counts(ShareArgs a0, ShareArgs a1, unsigned blocks0, int nc1, int rank, int world) {
    const bool first = blockIdx.x < blocks0;
    const ShareArgs &a = first ? a0 : a1;
    const int c = (int)((first ? blockIdx.x : blockIdx.x - blocks0) * blockDim.x + threadIdx.x);
    if (c >= nc1) return;
    const int v = a.cnt_me[c];
    if (v == 0 && a.prev[c] == 0) return;
    a.prev[c] = v;
    for (int r = 0; r < world; ++r) if (r != rank) a.dst[r][(size_t)rank * nc1 + c] = v;
}
// members of every cell over all ranks (-> global cell_start after the scan) and members that come from lower ranks
struct TotalsArgs { const int *cnt_all; int *cell_start; int *off_me; };
__global__ void k_cell_totals(TotalsArgs a0, TotalsArgs a1, unsigned blocks0, int nc, int rank, int world) {
    const bool first = blockIdx.x < blocks0;
    const TotalsArgs &a = first ? a0 : a1;
    const int c = (int)((first ? blockIdx.x : blockIdx.x - blocks0) * blockDim.x + threadIdx.x);
    if (c >= nc) return;
    int tot = 0, off = 0;
    for (int g = 0; g < world; ++g) {
        const int v = a.cnt_all[(size_t)g * (nc + 1) + c];
        if (g < rank) off += v;
        tot += v;
    }
    a.cell_start[c] = tot; a.off_me[c] = off;
}

// ranks that own a bonded partner of an owned protein (its x must reach them even if the cells are not stencil neighbours)
// Also collects the bonds with at least one owned atom (my_bonds[0] = count, then bond indices): k_bonded walks that list
// instead of the whole bond array.
__global__ void k_bond_mask(const int *__restrict__ bonds, size_t n_bonds, const int *__restrict__ tag2idx, const int *__restrict__ range,
                            const int *__restrict__ cs_p, CellOwners own, unsigned *__restrict__ pmask32, int *__restrict__ my_bonds, int my_cap, int *__restrict__ flags) {
    const size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_bonds) return;
    const int p1 = tag2idx[bonds[3 * l + 1]], p2 = tag2idx[bonds[3 * l + 2]];
    const int lo = range[2], hi = range[3];
    const bool own1 = p1 >= lo && p1 < hi, own2 = p2 >= lo && p2 < hi;
    if (own1 || own2) { const int k = atomicAdd(my_bonds, 1); if (k < my_cap) my_bonds[1 + k] = (int)l; else atomicExch(flags + 3, -(k + 1)); }
    if (own1 == own2) return;
    const int mine = own1 ? p1 : p2, other = own1 ? p2 : p1;
    int g = 0;
    for (int r = 1; r < own.world; ++r) g += (other >= cs_p[own.beg[r]]);
    atomicOr(pmask32 + (mine >> 2), (1u << g) << (8 * (mine & 3)));
}

// x, n of the owned particles of boundary cells -> the ranks that read them (after a migration; the per-step push is fused
// into the integrator)
struct HaloDst { float4 *x[kMaxWorld], *nn[kMaxWorld]; };
struct HaloArgs { const int *range2; const unsigned char *cell_mask, *pmask; const int *cellid; const float4 *x, *nn; HaloDst d; };
__global__ void k_halo_push(HaloArgs a0, HaloArgs a1, unsigned blocks0) {
    const bool first = blockIdx.x < blocks0;
    const HaloArgs &a = first ? a0 : a1;
    const int i = a.range2[0] + (int)((first ? blockIdx.x : blockIdx.x - blocks0) * blockDim.x + threadIdx.x);
    if (i >= a.range2[1]) return;
    unsigned m = a.cell_mask[a.cellid[i]];
    if (a.pmask) m |= a.pmask[i];
    if (!m) return;
    const float4 xv = a.x[i], nv = a.nn[i];
    while (m) {
        const int r = __ffs(m) - 1; m &= m - 1;
        a.d.x[r][i] = xv; a.d.nn[r][i] = nv;
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------------------
inline CellOwners cell_owners(const orbc_ctx *c) {
    CellOwners o;
    const int w = c->mg.on ? c->mg.world : 1;
    o.world = w;
    for (int g = 0; g <= kMaxWorld; ++g) o.beg[g] = g >= w ? c->n_cells : (int)((long long)g * c->n_cells / w);
    return o;
}
inline bool mg_active(const orbc_ctx *c) { return c->mg.on && c->mg.world > 1; }
inline size_t owned_bound(const orbc_ctx *c, int sp) { return mg_active(c) ? c->mg.own_cap[sp] : c->sp[sp].n; }

// every rank's partial kinetic energy (d_acc[0]) -> slot `rank` of every rank's ke_all; the caller puts a barrier behind it
inline int mg_share_ke(orbc_ctx *c);
inline int mg_barrier(orbc_ctx *c) {
    if (!mg_active(c)) return ORBC_OK;
    if (!c->mg.connected) return fail(ORBC_ERR_ARG, "decomposed run: orbc_mg_connect has not been called");
    BarrierArgs b;
    for (int r = 0; r < kMaxWorld; ++r) b.peer_flags[r] = c->mg.peers.flags[r];
    b.mine = c->mg.flags; b.rank = c->mg.rank; b.world = c->mg.world; b.epoch = ++c->mg.epoch; b.err = c->d_flags;
    ORBC_LAUNCH(c, k_mg_barrier, 1, 32, 0, b);
    return ORBC_OK;
}

// two sets of slots used alternately: a fast rank may publish its next partial sum while a slow one still reads this set
inline const double *mg_ke_slots(const orbc_ctx *c) { return c->mg.ke_all + c->mg.ke_par * kMaxWorld; }
inline int mg_share_ke(orbc_ctx *c) {
    if (!mg_active(c)) return ORBC_OK;
    c->mg.ke_par ^= 1;
    KeDst d; for (int r = 0; r < kMaxWorld; ++r) d.dst[r] = c->mg.peers.ke_all[r] + c->mg.ke_par * kMaxWorld;
    ORBC_LAUNCH(c, k_share_ke, 1, 32, 0, c->d_acc, c->mg.rank, c->mg.world, d);
    return ORBC_OK;
}

// the pointers peers write through, in a fixed order (identical on every rank)
constexpr int kMgShared = 29;
inline void mg_shared_list(orbc_ctx *c, void *out[kMgShared]) {
    int k = 0;
    for (int s = 0; s < 2; ++s) for (int b = 0; b < 2; ++b) {
        Species &S = c->sp[s];
        out[k++] = S.x[b]; out[k++] = S.nn[b]; out[k++] = S.v[b]; out[k++] = S.o[b]; out[k++] = S.cellid[b];
    }
    out[k++] = c->mg.cen_buf[0]; out[k++] = c->mg.cen_buf[1];
    out[k++] = c->mg.cnt_all[0]; out[k++] = c->mg.cnt_all[1];
    out[k++] = c->tag2idx; out[k++] = c->mg.flags; out[k++] = c->mg.ke_all;
    out[k++] = c->mg.vol_all; out[k++] = c->mg.cv_ptype;
}
inline void mg_fill_peers(orbc_ctx *c, int r, void *const p[kMgShared]) {
    PeerTable &t = c->mg.peers;
    int k = 0;
    for (int s = 0; s < 2; ++s) for (int b = 0; b < 2; ++b) {
        t.x[s][b][r] = (float4 *)p[k++]; t.nn[s][b][r] = (float4 *)p[k++]; t.v[s][b][r] = (float4 *)p[k++]; t.o[s][b][r] = (float4 *)p[k++];
        t.cellid[s][b][r] = (int *)p[k++];
    }
    t.centroid[0][r] = (float4 *)p[k++]; t.centroid[1][r] = (float4 *)p[k++];
    t.cnt_all[0][r] = (int *)p[k++]; t.cnt_all[1][r] = (int *)p[k++];
    t.tag2idx[r] = (int *)p[k++]; t.flags[r] = (unsigned *)p[k++]; t.ke_all[r] = (double *)p[k++];
    t.vol_all[r] = (double *)p[k++]; t.cv_ptype[r] = (int *)p[k++];
}

struct MgEntry { unsigned long long raw; cudaIpcMemHandle_t handle; };
struct MgBlob {
    unsigned magic; int pid, device, rank, world, n_cells;
    unsigned long long n_l, n_p;
    MgEntry e[kMgShared];
};
constexpr unsigned kMgMagic = 0x4f524243u;   // "ORBC"

} // namespace orbc
