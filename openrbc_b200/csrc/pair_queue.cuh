// pair_queue.cuh — the production pair-force kernels.
//
// Same candidate sets (cells whose CENTROIDS are closer than 6 / 8 / 9, compute_pairwise_fused.h:260,278,299) and the same
// per-pair arithmetic as pair.cuh (the simple kernels kept as an independent cross-check), organised for the SIMT machine:
//
//   k_cell_bounds  per Voronoi cell, a bounding sphere of its current lipids (centre = centroid of the last rebuild) and of
//                  its current proteins (centre = their mean).  A particle can only interact with members of a cell whose
//                  sphere it approaches to within the cutoff, so whole (particle, cell) pairs are skipped by one distance
//                  test; the skip is conservative (margin kCullEps), results are unchanged.
//   k_pair_ll_r    one thread per lipid over precomputed candidate RUNS (k_lipid_runs).  Phase 1 walks the runs and only TESTS the
//                  cutoff (r2 < 6.76 && r2 > 1e-5, compute_pairwise_fused.h:109,134); indices that pass go to a per-lane queue in
//                  shared memory (CPU precedent: the reference's implicit-SIMD "enqueue pairs that pass the cutoff" path,
//                  pairwise_kernel_implicit_simd.h:25-100).  Phase 2 evaluates the queued pairs on dense lanes.  One-sided, no
//                  atomics, fixed summation order.  Since round 2 it is the FALLBACK of the tiled kernel k_pair_ll_t
//                  (pair_tile.cuh): it runs when a cell's candidates do not fit the tile, or on request (option "ll_variant" 1).
//   k_pair_prot    one thread per protein: protein-protein over r<9 (one-sided) and protein-lipid over r<8.  Hits are rare
//                  (~1 per protein per step on the RBC), so each protein-lipid pair is evaluated ONCE, here, and the lipid
//                  receives its share through atomicAdd (fp32 RED) instead of re-testing every pair from the lipid side.
#pragma once
#include "common.cuh"
#include "pair.cuh"

namespace orbc {

constexpr int kQCap = 32;          // queue slots per lane
constexpr int kLLBlock = 64;
constexpr float kCullEps = 4e-3f;  // slack of the bounding-sphere test (absolute, length units)

struct CullTable {                 // per protein type: largest interaction range against lipids / against the protein types present
    float cut_l[kNType], cut_p[kNType];
    float cutsq_p[kNType];         // the largest SQUARED protein-protein cutoff itself (the pair test must not square a rounded root)
};

// ---- bounding spheres ------------------------------------------------------------------------------------------------------
__global__ void k_cell_bounds(const float4 *__restrict__ centroid, int n_cells, const int *__restrict__ cs_l, const float4 *__restrict__ xl,
                              const int *__restrict__ cs_p, const float4 *__restrict__ xp, float4 *__restrict__ lbound, float4 *__restrict__ pbound,
                              const int *__restrict__ need, int need_epoch, const int *__restrict__ gate) {
    if (gate && *gate == 0) return;                          // a list-walking evaluation: nobody culls
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    if (need && need[c] != need_epoch) return;             // decomposed run: neither owned nor halo, its particles are stale here
    {
        const float4 q = centroid[c];
        const int b = cs_l[c], e = cs_l[c + 1];
        float r2 = -1.f;                                   // empty cell (or NaN centroid): never passes the test
        float4 ctr = q;
        if (!(q.x == q.x)) { ctr = e > b ? xl[b] : make_float4(0, 0, 0, 0); }
        for (int j = b; j < e; ++j) {
            const float4 p = xl[j];
            const float dx = p.x - ctr.x, dy = p.y - ctr.y, dz = p.z - ctr.z;
            r2 = fmaxf(r2, dx * dx + dy * dy + dz * dz);
        }
        lbound[c] = make_float4(ctr.x, ctr.y, ctr.z, r2 < 0.f ? -1.f : sqrtf(r2) * 1.0001f);
    }
    if (cs_p) {
        const int b = cs_p[c], e = cs_p[c + 1];
        float mx = 0, my = 0, mz = 0;
        for (int j = b; j < e; ++j) { const float4 p = xp[j]; mx += p.x; my += p.y; mz += p.z; }
        const float s = e > b ? 1.0f / (float)(e - b) : 0.f;
        mx *= s; my *= s; mz *= s;
        float r2 = -1.f;
        for (int j = b; j < e; ++j) {
            const float4 p = xp[j];
            const float dx = p.x - mx, dy = p.y - my, dz = p.z - mz;
            r2 = fmaxf(r2, dx * dx + dy * dy + dz * dz);
        }
        pbound[c] = make_float4(mx, my, mz, r2 < 0.f ? -1.f : sqrtf(r2) * 1.0001f);
    }
}

// true when no member of the sphere `b` can be within `cut` of the point (x, y, z)
__device__ __forceinline__ bool culled(float4 b, float x, float y, float z, float cut) {
    const float dx = x - b.x, dy = y - b.y, dz = z - b.z;
    const float lim = b.w + cut + kCullEps;
    return !(b.w >= 0.f) || (dx * dx + dy * dy + dz * dz > lim * lim);
}

// ---- lipid-lipid --------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float rsqrt_fast(float x) {      // x is a squared distance in (1e-5, 14.6): never denormal
    float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}

__device__ __forceinline__ void sts_i32(unsigned addr, int v) { asm volatile("st.shared.b32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ int lds_i32(unsigned addr) { int v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }

// pairwise_kernel.h:30-68 for particle 1 only (the gathering side), in coefficient form.  With u = d / r, p_k = n_k - (n_k.u) u:
//   f_i  = F_r u + (alpha ua / r) [ (n_j.u) p_i + (n_i.u) p_j ]  =  A1 d + C n_j + B n_i
//   t_i -= alpha ua p_j                                          =>  t_i += -aua n_j + B d
// with aua = alpha att rc^4, B = aua (n_j.u) / r, C = aua (n_i.u) / r, A1 = (F_r - 2 aua (n_i.u)(n_j.u) / r) / r.
// n_i is the lane's own director, so its coefficient is summed as ONE scalar (sB) and applied after the loop.
struct LLConst { float cut, rep8, att4, alpha, alpha_att, one_m_alpha, cutsq; };
// RECHECK: the entry comes from a hit list built with a skin (or is padding): the reference's own guards decide here
// (r2 < cutsq && r2 > 1e-5, compute_pairwise_fused.h:109,134), on the current positions.
template <bool RECHECK>
__device__ __forceinline__ void ll_pair(const LLConst &k, F3 xi, F3 mi, float4 xj, float4 nj,
                                        float &fx, float &fy, float &fz, float &tx, float &ty, float &tz, float &sB) {
    // (explicit fmaf / __fmul_rn everywhere: the searching kernel, the recording kernel and the list walker must round alike, so
    //  that walking a list gives the very bits a search would)
    const float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
    const float r2 = fmaf(dz, dz, fmaf(dy, dy, __fmul_rn(dx, dx)));
    if (RECHECK && !(__float_as_uint(r2) - (__float_as_uint(1e-5f) + 1u) < __float_as_uint(k.cutsq) - (__float_as_uint(1e-5f) + 1u))) return;
    const float rinv = rsqrt_fast(r2);
    const float r = __fmul_rn(r2, rinv);
    const float ninj = fmaf(mi.z, nj.z, fmaf(mi.y, nj.y, __fmul_rn(mi.x, nj.x)));
    const float niu = __fmul_rn(fmaf(mi.z, dz, fmaf(mi.y, dy, __fmul_rn(mi.x, dx))), rinv);
    const float nju = __fmul_rn(fmaf(nj.z, dz, fmaf(nj.y, dy, __fmul_rn(nj.x, dx))), rinv);
    const float A = fmaf(k.alpha, fmaf(-niu, nju, ninj), k.one_m_alpha);   // 1 + alpha (a - 1)
    const float rc = __fsub_rn(k.cut, r);
    const float rc2 = __fmul_rn(rc, rc), rc3 = __fmul_rn(rc2, rc), rc4 = __fmul_rn(rc2, rc2);
    const float fra = fmaf(k.rep8, __fmul_rn(rc3, rc4), __fmul_rn(k.att4, __fmul_rn(A, rc3)));   // 8 rep rc^7 + 4 A att rc^3
    const float aua = __fmul_rn(k.alpha_att, rc4);                          // alpha * att * rc^4
    const float auar = __fmul_rn(aua, rinv);
    const float B = __fmul_rn(auar, nju), C = __fmul_rn(auar, niu);
    const float A1 = __fmul_rn(fmaf(__fmul_rn(-2.0f, C), nju, fra), rinv);
    fx = fmaf(A1, dx, fmaf(C, nj.x, fx)); fy = fmaf(A1, dy, fmaf(C, nj.y, fy)); fz = fmaf(A1, dz, fmaf(C, nj.z, fz));
    tx = fmaf(B, dx, fmaf(-aua, nj.x, tx)); ty = fmaf(B, dy, fmaf(-aua, nj.y, ty)); tz = fmaf(B, dz, fmaf(-aua, nj.z, tz));
    sB = __fadd_rn(sB, B);
}
template <bool RECHECK>
__device__ __forceinline__ void ll_eval(const LLConst &k, const float4 *__restrict__ xl, const float4 *__restrict__ nl, F3 xi, F3 mi, int j,
                                        float &fx, float &fy, float &fz, float &tx, float &ty, float &tz, float &sB) {
    ll_pair<RECHECK>(k, xi, mi, __ldg(xl + j), __ldg(nl + j), fx, fy, fz, tx, ty, tz, sB);
}

// Common tail of the lipid kernels: the n_i component of the force, (decomposed runs) the lipid side of the protein-lipid pairs
// whose protein lives on another rank, and the store.
__device__ __forceinline__ void ll_finish(const PairArgs &a, int i, bool live, const int *st, F3 xi, F3 mi,
                                          float fx, float fy, float fz, float tx, float ty, float tz, float sB) {
    fx = fmaf(sB, mi.x, fx); fy = fmaf(sB, mi.y, fy); fz = fmaf(sB, mi.z, fz);
    if (a.world > 1 && live && a.n_p) {
        // decomposed run: the lipid side of the protein-lipid pairs whose protein lives on another rank (that rank evaluates the
        // protein side) — the reference's one-sided evaluation across thread ranges, compute_pairwise_fused.h:287-295
        const int c = a.cell_l[i];
        if (a.dest_mask[c]) {
            const int n8 = (a.stencil_cnt[c] >> 8) & 255;
            for (int k = 0; k < n8; ++k) {
                const int c2 = __ldg(st + k);
                if (c2 >= a.cb && c2 < a.ce) continue;
                const int jb = __ldg(a.cs_p + c2), je = __ldg(a.cs_p + c2 + 1);
                for (int j = jb; j < je; ++j) {
                    const float4 xj = __ldg(a.xp + j);
                    const int type = __float_as_int(xj.w);
                    const F3 d = {xj.x - xi.x, xj.y - xi.y, xj.z - xi.z};      // x_protein - x_lipid (compute_pairwise_fused.h:167)
                    const float r2 = dot3(d, d);
                    if (r2 < c_ff.cutsqlp[type] && r2 > 1e-5f) {
                        const float4 nj = __ldg(a.np + j);
                        F3 f, q1, q2;
                        poly48(c_ff.cutlp[type], c_ff.attlp[type], c_ff.replp[type], c_ff.alphalp[type], d, r2, {nj.x, nj.y, nj.z}, mi, f, q1, q2);
                        fx -= f.x; fy -= f.y; fz -= f.z; tx -= q2.x; ty -= q2.y; tz -= q2.z;
                    } else if (r2 < c_ff.lj_cutsq[type] && r2 > 1e-5f) {
                        const F3 f = lj126(c_ff.lj_lj1[type], c_ff.lj_lj2[type], d, r2);
                        fx -= f.x; fy -= f.y; fz -= f.z;
                    }
                }
            }
        }
    }
    if (live) {
        if (a.accumulate) {
            float4 f = a.fl[i], t = a.tl[i];
            f.x += fx; f.y += fy; f.z += fz; t.x += tx; t.y += ty; t.z += tz;
            a.fl[i] = f; a.tl[i] = t;
        } else {
            a.fl[i] = make_float4(fx, fy, fz, 0.f); a.tl[i] = make_float4(tx, ty, tz, 0.f);
        }
    }
}

// ---- candidate RUNS ---------------------------------------------------------------------------------------------------------------------
// The candidates of a lipid are the members of the r<6 stencil cells of its cell, visited in ascending cell id.  Cells are
// numbered in Morton order and particles are stored sorted by cell, so neighbouring stencil cells usually hold neighbouring
// slot ranges: k_lipid_runs merges them, once per rebuild, into a few (first slot, length) runs per cell (full RBC: 4.4 on
// average, 10 at most).  A run is what one bulk copy moves into the tile of k_pair_ll_t, and what one 8-byte load advances the
// candidate stream of k_pair_ll_r by.  Per cell it also packs  runs | candidates << 6 | (tile slot of the cell's own first lipid) << 19
// and raises `tile_overflow` when a cell has more candidates than a tile holds (kTileCap, pair_tile.cuh).
constexpr int kRunStride = 32;     // a cell has at most 32 stencil cells of the r<6 class (k_stencil_build raises a flag otherwise)
__global__ void k_lipid_runs(int cb, int ce, const int *__restrict__ stencil, const int *__restrict__ stencil_cnt, const int *__restrict__ cs_l,
                             int2 *__restrict__ lruns, int *__restrict__ lrun_info, int tile_cap, int *__restrict__ tile_overflow) {
    const int c = cb + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ce) return;
    const int n6 = min(stencil_cnt[c] & 255, 32);
    const int *st = stencil + (size_t)c * kStencilStride;
    int2 *out = lruns + (size_t)c * kRunStride;
    int nr = 0, rb = 0, re = -1, done = 0, own = 0;              // done = candidates in the runs already written
    for (int k = 0; k < n6; ++k) {
        const int c2 = st[k];
        const int b = cs_l[c2], e = cs_l[c2 + 1];
        if (e <= b) continue;
        if (b != re) {
            if (re > rb) { out[nr++] = make_int2(rb, re - rb); done += re - rb; }
            rb = b;
        }
        re = e;
        if (c2 == c) own = done + (b - rb);
    }
    if (re > rb) { out[nr++] = make_int2(rb, re - rb); done += re - rb; }
    if (done > tile_cap || done > 8191) { atomicExch(tile_overflow, 1); done = min(done, 8191); own = 0; }
    lrun_info[c] = nr | done << 6 | own << 19;
}

// ---- hit lists (Verlet lists with a skin) ---------------------------------------------------------------------------------------------------
// Between two rebuilds the partition, the stencils and the storage order do not change; only the particles move, by ~0.005 per
// step.  A force evaluation right after a rebuild therefore records, per lipid, every candidate closer than cut + skin (BUILD), and
// the evaluations up to the next rebuild walk those lists instead of the ~118 candidates per lipid (k_pair_ll_list), re-testing every
// entry with the reference's exact guards on the current positions.  The lists are a superset of the reference's hits as long as
// no particle has moved further than skin / 2 since they were built: the integrators record the largest displacement of every
// step (NlState::disp), k_nl_gate adds them up and orders a fresh build when 2 x (sum of maxima) exceeds the skin.  Same hits,
// same order of evaluation per lipid: the forces are bit-identical to an evaluation without lists.
// Layout: groups of 64 consecutive lipid slots, [group][entry][64] (coalesced for the thread-per-lipid readers), `cap` entries.
constexpr int kNlBackoffMax = 3;    // rebuilds to sit out, at most, after a recording that was never walked
struct NlState {                    // one per context, in device memory
    unsigned disp[64];              // largest squared displacement of a particle in the integration steps since the last gate (float bits)
    float accum;                    // sum of the per-step maxima since the lists were recorded
    int need;                       // what this evaluation does: 0 walks the lists, 1 searches AND records them, 2 searches without recording
    int overflow;                   // a list did not fit its row: the lists are unusable until the next recording
    unsigned builds, reuses;        // statistics: evaluations that recorded / that walked
    int valid;                      // lists were recorded for the current partition
    float d_last;                   // the largest single-step displacement seen last
    unsigned searches;              // statistics: evaluations that searched without recording
    float skin;                     // the skin of the current lists (chosen by the gate at the recording, between the option's value and its ceiling)
    int used, backoff, wait;        // walks of the current lists; rebuilds to sit out after lists that were never walked (doubles, halves)
    unsigned work[8];               // tickets of the gated kernels of this evaluation (next_piece), zeroed by the gate
};
// The decision, once per force evaluation (one warp).  `force`: the partition has changed since the last evaluation (host's knowledge).
//   after a rebuild   record, with a skin that leaves room for a step 25 % longer than the one just seen (at least the option's skin;
//                     a system with a few fast particles gets thicker lists instead of none), if that skin stays under the ceiling
//                     (skin_max: lists grow with the cube of cutoff + skin) -- unless lists were lately recorded and never walked
//                     (recording costs ~16 % on top of a search and pays if one of three recordings is walked): then sit out
//                     1, then 3 rebuilds (kNlBackoffMax; halving with every recording that was walked)
//   otherwise         walk, if lists exist and 2 x (sum of the per-step maxima since the recording) <= their skin; search if not
// `shared` (decomposed run): the per-rank maxima of the last integration step, published by k_nl_share into every rank's table
// (the barrier behind the halo push stands between the two kernels): every rank takes the same maximum and decides alike.
__global__ void k_nl_gate(NlState *st, int force, int fixed_mode, int moves, float skin_min, float skin_max, const unsigned *__restrict__ shared, int world) {
    const int lane = threadIdx.x;
    unsigned m;
    if (shared) m = lane < world ? shared[lane] : 0u;
    else { m = max(st->disp[lane], st->disp[lane + 32]); st->disp[lane] = 0u; st->disp[lane + 32] = 0u; }
    m = __reduce_max_sync(0xffffffffu, m);
    if (lane < 8) st->work[lane] = 0u;
    if (lane == 0) {
        const float d = sqrtf(__uint_as_float(m)) * 1.0001f + 1e-4f;          // + the rounding of x + v dt at |x| ~ 1000
        if (moves > 0) st->d_last = d;
        int mode;
        if (force) {
            const float skin = fmaxf(skin_min, 2.5f * st->d_last);
            st->skin = fminf(skin, fmaxf(skin_min, skin_max));
            if (fixed_mode > 0) mode = fixed_mode;               // (measurement aid)
            else if (st->wait > 0) { st->wait--; mode = 2; }
            else mode = skin <= st->skin ? 1 : 2;
            st->valid = mode == 1; st->accum = 0.f; st->overflow = 0; st->used = 0;
        } else if (!st->valid || st->overflow) {
            if (st->valid && st->used == 0) { st->backoff = min(2 * st->backoff + 1, kNlBackoffMax); st->wait = st->backoff; }   // (rows overflowed: recorded for nothing)
            mode = 2; st->valid = 0;
        }
        else {
            const float acc = st->accum + (float)moves * d;
            if (2.0f * acc <= st->skin) { mode = 0; st->accum = acc; if (st->used++ == 0) st->backoff >>= 1; }
            else {
                mode = 2; st->valid = 0;
                if (st->used == 0) { st->backoff = min(2 * st->backoff + 1, kNlBackoffMax); st->wait = st->backoff; }   // recorded for nothing
            }
        }
        if (mode == 0) st->reuses++; else if (mode == 1) st->builds++; else st->searches++;
        st->need = mode;
    }
}
// decomposed run, after the integrator kernels of a step: this rank's largest squared displacement -> slot `rank` of every rank
struct NlShare { unsigned *dst[kMaxWorld]; };
__global__ void k_nl_share(NlState *st, int rank, int world, NlShare d) {
    const int lane = threadIdx.x;
    unsigned m = max(st->disp[lane], st->disp[lane + 32]);
    st->disp[lane] = 0u; st->disp[lane + 32] = 0u;
    m = __reduce_max_sync(0xffffffffu, m);
    if (lane < world) d.dst[lane][rank] = m;
}
// what the integrators call, with EVERY thread of the block (d2 = 0 for idle ones): the largest squared displacement of the step,
// one RED per block into 64 slots (one per warp cost 11 us per launch on the RBC)
__device__ __forceinline__ void nl_track(unsigned *disp, float d2) {
    if (!disp) return;                                           // (uniform: a kernel argument)
    __shared__ unsigned s_max;
    if (threadIdx.x == 0) s_max = 0u;
    __syncthreads();
    const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(d2));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(&s_max, m);
    __syncthreads();
    if (threadIdx.x == 0 && s_max) atomicMax(disp + (blockIdx.x & 63), s_max);
}

// The pieces of work of a block.  Without a ticket counter (work == nullptr): blockIdx.x, then on by the size of the grid.  The
// gated kernels of the hit lists are launched three at a time (walk / record / search) and two of them return at once: they get a
// grid that fills the GPU once, and their blocks draw pieces from a ticket counter, in order, as the block scheduler would have
// handed them out.  (The lipid kernels only: the protein kernels, 8.8 k blocks with the expensive proteins first, are faster over
// their natural grid -- walker 67 us against 107 us with tickets -- and a launch that returns at once costs them ~4 us.)
__device__ __forceinline__ bool next_piece(unsigned *work, unsigned *s_piece, unsigned &piece, bool first) {
    if (first) { piece = blockIdx.x; return true; }             // (no ticket for the first piece: a burst of same-address atomics costs ~10 us)
    if (!work) { piece += gridDim.x; return true; }
    __syncthreads();                                            // (everybody has read the previous ticket)
    if (threadIdx.x == 0) *s_piece = gridDim.x + atomicAdd(work, 1u);
    __syncthreads();
    piece = *s_piece;
    return true;
}
struct LLList { int *list; int *cnt; int cap; NlState *st; };
__device__ __forceinline__ size_t ll_row(int i, int cap) { return ((size_t)(i >> 6) * cap) * 64 + (i & 63); }

// Phase 2 of the recording kernel: the lane's queue of hits, two at a time -- both partners' x and n are gathered before the first force
// body starts, so the second pair's loads fly under the first pair's arithmetic (the bodies run in queue order: same sums).  The
// warp's lanes write the same row index (the longest queue decides, `len`; shorter queues are padded
// with the lane's own slot, which fails the guards and costs no gather): one coalesced 128-byte store per entry instead of 32 sectors.
__device__ __forceinline__ void ll_drain_record(const LLConst &kc, const float4 *__restrict__ xl, const float4 *__restrict__ nl_, F3 xi, F3 mi, unsigned q0, unsigned qp,
                                                unsigned len, int self, bool live, int *__restrict__ row, int cap, int &total,
                                                float &fx, float &fy, float &fz, float &tx, float &ty, float &tz, float &sB) {
    const unsigned mine = qp - q0;
    const float4 own = make_float4(xi.x, xi.y, xi.z, 0.f);
    for (unsigned off = 0; off < len; off += 256) {
        const bool h0 = off < mine, h1 = off + 128 < mine, second = off + 128 < len;
        const int j0 = h0 ? lds_i32(q0 + off) : self, j1 = h1 ? lds_i32(q0 + off + 128) : self;
        if (live && total < cap) row[(size_t)total * 64] = j0;
        ++total;
        if (second) { if (live && total < cap) row[(size_t)total * 64] = j1; ++total; }
        float4 x0 = own, x1 = own, n0 = own, n1 = own;
        if (h0) { x0 = __ldg(xl + j0); n0 = __ldg(nl_ + j0); }
        if (h1) { x1 = __ldg(xl + j1); n1 = __ldg(nl_ + j1); }
        if (h0) ll_pair<true>(kc, xi, mi, x0, n0, fx, fy, fz, tx, ty, tz, sB);
        if (h1) ll_pair<true>(kc, xi, mi, x1, n1, fx, fy, fz, tx, ty, tz, sB);
    }
}

// W = candidates per lane and iteration.  Measured on the full RBC (B200): W = 4 with 20 resident blocks 470 us; W = 8 505-515 us
// (longer partial groups, 64 registers); 24 resident blocks at 40 registers 538 us (spills); an L1 prefetch 4-16 candidates ahead
// of the stream changes nothing.  Round 2 (profiles/r02_*): partners gathered as interleaved 32-byte (x, n) records with one
// 256-bit load 454 us (no gain: the stalls are load LATENCY in phase 1, 31 % of the samples, not L1 wavefronts); the warp-per-cell
// tile kernel (pair_tile.cuh) 713 us.
// BUILD: also record the hit lists (window 0 <= r2 < (cut + skin)^2, exact guards at evaluation).  `gate`/`want`: run only if
// *gate == want (the list walker and this kernel are launched together on steps without a rebuild; one of them returns at once).
// The grid may be smaller than the number of 64-lipid groups (grid-stride), so that a gated launch that returns costs nothing.
// (Two hits at a time in phase 2 -- ll_drain_record -- pays where the registers are there anyway: the recording kernel, 64 registers at 16
// blocks per SM, 587 -> 528 us.  The plain search stays one at a time: 48 registers at 20 blocks, 457 us; two at a time it spills at 20
// blocks, 500 us, and runs at 465 / 467 us with 16 / 18 blocks; a depth-one software pipeline of the gathers: 460 us.)
template <int MINB, int W, bool BUILD>
__global__ void __launch_bounds__(kLLBlock, MINB) k_pair_ll_r(PairArgs a, const LLConst kc, const int2 *__restrict__ lruns, const int *__restrict__ lrun_info,
                                                               const int *__restrict__ gate, int want, LLList nl, unsigned *work) {
    if (gate && *gate != want) return;
    const float skin = BUILD ? nl.st->skin : 0.f;                // (the gate's choice for this recording)
    __shared__ unsigned s_piece;
    __shared__ int s_q[kLLBlock / 32][kQCap * 32];
    const int lane = threadIdx.x & 31;
    int *const q = s_q[threadIdx.x >> 5] + lane;
    const float4 *__restrict__ xl = a.xl;
    const float4 *__restrict__ nl_ = a.nl;
    // (kc: the lipid-lipid constants as a kernel parameter — read from c_ff they were rematerialised inside the loops)
    // r2 > 1e-5 && r2 < cutsq (compute_pairwise_fused.h:109,134) as ONE unsigned comparison of the bit patterns: r2 is a sum of
    // squares (never negative), and non-negative floats order like their bits; a NaN lies above every finite pattern
    const float lim = BUILD ? (kc.cut + skin) * (kc.cut + skin) : kc.cutsq;
    const unsigned lo_bits = BUILD ? 0u : __float_as_uint(1e-5f) + 1u, span = __float_as_uint(lim) - lo_bits;
    const unsigned q0 = (unsigned)__cvta_generic_to_shared(q);
    const unsigned q_full = q0 + (kQCap - W) * 128;              // a group of W always fits below this mark
    const int l0 = a.range[0], l1 = a.range[1];
    unsigned piece;
    for (bool first = true; next_piece(work, &s_piece, piece, first); first = false) {
        const int base = l0 + (int)piece * kLLBlock;
        if (base >= l1) break;
        const int i = base + threadIdx.x;
        const bool live = i < l1;
        float fx = 0, fy = 0, fz = 0, tx = 0, ty = 0, tz = 0, sB = 0;
        F3 xi = {0, 0, 0}, mi = {0, 0, 0};
        const int *st = a.stencil;
        const int2 *rp = lruns;
        int nr = 0;
        if (live) {
            const float4 xi4 = xl[i], ni4 = nl_[i];
            xi = {xi4.x, xi4.y, xi4.z}; mi = {ni4.x, ni4.y, ni4.z};
            const int c = a.cell_l[i];
            st += (size_t)c * kStencilStride;
            rp += (size_t)c * kRunStride;
            nr = __ldg(lrun_info + c) & 63;
        }
        const int self = live ? i : l0;
        int *row = BUILD ? nl.list + ll_row(self, nl.cap) : nullptr;
        int total = 0;                                           // entries recorded so far (BUILD): the same number in every lane of the warp
        unsigned qp = q0;
        int2 nx = make_int2(0, 0);                               // the next run, loaded one advance ahead
        if (nr > 0) nx = __ldg(rp);
        int k = 0, cur = 0, rem = 0;
        for (;;) {
            if (rem <= 0 && k < nr) { cur = nx.x; rem = nx.y; ++k; if (k < nr) nx = __ldg(rp + k); }
            if (!__any_sync(0xffffffffu, rem > 0)) break;
            if (__any_sync(0xffffffffu, qp > q_full)) {          // make room: every lane drains its queue (dense)
                if (BUILD) {
                    // recording: the warp's lanes write the same row index (the longest queue decides, shorter ones are padded with
                    // the lane's own slot, which fails the guards): one coalesced 128-byte store per entry instead of 32 sectors
                    const unsigned len = __reduce_max_sync(0xffffffffu, qp - q0);
                    ll_drain_record(kc, xl, nl_, xi, mi, q0, qp, len, self, live, row, nl.cap, total, fx, fy, fz, tx, ty, tz, sB);
                } else
                    for (unsigned e = q0; e < qp; e += 128) ll_eval<false>(kc, xl, nl_, xi, mi, lds_i32(e), fx, fy, fz, tx, ty, tz, sB);
                qp = q0;
            }
            const float4 *__restrict__ p = xl + cur;
            float4 xj[W];
            #pragma unroll
            for (int u = 0; u < W; ++u) xj[u] = __ldg(p + u);
            #pragma unroll
            for (int u = 0; u < W; ++u) {
                const float dx = xi.x - xj[u].x, dy = xi.y - xj[u].y, dz = xi.z - xj[u].z;
                const float r2 = dx * dx + dy * dy + dz * dz;
                if (u < rem && __float_as_uint(r2) - lo_bits < span) { sts_i32(qp, cur + u); qp += 128; }
            }
            if (rem > 0) cur += W;
            rem -= W;
        }
        if (BUILD) {
            const unsigned len = __reduce_max_sync(0xffffffffu, qp - q0);
            ll_drain_record(kc, xl, nl_, xi, mi, q0, qp, len, self, live, row, nl.cap, total, fx, fy, fz, tx, ty, tz, sB);
        } else
            for (unsigned e = q0; e < qp; e += 128) ll_eval<false>(kc, xl, nl_, xi, mi, lds_i32(e), fx, fy, fz, tx, ty, tz, sB);
        if (BUILD && live) {
            nl.cnt[i] = min(total, nl.cap);
            if (total > nl.cap) atomicExch(&nl.st->overflow, 1);
        }
        ll_finish(a, i, live, st, xi, mi, fx, fy, fz, tx, ty, tz, sB);
        __syncwarp();
    }
}

// The list walker: one thread per lipid, entries read coalesced, two partners in flight per lane.  An entry beyond the lane's
// count is replaced by the lane's own slot (r2 = 0 fails the guards), which keeps the loop free of branches around the loads.
// (Interleaved 32-byte (x, n) records, one 256-bit load per partner instead of two 128-bit loads, were measured and removed: 243 against
// 238 us per launch -- the walker is limited by the bytes through the L1 data path, not by the number of requests.)
template <int MINB>
__global__ void __launch_bounds__(kLLBlock, MINB) k_pair_ll_list(PairArgs a, const LLConst kc, const int *__restrict__ gate, int want, LLList nl, unsigned *work) {
    if (gate && *gate != want) return;
    __shared__ unsigned s_piece;
    const float4 *__restrict__ xl = a.xl;
    const float4 *__restrict__ nl_ = a.nl;
    const int l0 = a.range[0], l1 = a.range[1];
    unsigned piece;
    for (bool first = true; next_piece(work, &s_piece, piece, first); first = false) {
        const int base = l0 + (int)piece * kLLBlock;
        if (base >= l1) break;
        const int i = base + threadIdx.x;
        const bool live = i < l1;
        float fx = 0, fy = 0, fz = 0, tx = 0, ty = 0, tz = 0, sB = 0;
        F3 xi = {0, 0, 0}, mi = {0, 0, 0};
        int cnt = 0;
        const int self = live ? i : l0;
        if (live) {
            const float4 xi4 = xl[i], ni4 = nl_[i];
            xi = {xi4.x, xi4.y, xi4.z}; mi = {ni4.x, ni4.y, ni4.z};
            cnt = __ldg(nl.cnt + i);
        }
        const int *row = nl.list + ll_row(self, nl.cap);
        const int maxc = __reduce_max_sync(0xffffffffu, cnt);
        for (int s = 0; s < maxc; s += 4) {
            int j[4]; float4 xj[4], nj[4];
            #pragma unroll
            for (int u = 0; u < 4; ++u) { j[u] = s + u < cnt ? __ldg(row + (size_t)(s + u) * 64) : -1; if (j[u] == self) j[u] = -1; }   // (recorded padding = the lane's own slot)
            // four partners in flight per lane.  A gather costs one L1 request per LANE (the walker runs at that limit: 2 requests
            // per pair, ~1 per cycle and SM), so padding entries must not gather: they get the lane's own position, which fails the guards
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                xj[u] = make_float4(xi.x, xi.y, xi.z, 0.f); nj[u] = xj[u];
                if (j[u] >= 0) { xj[u] = __ldg(xl + j[u]); nj[u] = __ldg(nl_ + j[u]); }
            }
            #pragma unroll
            for (int u = 0; u < 4; ++u) ll_pair<true>(kc, xi, mi, xj[u], nj[u], fx, fy, fz, tx, ty, tz, sB);
        }
        ll_finish(a, i, live, a.stencil + (live ? (size_t)a.cell_l[i] * kStencilStride : 0), xi, mi, fx, fy, fz, tx, ty, tz, sB);
    }
}

// ---- proteins -------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_add3(float4 *dst, float x, float y, float z) {
    atomicAdd(dst, make_float4(x, y, z, 0.f));                   // one 16-byte RED (REDG.ADD.F32x4); .w of f and t is unused
}

// Thread -> protein map of k_pair_prot: proteins whose type reaches far into the bilayer (band-3, glycophorin: 2.6) first,
// the LJ-core-only types (actin, spectrin: 1.1225) after them, each class in storage order.  The work of a protein thread
// scales with the square of its interaction range, and a warp is as slow as its slowest lane, so mixed warps would run at
// the pace of the two or three heavy proteins in them.  Built after every protein reorder: flag -> scan -> scatter.
// Both kernels run over `cap` threads (the launch bound of the owned proteins); thread k stands for protein slot p0 + k.
__global__ void k_porder_flag(const float4 *__restrict__ xp, const int *__restrict__ range, size_t cap, CullTable ct, float heavy_cut, int *__restrict__ flag) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cap) return;
    const int i = range[2] + (int)k;
    int h = 0;
    if (i < range[3]) { const int t = __float_as_int(xp[i].w); h = (t >= 0 && t < kNType && ct.cut_l[t] >= heavy_cut) ? 1 : 0; }
    flag[k] = h;
}
__global__ void k_porder_scatter(const int *__restrict__ scan /* exclusive, scan[cap] = n_heavy */, const int *__restrict__ range, size_t cap, const float4 *__restrict__ xp,
                                 CullTable ct, float heavy_cut, int *__restrict__ porder) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = range[2] + (int)k;
    if (k >= cap || i >= range[3]) return;
    const int t = __float_as_int(xp[i].w);
    const bool heavy = t >= 0 && t < kNType && ct.cut_l[t] >= heavy_cut;
    const int before = scan[k];                                   // heavy proteins with a lower index
    porder[heavy ? before : scan[cap] + ((int)k - before)] = i;
}

constexpr int kPBlock = 64;
constexpr int kRangeCap = 8;       // stencil slots handled per round

// LPP lanes per protein (adjacent lanes; each takes every LPP-th cell of the stencil, the partial sums are combined with
// shuffles at the end).  The kernel is bound by chains of dependent loads (stencil -> bounds / ranges -> members -> directors):
// with one lane per protein a warp needs ~100 serial round trips to memory and a rank of a decomposed run has too few
// warps to hide them; four lanes per protein make the chains four times shorter and give four times as many warps.
// Phase 0 culls the member lists of the stencil cells against the bounding spheres and COMPACTS the
// survivors into per-lane range lists in shared memory (a lane-level `if (culled) skip` would save nothing on a SIMT machine;
// the compaction is what turns skipped cells into skipped warp iterations).  Phase 1 walks the r-th surviving range of every
// lane together, warp-uniform trip counts, predicated bodies.
struct PLists { int *pl, *pl_cnt, *pp, *pp_cnt; int cap_pl, cap_pp; NlState *st; };   // rows indexed by the THREAD id (porder position), [group of 64][entry][64]

// one protein-lipid pair from the protein's side with the reference's branch structure (compute_pairwise_fused.h:167-176): the
// protein accumulates in registers, the lipid (if this rank owns it) gets its share by a 16-byte RED
__device__ __forceinline__ void pl_pair(const PairArgs &a, int type1, float cutsq, float ljcut, F3 xi, F3 mi, int j, float4 xj, int l0, int l1,
                                        float &fx, float &fy, float &fz, float &tx, float &ty, float &tz) {
    const F3 d = {xi.x - xj.x, xi.y - xj.y, xi.z - xj.z};          // x_protein - x_lipid (compute_pairwise_fused.h:167)
    const float r2 = dot3(d, d);
    if (!(r2 > 1e-5f)) return;
    if (r2 < cutsq) {
        const float4 nj = __ldg(a.nl + j);
        F3 f, q1, q2;
        poly48(c_ff.cutlp[type1], c_ff.attlp[type1], c_ff.replp[type1], c_ff.alphalp[type1], d, r2, mi, {nj.x, nj.y, nj.z}, f, q1, q2);
        fx += f.x; fy += f.y; fz += f.z; tx -= q1.x; ty -= q1.y; tz -= q1.z;
        if (j >= l0 && j < l1) { atomic_add3(a.fl + j, -f.x, -f.y, -f.z); atomic_add3(a.tl + j, -q2.x, -q2.y, -q2.z); }
    } else if (r2 < ljcut) {
        const F3 f = lj126(c_ff.lj_lj1[type1], c_ff.lj_lj2[type1], d, r2);
        fx += f.x; fy += f.y; fz += f.z;
        if (j >= l0 && j < l1) atomic_add3(a.fl + j, -f.x, -f.y, -f.z);
    }
}
// one protein-protein pair, gathering side only (compute_pairwise_fused.h:196-208)
__device__ __forceinline__ void pp_pair(int type1, F3 xi, float4 xj, const float *s_cutsqpp, const float *s_ljcutsq, float &fx, float &fy, float &fz) {
    const F3 d = {xi.x - xj.x, xi.y - xj.y, xi.z - xj.z};
    const float r2 = dot3(d, d);
    if (!(r2 > 1e-5f)) return;
    const int type12 = type1 + __float_as_int(xj.w) * kNType;
    if (r2 < s_cutsqpp[type12]) {
        const F3 f = rep8(c_ff.cutpp[type12], c_ff.reppp[type12], d, r2);
        fx += f.x; fy += f.y; fz += f.z;
    } else if (r2 < s_ljcutsq[type12]) {
        const F3 f = lj126(c_ff.lj_lj1[type12], c_ff.lj_lj2[type12], d, r2);
        fx += f.x; fy += f.y; fz += f.z;
    }
}

// BUILD: also record, per THREAD (LPP threads share a protein, each with its own row), every lipid closer than the protein's range
// + skin and every protein closer than the pair's range + skin (the hit lists of k_pair_prot_list, see the lipid kernels above).
template <int LPP, bool BUILD>
__global__ void __launch_bounds__(kPBlock, 16) k_pair_prot(PairArgs a, const float4 *__restrict__ lbound, const float4 *__restrict__ pbound, CullTable ct,
                                                             const int *__restrict__ porder, const int *__restrict__ gate, int want, PLists nl) {
    if (gate && *gate != want) return;
    const float skin = BUILD ? nl.st->skin : 0.f;                // (the gate's choice for this recording)
    __shared__ float s_cutsqpp[36], s_ljcutsq[36], s_recsq[36];
    __shared__ int s_jb[2][kPBlock / 32][kRangeCap * 32];
    __shared__ unsigned short s_len[2][kPBlock / 32][kRangeCap * 32];
    for (int k = threadIdx.x; k < 36; k += blockDim.x) {
        s_cutsqpp[k] = c_ff.cutsqpp[k]; s_ljcutsq[k] = c_ff.lj_cutsq[k];
        const float r = sqrtf(fmaxf(c_ff.cutsqpp[k], c_ff.lj_cutsq[k]));                 // range of this pair of types; recorded up to range + skin
        s_recsq[k] = r > 0.f ? (r + skin) * (r + skin) * 1.00001f : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int *const ljb = s_jb[0][w] + lane, *const pjb = s_jb[1][w] + lane;
    unsigned short *const llen = s_len[0][w] + lane, *const plen = s_len[1][w] + lane;
    const int n_own = a.range[3] - a.range[2];
    const int l0 = a.range[0], l1 = a.range[1];                  // lipids of other ranks get their share from the lipid kernel's epilogue over there
    for (int tid0 = blockIdx.x * kPBlock; tid0 / LPP < n_own; tid0 += gridDim.x * kPBlock) {
    const int tid = tid0 + threadIdx.x;
    const int pid = tid / LPP, sub = tid % LPP;
    const bool live = pid < n_own;
    const int i = live ? porder[pid] : 0;
    F3 xi = {0, 0, 0}, mi = {0, 0, 0};
    int type1 = 0, n8 = 0, n9 = 0;
    const int *st = a.stencil;
    if (live) {
        const float4 xi4 = a.xp[i], ni4 = a.np[i];
        xi = {xi4.x, xi4.y, xi4.z}; mi = {ni4.x, ni4.y, ni4.z};
        type1 = __float_as_int(xi4.w);
        const int c = a.cell_p[i];
        const int cnt = a.stencil_cnt[c];
        n8 = (cnt >> 8) & 255; n9 = cnt >> 16;
        st += (size_t)c * kStencilStride;
    }
    // per-type constants in registers (type1 is fixed for the thread)
    const float cutsq = c_ff.cutsqlp[type1], ljcut = c_ff.lj_cutsq[type1];
    float cull_l = type1 == 0 ? ct.cut_l[0] : type1 == 1 ? ct.cut_l[1] : type1 == 2 ? ct.cut_l[2] : type1 == 3 ? ct.cut_l[3] : type1 == 4 ? ct.cut_l[4] : ct.cut_l[5];
    float cull_p = type1 == 0 ? ct.cut_p[0] : type1 == 1 ? ct.cut_p[1] : type1 == 2 ? ct.cut_p[2] : type1 == 3 ? ct.cut_p[3] : type1 == 4 ? ct.cut_p[4] : ct.cut_p[5];
    const bool any_l = cull_l > 0.f, any_p = cull_p > 0.f;       // a type with no interaction at all against lipids / the proteins present
    // squared test radii: the largest squared cutoff of the type (exactly the table's values), widened by the skin when recording
    float testsq_l = fmaxf(cutsq, ljcut);
    float testsq_p = type1 == 0 ? ct.cutsq_p[0] : type1 == 1 ? ct.cutsq_p[1] : type1 == 2 ? ct.cutsq_p[2] : type1 == 3 ? ct.cutsq_p[3] : type1 == 4 ? ct.cutsq_p[4] : ct.cutsq_p[5];
    if (BUILD) { cull_l += skin; cull_p += skin; testsq_l = cull_l * cull_l * 1.00001f; testsq_p = cull_p * cull_p * 1.00001f; }
    int *row_l = nullptr, *row_p = nullptr; int rec_l = 0, rec_p = 0;
    if (BUILD) { row_l = nl.pl + ll_row(tid, nl.cap_pl); row_p = nl.pp + ll_row(tid, nl.cap_pp); }
    float fx = 0, fy = 0, fz = 0, tx = 0, ty = 0, tz = 0;
    const int nmax = __reduce_max_sync(0xffffffffu, n9);
    for (int kb = 0; kb < nmax; kb += kRangeCap * LPP) {
        // ---- phase 0: cull both member lists of every stencil cell, compact the survivors ------------------------------------------
        int nl_ = 0, np_ = 0;
        const int kend = min(kRangeCap, (nmax - kb + LPP - 1) / LPP);
        #pragma unroll 4
        for (int kk = 0; kk < kend; ++kk) {
            const int k = kb + kk * LPP + sub;
            const bool in9 = k < n9, in8 = k < n8;
            int c2 = 0;
            if (in9) c2 = __ldg(st + k);
            float4 bp = make_float4(0.f, 0.f, 0.f, -1.f), bl = bp;
            int pb = 0, pe = 0, lb = 0, le = 0;
            if (in9) { bp = __ldg(pbound + c2); pb = __ldg(a.cs_p + c2); pe = __ldg(a.cs_p + c2 + 1); }
            if (in8) { bl = __ldg(lbound + c2); lb = __ldg(a.cs_l + c2); le = __ldg(a.cs_l + c2 + 1); }
            if (in9 && pe > pb && any_p && !culled(bp, xi.x, xi.y, xi.z, cull_p)) { pjb[np_ * 32] = pb; plen[np_ * 32] = (unsigned short)min(pe - pb, 65535); ++np_; }
            if (in8 && le > lb && any_l && !culled(bl, xi.x, xi.y, xi.z, cull_l)) { ljb[nl_ * 32] = lb; llen[nl_ * 32] = (unsigned short)min(le - lb, 65535); ++nl_; }
        }
        // ---- protein-lipid: evaluated once, here; the lipid gets its share by atomics (compute_pairwise_fused.h:143-179,278-297) ---
        // every lane streams through ITS surviving ranges four candidates at a time (same scheme as k_pair_ll_r)
        {
            int taken = 0, cur = 0, rem = 0;
            for (;;) {
                if (rem <= 0 && taken < nl_) { cur = ljb[taken * 32]; rem = llen[taken * 32]; ++taken; }
                if (!__any_sync(0xffffffffu, rem > 0 || taken < nl_)) break;
                const float4 *__restrict__ p = a.xl + cur;
                float4 xj4[4];
                #pragma unroll
                for (int u = 0; u < 4; ++u) xj4[u] = __ldg(p + u);
                #pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 xj = xj4[u];
                    const float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                    const float r2 = dx * dx + dy * dy + dz * dz;
                    if (u < rem && r2 < testsq_l) {
                        if (BUILD) { if (rec_l < nl.cap_pl) row_l[(size_t)rec_l * 64] = cur + u; ++rec_l; }
                        pl_pair(a, type1, cutsq, ljcut, xi, mi, cur + u, xj, l0, l1, fx, fy, fz, tx, ty, tz);
                    }
                }
                if (rem > 0) cur += 4;
                rem -= 4;
            }
        }
        // ---- protein-protein, one-sided (compute_pairwise_fused.h:182-236, 262-276) ----------------------------------------------------
        {
            int taken = 0, cur = 0, rem = 0;
            for (;;) {
                if (rem <= 0 && taken < np_) { cur = pjb[taken * 32]; rem = plen[taken * 32]; ++taken; }
                if (!__any_sync(0xffffffffu, rem > 0 || taken < np_)) break;
                const float4 *__restrict__ p = a.xp + cur;
                float4 xj4[2];
                #pragma unroll
                for (int u = 0; u < 2; ++u) xj4[u] = __ldg(p + u);
                #pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float4 xj = xj4[u];
                    const float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                    const float r2 = dx * dx + dy * dy + dz * dz;
                    if (u < rem && r2 < testsq_p) {
                        if (BUILD && r2 < s_recsq[type1 + __float_as_int(xj.w) * kNType]) { if (rec_p < nl.cap_pp) row_p[(size_t)rec_p * 64] = cur + u; ++rec_p; }
                        pp_pair(type1, xi, xj, s_cutsqpp, s_ljcutsq, fx, fy, fz);
                    }
                }
                if (rem > 0) cur += 2;
                rem -= 2;
            }
        }
    }
    if (BUILD && live) {
        nl.pl_cnt[tid] = min(rec_l, nl.cap_pl); nl.pp_cnt[tid] = min(rec_p, nl.cap_pp);
        if (rec_l > nl.cap_pl || rec_p > nl.cap_pp) atomicExch(&nl.st->overflow, 1);
    }
    #pragma unroll
    for (int o = 1; o < LPP; o <<= 1) {
        fx += __shfl_xor_sync(0xffffffffu, fx, o); fy += __shfl_xor_sync(0xffffffffu, fy, o); fz += __shfl_xor_sync(0xffffffffu, fz, o);
        tx += __shfl_xor_sync(0xffffffffu, tx, o); ty += __shfl_xor_sync(0xffffffffu, ty, o); tz += __shfl_xor_sync(0xffffffffu, tz, o);
    }
    if (live && sub == 0) {                                      // this lane owns protein i: plain read-modify-write, or plain write
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f), t = f;
        if (a.accumulate) { f = a.fp[i]; t = a.tp[i]; }
        f.x += fx; f.y += fy; f.z += fz; t.x += tx; t.y += ty; t.z += tz;
        a.fp[i] = f; a.tp[i] = t;
    }
    __syncwarp();
    }
}

// The list walker of the proteins: the same thread -> (protein, lane) map as the kernel that recorded, every recorded partner
// re-tested with the reference's guards on the current positions.  ~1 protein-lipid hit per protein and step on the RBC.
template <int LPP>
__global__ void __launch_bounds__(kPBlock) k_pair_prot_list(PairArgs a, const int *__restrict__ porder, const int *__restrict__ gate, int want, PLists nl) {
    if (gate && *gate != want) return;
    __shared__ float s_cutsqpp[36], s_ljcutsq[36];
    for (int k = threadIdx.x; k < 36; k += blockDim.x) { s_cutsqpp[k] = c_ff.cutsqpp[k]; s_ljcutsq[k] = c_ff.lj_cutsq[k]; }
    __syncthreads();
    const int n_own = a.range[3] - a.range[2];
    const int l0 = a.range[0], l1 = a.range[1];
    for (int tid0 = blockIdx.x * kPBlock; tid0 / LPP < n_own; tid0 += gridDim.x * kPBlock) {
        const int tid = tid0 + threadIdx.x;
        const int pid = tid / LPP, sub = tid % LPP;
        const bool live = pid < n_own;
        const int i = live ? porder[pid] : 0;
        float fx = 0, fy = 0, fz = 0, tx = 0, ty = 0, tz = 0;
        if (live) {
            const float4 xi4 = a.xp[i], ni4 = a.np[i];
            const F3 xi = {xi4.x, xi4.y, xi4.z}, mi = {ni4.x, ni4.y, ni4.z};
            const int type1 = __float_as_int(xi4.w);
            const float cutsq = c_ff.cutsqlp[type1], ljcut = c_ff.lj_cutsq[type1];
            const int nl_ = __ldg(nl.pl_cnt + tid), np_ = __ldg(nl.pp_cnt + tid);
            const int *row_l = nl.pl + ll_row(tid, nl.cap_pl), *row_p = nl.pp + ll_row(tid, nl.cap_pp);
            for (int s = 0; s < nl_; s += 4) {                   // four partners at a time: the chain list -> position -> director is all latency
                int j[4]; float4 xj[4];
                #pragma unroll
                for (int u = 0; u < 4; ++u) j[u] = s + u < nl_ ? __ldg(row_l + (size_t)(s + u) * 64) : -1;
                #pragma unroll
                for (int u = 0; u < 4; ++u) xj[u] = j[u] >= 0 ? __ldg(a.xl + j[u]) : make_float4(xi.x, xi.y, xi.z, 0.f);   // padding: r2 = 0 fails the guard
                #pragma unroll
                for (int u = 0; u < 4; ++u) pl_pair(a, type1, cutsq, ljcut, xi, mi, max(j[u], 0), xj[u], l0, l1, fx, fy, fz, tx, ty, tz);
            }
            for (int s = 0; s < np_; s += 4) {                   // (same: the padding is the protein itself, r2 = 0 fails the guard)
                int j[4]; float4 xj[4];
                #pragma unroll
                for (int u = 0; u < 4; ++u) j[u] = s + u < np_ ? __ldg(row_p + (size_t)(s + u) * 64) : -1;
                #pragma unroll
                for (int u = 0; u < 4; ++u) xj[u] = j[u] >= 0 ? __ldg(a.xp + j[u]) : make_float4(xi.x, xi.y, xi.z, 0.f);
                #pragma unroll
                for (int u = 0; u < 4; ++u) pp_pair(type1, xi, xj[u], s_cutsqpp, s_ljcutsq, fx, fy, fz);
            }
        }
        #pragma unroll
        for (int o = 1; o < LPP; o <<= 1) {
            fx += __shfl_xor_sync(0xffffffffu, fx, o); fy += __shfl_xor_sync(0xffffffffu, fy, o); fz += __shfl_xor_sync(0xffffffffu, fz, o);
            tx += __shfl_xor_sync(0xffffffffu, tx, o); ty += __shfl_xor_sync(0xffffffffu, ty, o); tz += __shfl_xor_sync(0xffffffffu, tz, o);
        }
        if (live && sub == 0) {
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f), t = f;
            if (a.accumulate) { f = a.fp[i]; t = a.tp[i]; }
            f.x += fx; f.y += fy; f.z += fz; t.x += tx; t.y += ty; t.z += tz;
            a.fp[i] = f; a.tp[i] = t;
        }
    }
}

} // namespace orbc
