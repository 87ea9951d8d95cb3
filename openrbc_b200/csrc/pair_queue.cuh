// pair_queue.cuh — the production pair-force kernels.
//
// Same candidate sets (cells whose CENTROIDS are closer than 6 / 8 / 9, compute_pairwise_fused.h:260,278,299) and the same
// per-pair arithmetic as pair.cuh (the simple kernels kept as an independent cross-check), organised for the SIMT machine:
//
//   k_cell_bounds  per Voronoi cell, a bounding sphere of its current lipids (centre = centroid of the last rebuild) and of
//                  its current proteins (centre = their mean).  A particle can only interact with members of a cell whose
//                  sphere it approaches to within the cutoff, so whole (particle, cell) pairs are skipped by one distance
//                  test; the skip is conservative (margin kCullEps), results are unchanged.
//   k_pair_ll      one thread per lipid.  Phase 1 walks the r<6 stencil of the lipid's cell and only TESTS the cutoff
//                  (r2 < 6.76 && r2 > 1e-5, compute_pairwise_fused.h:109,134); indices that pass go to a per-lane queue in
//                  shared memory (CPU precedent: the reference's implicit-SIMD "enqueue pairs that pass the cutoff" path,
//                  pairwise_kernel_implicit_simd.h:25-100).  Phase 2 evaluates the queued pairs, so the ~60-instruction
//                  force body runs on dense lanes instead of on the ~19 % of lanes that hit in any one iteration.
//                  One-sided (every lipid gathers its own force), no atomics, fixed summation order.
//   k_pair_prot    one thread per protein: protein-protein over r<9 (one-sided) and protein-lipid over r<8.  Hits are rare
//                  (~1 per protein per step on the RBC), so each protein-lipid pair is evaluated ONCE, here, and the lipid
//                  receives its share through atomicAdd (fp32 RED) instead of re-testing every pair from the lipid side.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "pair.cuh"

namespace orbc {

constexpr int kQCap = 32;          // queue slots per lane
constexpr int kLLBlock = 64;
constexpr float kCullEps = 4e-3f;  // slack of the bounding-sphere test (absolute, length units)

struct CullTable {                 // per protein type: largest interaction range against lipids / against the protein types present
    float cut_l[kNType], cut_p[kNType];
};

// ---- bounding spheres ------------------------------------------------------------------------------------------------------
// It also writes every lipid's position relative to the centre of its cell's sphere in HALF precision, two lipids per 16-byte
// record (slots 2p and 2p+1: x pair, y pair, z pair, pad) — the operands of k_pair_ll_h's packed cutoff prefilter — and raises
// rel_flag when a component does not fit the prefilter's error budget (|rel| >= 8), which sends that step to k_pair_ll.
constexpr float kRelMax = 8.0f;
__global__ void k_cell_bounds(const float4 *__restrict__ centroid, int n_cells, const int *__restrict__ cs_l, const float4 *__restrict__ xl,
                              const int *__restrict__ cs_p, const float4 *__restrict__ xp, float4 *__restrict__ lbound, float4 *__restrict__ pbound,
                              const int *__restrict__ need, int need_epoch, uint4 *__restrict__ rel16, int *__restrict__ rel_flag) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    if (need && need[c] != need_epoch) return;             // decomposed run: neither owned nor halo, its particles are stale here
    {
        const float4 q = centroid[c];
        const int b = cs_l[c], e = cs_l[c + 1];
        float r2 = -1.f;                                   // empty cell (or NaN centroid): never passes the test
        float4 ctr = q;
        if (!(q.x == q.x)) { ctr = e > b ? xl[b] : make_float4(0, 0, 0, 0); }
        bool fits = true;
        for (int j = b; j < e; ++j) {
            const float4 p = xl[j];
            const float dx = p.x - ctr.x, dy = p.y - ctr.y, dz = p.z - ctr.z;
            r2 = fmaxf(r2, dx * dx + dy * dy + dz * dz);
            if (rel16) {
                __half *h = reinterpret_cast<__half *>(rel16 + (j >> 1)) + (j & 1);      // x pair at halves 0-1, y at 2-3, z at 4-5
                h[0] = __float2half_rn(dx); h[2] = __float2half_rn(dy); h[4] = __float2half_rn(dz);
                fits = fits && fabsf(dx) < kRelMax && fabsf(dy) < kRelMax && fabsf(dz) < kRelMax;
            }
        }
        if (!fits) atomicExch(rel_flag, 1);
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
// RECHECK: the queue was filled by the half-precision prefilter, which admits a few pairs just outside the cutoff (never the
// reverse); the exact fp32 test of the reference (compute_pairwise_fused.h:109,134) decides here.
template <bool RECHECK>
__device__ __forceinline__ void ll_eval(const LLConst &k, const float4 *__restrict__ xl, const float4 *__restrict__ nl, F3 xi, F3 mi, int j,
                                        float &fx, float &fy, float &fz, float &tx, float &ty, float &tz, float &sB) {
    const float4 xj = __ldg(xl + j), nj = __ldg(nl + j);
    const float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
    const float r2 = dx * dx + dy * dy + dz * dz;
    if (RECHECK && !(r2 < k.cutsq && r2 > 1e-5f)) return;
    const float rinv = rsqrt_fast(r2);
    const float r = r2 * rinv;
    const float ninj = mi.x * nj.x + mi.y * nj.y + mi.z * nj.z;
    const float niu = (mi.x * dx + mi.y * dy + mi.z * dz) * rinv;
    const float nju = (nj.x * dx + nj.y * dy + nj.z * dz) * rinv;
    const float A = fmaf(k.alpha, fmaf(-niu, nju, ninj), k.one_m_alpha);   // 1 + alpha (a - 1)
    const float rc = k.cut - r;
    const float rc2 = rc * rc, rc3 = rc2 * rc, rc4 = rc2 * rc2;
    const float fra = fmaf(k.rep8, rc3 * rc4, k.att4 * (A * rc3));         // 8 rep rc^7 + 4 A att rc^3
    const float aua = k.alpha_att * rc4;                                    // alpha * att * rc^4
    const float auar = aua * rinv;
    const float B = auar * nju, C = auar * niu;
    const float A1 = fmaf(-2.0f * C, nju, fra) * rinv;
    fx = fmaf(A1, dx, fmaf(C, nj.x, fx)); fy = fmaf(A1, dy, fmaf(C, nj.y, fy)); fz = fmaf(A1, dz, fmaf(C, nj.z, fz));
    tx = fmaf(B, dx, fmaf(-aua, nj.x, tx)); ty = fmaf(B, dy, fmaf(-aua, nj.y, ty)); tz = fmaf(B, dz, fmaf(-aua, nj.z, tz));
    sB += B;
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

// One thread per lipid.  A lane's candidates are the members of the cells in the r<6 stencil of its own cell; the lane walks
// them as ONE stream, four at a time, independent of what the other lanes of the warp are looking at (lanes of one warp
// belong to ~3 different cells with different stencils and different cell sizes — aligning them slot by slot would make every
// lane wait for the largest cell of every slot).
//   phase 0  cull: a stencil cell whose bounding sphere (k_cell_bounds) stays further than the cutoff from THIS lipid cannot
//            hold a partner; the surviving cells are a bit mask, so culled cells cost no loop iteration
//   phase 1  test the cutoff, push hits on the lane's queue (shared memory, slot-major: conflict-free)
//   phase 2  drain the queue through the force body on dense lanes; the whole warp drains early if a queue could overflow
template <bool CULL, int MINB>
__global__ void __launch_bounds__(kLLBlock, MINB) k_pair_ll(PairArgs a, const float4 *__restrict__ lbound, const int *__restrict__ run_if, int run_value) {
    if (run_if && *run_if != run_value) return;                  // the packed-prefilter kernel handles this step (or the other way round)
    __shared__ int s_q[kLLBlock / 32][kQCap * 32];
    const int lane = threadIdx.x & 31;
    int *const q = s_q[threadIdx.x >> 5] + lane;
    const int i = a.range[0] + blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < a.range[1];
    const float4 *__restrict__ xl = a.xl;
    const float4 *__restrict__ nl = a.nl;
    const int *__restrict__ cs = a.cs_l;
    float fx = 0, fy = 0, fz = 0, tx = 0, ty = 0, tz = 0, sB = 0;
    F3 xi = {0, 0, 0}, mi = {0, 0, 0};
    const int *st = a.stencil;
    const float cutsq = c_ff.cutsqll;
    const LLConst kc = {c_ff.cutll, 8.0f * c_ff.repll, 4.0f * c_ff.attll, c_ff.alphall, c_ff.alphall * c_ff.attll, 1.0f - c_ff.alphall, c_ff.cutsqll};
    unsigned keep = 0;                                           // stencil slots (<= 32 of the r<6 class) still to visit
    if (live) {
        const float4 xi4 = xl[i], ni4 = nl[i];
        xi = {xi4.x, xi4.y, xi4.z}; mi = {ni4.x, ni4.y, ni4.z};
        const int c = a.cell_l[i];
        const int n6 = min(a.stencil_cnt[c] & 255, 32);
        st += (size_t)c * kStencilStride;
        if (CULL) {
            // four stencil slots at a time, ids first, then their spheres, then the tests: two round trips to memory per four cells
            for (int k0 = 0; k0 < n6; k0 += 4) {
                int c2[4]; float4 b[4];
                #pragma unroll
                for (int u = 0; u < 4; ++u) c2[u] = __ldg(st + min(k0 + u, n6 - 1));
                #pragma unroll
                for (int u = 0; u < 4; ++u) b[u] = __ldg(lbound + c2[u]);
                #pragma unroll
                for (int u = 0; u < 4; ++u) if (k0 + u < n6 && !culled(b[u], xi.x, xi.y, xi.z, kc.cut)) keep |= 1u << (k0 + u);
            }
        } else keep = n6 >= 32 ? 0xffffffffu : (1u << n6) - 1u;
    }
    // the lane's hit queue, addressed with 32-bit shared-window addresses (slot stride = 32 lanes x 4 B)
    const unsigned q0 = (unsigned)__cvta_generic_to_shared(q);
    const unsigned q_full = q0 + (kQCap - 4) * 128;              // a group of four always fits below this mark
    unsigned qp = q0;
    // two-deep prefetch so that the stream never waits for the stencil: (jb_n, len_n) = member range of the next cell to
    // visit, c2_nn = id of the one after it; n_next = cells not yet entered
    int n_next = __popc(keep), jb_n = 0, len_n = 0, c2_nn = 0;
    if (n_next > 0) { const int c2 = __ldg(st + (__ffs(keep) - 1)); keep &= keep - 1; jb_n = __ldg(cs + c2); len_n = __ldg(cs + c2 + 1) - jb_n; }
    if (n_next > 1) { c2_nn = __ldg(st + (__ffs(keep) - 1)); keep &= keep - 1; }
    // cur stays a valid index for idle lanes: loads are unconditional, only the queue push is predicated (the arrays are
    // allocated with 64 spare elements, so reading up to three elements past a cell's range is always in bounds)
    int cur = 0, rem = 0;
    for (;;) {
        if (rem <= 0 && n_next > 0) {                            // advance to the next cell of the stencil
            cur = jb_n; rem = len_n; --n_next;
            if (n_next > 0) { jb_n = __ldg(cs + c2_nn); len_n = __ldg(cs + c2_nn + 1) - jb_n; }
            if (n_next > 1) { c2_nn = __ldg(st + (__ffs(keep) - 1)); keep &= keep - 1; }
        }
        if (!__any_sync(0xffffffffu, rem > 0 || n_next > 0)) break;
        if (__any_sync(0xffffffffu, qp > q_full)) {              // make room: every lane drains its queue (dense)
            for (unsigned e = q0; e < qp; e += 128) ll_eval<false>(kc, xl, nl, xi, mi, lds_i32(e), fx, fy, fz, tx, ty, tz, sB);
            qp = q0;
        }
        const float4 *__restrict__ p = xl + cur;
        float4 xj[4];
        #pragma unroll
        for (int u = 0; u < 4; ++u) xj[u] = __ldg(p + u);
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float dx = xi.x - xj[u].x, dy = xi.y - xj[u].y, dz = xi.z - xj[u].z;
            const float r2 = dx * dx + dy * dy + dz * dz;
            if (u < rem && r2 < cutsq && r2 > 1e-5f) { sts_i32(qp, cur + u); qp += 128; }
        }
        if (rem > 0) cur += 4;
        rem -= 4;
    }
    for (unsigned e = q0; e < qp; e += 128) ll_eval<false>(kc, xl, nl, xi, mi, lds_i32(e), fx, fy, fz, tx, ty, tz, sB);
    ll_finish(a, i, live, st, xi, mi, fx, fy, fz, tx, ty, tz, sB);
}

// ---- k_pair_ll_r: the same kernel over precomputed candidate RUNS ---------------------------------------------------------------------
// The candidates of a lipid are the members of the r<6 stencil cells of its cell, visited in ascending cell id.  Cells are
// numbered in Morton order and particles are stored sorted by cell, so neighbouring stencil cells usually hold neighbouring
// slot ranges: k_lipid_runs merges them, once per rebuild, into a few (first slot, length) runs per cell.  The stream then
// advances by one 8-byte load per run instead of three dependent loads per cell (stencil id -> cell_start pair), wastes fewer
// lanes on the partial group at the end of every cell, and the loop control shrinks accordingly.  Same candidates, same order,
// same hits as k_pair_ll: results are bit-identical.
constexpr int kRunStride = 32;     // a cell has at most 32 stencil cells of the r<6 class (k_stencil_build raises a flag otherwise)
__global__ void k_lipid_runs(int cb, int ce, const int *__restrict__ stencil, const int *__restrict__ stencil_cnt, const int *__restrict__ cs_l,
                             int2 *__restrict__ lruns, int *__restrict__ lrun_cnt) {
    const int c = cb + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ce) return;
    const int n6 = min(stencil_cnt[c] & 255, 32);
    const int *st = stencil + (size_t)c * kStencilStride;
    int2 *out = lruns + (size_t)c * kRunStride;
    int nr = 0, rb = 0, re = -1;
    for (int k = 0; k < n6; ++k) {
        const int c2 = st[k];
        const int b = cs_l[c2], e = cs_l[c2 + 1];
        if (e <= b) continue;
        if (b == re) { re = e; continue; }
        if (re > rb) out[nr++] = make_int2(rb, re - rb);
        rb = b; re = e;
    }
    if (re > rb) out[nr++] = make_int2(rb, re - rb);
    lrun_cnt[c] = nr;
}

// W = candidates per lane and iteration.  Measured on the full RBC (B200): W = 4 with 20 resident blocks 470 us (k_pair_ll: 508);
// W = 8 505-515 us (longer partial groups, 64 registers); 24 resident blocks at 40 registers 538 us (spills); an L1 prefetch 4-16
// candidates ahead of the stream changes nothing.
template <int MINB, int W>
__global__ void __launch_bounds__(kLLBlock, MINB) k_pair_ll_r(PairArgs a, const int2 *__restrict__ lruns, const int *__restrict__ lrun_cnt) {
    __shared__ int s_q[kLLBlock / 32][kQCap * 32];
    const int lane = threadIdx.x & 31;
    int *const q = s_q[threadIdx.x >> 5] + lane;
    const int i = a.range[0] + blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < a.range[1];
    const float4 *__restrict__ xl = a.xl;
    const float4 *__restrict__ nl = a.nl;
    float fx = 0, fy = 0, fz = 0, tx = 0, ty = 0, tz = 0, sB = 0;
    F3 xi = {0, 0, 0}, mi = {0, 0, 0};
    const int *st = a.stencil;
    const LLConst kc = {c_ff.cutll, 8.0f * c_ff.repll, 4.0f * c_ff.attll, c_ff.alphall, c_ff.alphall * c_ff.attll, 1.0f - c_ff.alphall, c_ff.cutsqll};
    // r2 > 1e-5 && r2 < cutsq (compute_pairwise_fused.h:109,134) as ONE unsigned comparison of the bit patterns: r2 is a sum of
    // squares (never negative), and non-negative floats order like their bits; a NaN lies above every finite pattern
    const unsigned lo_bits = __float_as_uint(1e-5f) + 1u, span = __float_as_uint(c_ff.cutsqll) - lo_bits;
    const int2 *rp = lruns;
    int nr = 0;
    if (live) {
        const float4 xi4 = xl[i], ni4 = nl[i];
        xi = {xi4.x, xi4.y, xi4.z}; mi = {ni4.x, ni4.y, ni4.z};
        const int c = a.cell_l[i];
        st += (size_t)c * kStencilStride;
        rp += (size_t)c * kRunStride;
        nr = __ldg(lrun_cnt + c);
    }
    const unsigned q0 = (unsigned)__cvta_generic_to_shared(q);
    const unsigned q_full = q0 + (kQCap - W) * 128;              // a group of W always fits below this mark
    unsigned qp = q0;
    int2 nx = make_int2(0, 0);                                   // the next run, loaded one advance ahead
    if (nr > 0) nx = __ldg(rp);
    int k = 0, cur = 0, rem = 0;
    for (;;) {
        if (rem <= 0 && k < nr) { cur = nx.x; rem = nx.y; ++k; if (k < nr) nx = __ldg(rp + k); }
        if (!__any_sync(0xffffffffu, rem > 0)) break;
        if (__any_sync(0xffffffffu, qp > q_full)) {              // make room: every lane drains its queue (dense)
            for (unsigned e = q0; e < qp; e += 128) ll_eval<false>(kc, xl, nl, xi, mi, lds_i32(e), fx, fy, fz, tx, ty, tz, sB);
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
    for (unsigned e = q0; e < qp; e += 128) ll_eval<false>(kc, xl, nl, xi, mi, lds_i32(e), fx, fy, fz, tx, ty, tz, sB);
    ll_finish(a, i, live, st, xi, mi, fx, fy, fz, tx, ty, tz, sB);
}

// ---- k_pair_ll_h (EXPERIMENTAL, off by default: option "ll_half"): the cutoff test of phase 1 in packed half precision ---------
// Measured on the full RBC (profiles/r01_half_prefilter.txt): exact (hits, order and forces bit-identical to k_pair_ll), the
// packed test is 25 % cheaper than the fp32 one (16 instructions per two partners), but the kernel as a whole is SLOWER (540-590
// vs 487 us): queues fill 3 % fuller and drain less densely, the re-test and the per-cell frame shift cost instructions, and the
// 2-byte scattered stores that produce the records add 70 us to k_cell_bounds.  Kept as a measured negative result.
// Phase 1 is two thirds of k_pair_ll's instructions and 81 % of the pairs it tests are beyond the cutoff.  Here it runs on the
// cell-relative half-precision records written by k_cell_bounds, two partners per instruction (sub / mul / fma.f16x2, one
// setp.lt.f16x2 for both): 6.5 instead of 11 instructions per candidate and half the bytes.  It is a PREFILTER: the limit is
// cutsq + kHalfMargin, so that no pair the reference's fp32 test (r2 < 6.76) admits can be missed — with |rel| < 8 the
// partner's record is off by <= 2^-8, the lane's own position (fp32, shifted into the partner cell's frame, rounded once) by
// <= 2^-8, the difference by <= 2^-10 where it matters (|d| < 4), i.e. <= 0.009 per component and <= 0.09 in r2, plus three
// half-precision roundings of r2 (<= 0.012).  Phase 2 re-tests every queued pair exactly (ll_eval<true>), so the hits, their
// order and the forces are bit-identical to k_pair_ll's.  k_cell_bounds raises rel_flag when a lipid strays further than 8
// from its cell's origin; that step then runs through k_pair_ll instead (both kernels are launched, one returns at once).
constexpr float kHalfMargin = 0.2f;
__device__ __forceinline__ unsigned h2_bcast(float v) { const __half2 h = __float2half2_rn(v); return *reinterpret_cast<const unsigned *>(&h); }

// one record (two partners, slots j0 and j0 + 1 = positions s0, s0 + 1 of the lane's current run): test both, queue the hits
#define ORBC_LL_PAIR(REC, S0, FIRST)                                                                                            \
    asm volatile("{\n"                                                                                                          \
                 ".reg .b32 dx, dy, dz, r2;\n"                                                                                  \
                 ".reg .pred p, q;\n"                                                                                           \
                 "sub.f16x2 dx, %1, %4;\n"                                                                                      \
                 "sub.f16x2 dy, %2, %5;\n"                                                                                      \
                 "sub.f16x2 dz, %3, %6;\n"                                                                                      \
                 "mul.f16x2 r2, dx, dx;\n"                                                                                      \
                 "fma.rn.f16x2 r2, dy, dy, r2;\n"                                                                               \
                 "fma.rn.f16x2 r2, dz, dz, r2;\n"                                                                               \
                 "setp.lt.f16x2 p|q, r2, %7;\n"                                                                                 \
                 "setp.gt.and.s32 p, %8, %9, p;\n"                                                                              \
                 "setp.gt.and.s32 q, %8, %10, q;\n"                                                                             \
                 "setp.eq.and.s32 p, %13, 0, p;\n"                                                                              \
                 "@p st.shared.b32 [%0], %11;\n"                                                                                \
                 "@p add.u32 %0, %0, 128;\n"                                                                                    \
                 "@q st.shared.b32 [%0], %12;\n"                                                                                \
                 "@q add.u32 %0, %0, 128;\n"                                                                                    \
                 "}"                                                                                                            \
                 : "+r"(qp)                                                                                                     \
                 : "r"(xh), "r"(yh), "r"(zh), "r"((REC).x), "r"((REC).y), "r"((REC).z), "r"(lim), "r"(rem), "n"(S0), "n"((S0) + 1),      \
                   "r"(cur + (S0)), "r"(cur + (S0) + 1), "r"((FIRST) ? skip : 0)                                                  \
                 : "memory")

template <int MINB>
__global__ void __launch_bounds__(kLLBlock, MINB) k_pair_ll_h(PairArgs a, const float4 *__restrict__ lbound, const uint4 *__restrict__ rel16,
                                                               const int *__restrict__ run_if, int run_value) {
    if (run_if && *run_if != run_value) return;
    __shared__ int s_q[kLLBlock / 32][kQCap * 32];
    const int lane = threadIdx.x & 31;
    int *const q = s_q[threadIdx.x >> 5] + lane;
    const int i = a.range[0] + blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < a.range[1];
    const float4 *__restrict__ xl = a.xl;
    const float4 *__restrict__ nl = a.nl;
    const int *__restrict__ cs = a.cs_l;
    float fx = 0, fy = 0, fz = 0, tx = 0, ty = 0, tz = 0, sB = 0;
    F3 xi = {0, 0, 0}, mi = {0, 0, 0};
    const int *st = a.stencil;
    const LLConst kc = {c_ff.cutll, 8.0f * c_ff.repll, 4.0f * c_ff.attll, c_ff.alphall, c_ff.alphall * c_ff.attll, 1.0f - c_ff.alphall, c_ff.cutsqll};
    const unsigned lim = h2_bcast(c_ff.cutsqll + kHalfMargin);
    unsigned keep = 0;                                           // stencil slots (<= 32 of the r<6 class) still to visit
    if (live) {
        const float4 xi4 = xl[i], ni4 = nl[i];
        xi = {xi4.x, xi4.y, xi4.z}; mi = {ni4.x, ni4.y, ni4.z};
        const int c = a.cell_l[i];
        const int n6 = min(a.stencil_cnt[c] & 255, 32);
        st += (size_t)c * kStencilStride;
        keep = n6 >= 32 ? 0xffffffffu : (1u << n6) - 1u;
    }
    const unsigned q0 = (unsigned)__cvta_generic_to_shared(q);
    const unsigned q_full = q0 + (kQCap - 8) * 128;              // a group of eight always fits below this mark
    unsigned qp = q0;
    // two-deep prefetch: (c2_n, jb_n, len_n) = id and member range of the next cell, c2_nn = id of the one after it; the sphere
    // record of a cell (its frame origin) is pulled into L1 one advance before it is read
    int n_next = __popc(keep), c2_n = 0, jb_n = 0, len_n = 0, c2_nn = 0;
    if (n_next > 0) { c2_n = __ldg(st + (__ffs(keep) - 1)); keep &= keep - 1; jb_n = __ldg(cs + c2_n); len_n = __ldg(cs + c2_n + 1) - jb_n; }
    if (n_next > 1) { c2_nn = __ldg(st + (__ffs(keep) - 1)); keep &= keep - 1; }
    // cur = even slot the current run is read from, rem = slots of the run still to test counted from cur, skip = 1 when the run
    // starts on an odd slot (the first half of its first record belongs to the previous cell)
    int cur = 0, rem = 0, skip = 0;
    unsigned xh = 0, yh = 0, zh = 0;                             // the lane's position in the current cell's frame, half2-broadcast
    for (;;) {
        if (rem <= 0 && n_next > 0) {                            // advance to the next cell of the stencil
            const float4 o = __ldg(lbound + c2_n);
            xh = h2_bcast(xi.x - o.x); yh = h2_bcast(xi.y - o.y); zh = h2_bcast(xi.z - o.z);
            skip = jb_n & 1; cur = jb_n - skip; rem = len_n + skip; --n_next;
            if (n_next > 0) {
                c2_n = c2_nn; jb_n = __ldg(cs + c2_n); len_n = __ldg(cs + c2_n + 1) - jb_n;
                asm volatile("prefetch.global.L1 [%0];" :: "l"(lbound + c2_n));
            }
            if (n_next > 1) { c2_nn = __ldg(st + (__ffs(keep) - 1)); keep &= keep - 1; }
        }
        if (!__any_sync(0xffffffffu, rem > 0 || n_next > 0)) break;
        if (__any_sync(0xffffffffu, qp > q_full)) {              // make room: every lane drains its queue (dense)
            for (unsigned e = q0; e < qp; e += 128) ll_eval<true>(kc, xl, nl, xi, mi, lds_i32(e), fx, fy, fz, tx, ty, tz, sB);
            qp = q0;
        }
        const uint4 *__restrict__ p = rel16 + (cur >> 1);
        uint4 rec[4];
        #pragma unroll
        for (int u = 0; u < 4; ++u) rec[u] = __ldg(p + u);
        ORBC_LL_PAIR(rec[0], 0, true);
        ORBC_LL_PAIR(rec[1], 2, false);
        ORBC_LL_PAIR(rec[2], 4, false);
        ORBC_LL_PAIR(rec[3], 6, false);
        skip = 0;
        if (rem > 0) cur += 8;
        rem -= 8;
    }
    for (unsigned e = q0; e < qp; e += 128) ll_eval<true>(kc, xl, nl, xi, mi, lds_i32(e), fx, fy, fz, tx, ty, tz, sB);
    ll_finish(a, i, live, st, xi, mi, fx, fy, fz, tx, ty, tz, sB);
}
#undef ORBC_LL_PAIR

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
template <int LPP>
__global__ void __launch_bounds__(kPBlock, 16) k_pair_prot(PairArgs a, const float4 *__restrict__ lbound, const float4 *__restrict__ pbound, CullTable ct,
                                                             const int *__restrict__ porder) {
    __shared__ float s_cutsqpp[36], s_ljcutsq[36];
    __shared__ int s_jb[2][kPBlock / 32][kRangeCap * 32];
    __shared__ unsigned short s_len[2][kPBlock / 32][kRangeCap * 32];
    for (int k = threadIdx.x; k < 36; k += blockDim.x) { s_cutsqpp[k] = c_ff.cutsqpp[k]; s_ljcutsq[k] = c_ff.lj_cutsq[k]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int *const ljb = s_jb[0][w] + lane, *const pjb = s_jb[1][w] + lane;
    unsigned short *const llen = s_len[0][w] + lane, *const plen = s_len[1][w] + lane;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int pid = tid / LPP, sub = tid % LPP;
    const bool live = pid < a.range[3] - a.range[2];
    const int i = live ? porder[pid] : 0;
    const int l0 = a.range[0], l1 = a.range[1];                  // lipids of other ranks get their share from k_pair_lipid<true> over there
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
    const float cull_l = type1 == 0 ? ct.cut_l[0] : type1 == 1 ? ct.cut_l[1] : type1 == 2 ? ct.cut_l[2] : type1 == 3 ? ct.cut_l[3] : type1 == 4 ? ct.cut_l[4] : ct.cut_l[5];
    const float cull_p = type1 == 0 ? ct.cut_p[0] : type1 == 1 ? ct.cut_p[1] : type1 == 2 ? ct.cut_p[2] : type1 == 3 ? ct.cut_p[3] : type1 == 4 ? ct.cut_p[4] : ct.cut_p[5];
    const float testsq_l = fmaxf(cutsq, ljcut);
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
            if (in9 && pe > pb && cull_p > 0.f && !culled(bp, xi.x, xi.y, xi.z, cull_p)) { pjb[np_ * 32] = pb; plen[np_ * 32] = (unsigned short)min(pe - pb, 65535); ++np_; }
            if (in8 && le > lb && cull_l > 0.f && !culled(bl, xi.x, xi.y, xi.z, cull_l)) { ljb[nl_ * 32] = lb; llen[nl_ * 32] = (unsigned short)min(le - lb, 65535); ++nl_; }
        }
        // ---- protein-lipid: evaluated once, here; the lipid gets its share by atomics (compute_pairwise_fused.h:143-179,278-297) ---
        // every lane streams through ITS surviving ranges four candidates at a time (same scheme as k_pair_ll)
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
                    const F3 d = {xi.x - xj.x, xi.y - xj.y, xi.z - xj.z};          // x_protein - x_lipid (compute_pairwise_fused.h:167)
                    const float r2 = dot3(d, d);
                    if (u < rem && r2 < testsq_l && r2 > 1e-5f) {
                        const int j = cur + u;
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
                    const F3 d = {xi.x - xj.x, xi.y - xj.y, xi.z - xj.z};
                    const float r2 = dot3(d, d);
                    if (u < rem && r2 > 1e-5f && r2 < cull_p * cull_p) {
                        const int type12 = type1 + __float_as_int(xj.w) * kNType;
                        if (r2 < s_cutsqpp[type12]) {
                            const F3 f = rep8(c_ff.cutpp[type12], c_ff.reppp[type12], d, r2);
                            fx += f.x; fy += f.y; fz += f.z;
                        } else if (r2 < s_ljcutsq[type12]) {
                            const F3 f = lj126(c_ff.lj_lj1[type12], c_ff.lj_lj2[type12], d, r2);
                            fx += f.x; fy += f.y; fz += f.z;
                        }
                    }
                }
                if (rem > 0) cur += 2;
                rem -= 2;
            }
        }
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
}

} // namespace orbc
