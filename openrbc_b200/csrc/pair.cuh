// pair.cuh — pairwise forces: lipid-lipid, protein-lipid (+LJ), protein-protein (+LJ), and harmonic bonds.
//
// Replaces compute_pairwise_fused.h:238-320 (driver), pairwise_kernel.h:30-68, pairwise_kernel_fused.h:22-97 (physics),
// compute_bonded.h:89-146.  The reference evaluates each cell pair once with Newton's third law inside a thread's cell
// range and twice, one-sidedly, across ranges (compute_pairwise_fused.h:264-275,287-295,303-314); the device takes the
// one-sided form everywhere: every particle gathers its own force over exactly the reference's candidate set
// (cells whose CENTROIDS are closer than 6 / 8 / 9, not the physical cutoff sphere), so no force ever needs an atomic.
#pragma once
#include "common.cuh"

namespace orbc {

__constant__ orbc_forcefield c_ff;

struct PairArgs {
    const float4 *xl, *nl; const int *cs_l; const int *cell_l; int n_l;
    const float4 *xp, *np; const int *cs_p; const int *cell_p; int n_p;
    const int *stencil, *stencil_cnt;
    float4 *fl, *tl, *fp, *tp;
    const int *range;      // {l0, l1, p0, p1}: the particle slots this GPU computes (everything on a single GPU)
    int cb, ce, world;     // owned cells of a decomposed run ([0, n_cells) and 1 otherwise)
    const unsigned char *dest_mask;   // decomposed: per cell, the other ranks that own a cell of its r<9 stencil
    int accumulate;        // 1: f, t += (the reference's semantics); 0: f, t = (orbc_run_langevin, where they are known to be dead)
    unsigned long long *counters;   // k_pair_lipid / k_pair_protein (the cross-check kernels) count their tests and hits here: [1..6] = LL, PL, PP tests / hits
};

struct F3 { float x, y, z; };
__device__ __forceinline__ float dot3(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// poly 4-8 anisotropic pair (pairwise_kernel.h:30-68 / pairwise_kernel_fused.h:22-61).  d = x1 - x2, mu1/mu2 directors.
// f  : force on particle 1 (particle 2 gets -f)
// q1 : alphaua * pnj  -> t[1] -= q1          q2 : alphaua * pni  -> t[2] -= q2
__device__ __forceinline__ void poly48(float cut, float att, float rep, float alpha, F3 d, float r2, F3 mu1, F3 mu2, F3 &f, F3 &q1, F3 &q2) {
    const float rinv = rsqrtf(r2);
    const float r = r2 * rinv;
    const F3 u = {d.x * rinv, d.y * rinv, d.z * rinv};
    const float ninj = dot3(mu1, mu2), niu = dot3(mu1, u), nju = dot3(mu2, u);
    const float a = ninj - niu * nju;
    const float A = 1.0f + alpha * (a - 1.0f);
    const float rc = cut - r;
    const float rc3 = rc * rc * rc, rc4 = rc * rc3, rc7 = rc3 * rc4;
    const F3 pni = {mu1.x - niu * u.x, mu1.y - niu * u.y, mu1.z - niu * u.z};
    const F3 pnj = {mu2.x - nju * u.x, mu2.y - nju * u.y, mu2.z - nju * u.z};
    const float ua = att * rc4;
    const float alphaua = alpha * ua;
    const float alphauar = alphaua * rinv;
    const float fra = 8.0f * rep * rc7 + A * 4.0f * att * rc3;
    f = {fra * u.x + alphauar * (nju * pni.x + niu * pnj.x),
         fra * u.y + alphauar * (nju * pni.y + niu * pnj.y),
         fra * u.z + alphauar * (nju * pni.z + niu * pnj.z)};
    q1 = {alphaua * pnj.x, alphaua * pnj.y, alphaua * pnj.z};
    q2 = {alphaua * pni.x, alphaua * pni.y, alphaua * pni.z};
}
// pairwise_kernel_fused.h:63-77
__device__ __forceinline__ F3 lj126(float lj1, float lj2, F3 d, float r2) {
    const float r2inv = 1.0f / r2;
    const float r6inv = r2inv * r2inv * r2inv;
    const float fpair = r6inv * (lj1 * r6inv - lj2) * r2inv;
    return {d.x * fpair, d.y * fpair, d.z * fpair};
}
// pairwise_kernel_fused.h:79-97
__device__ __forceinline__ F3 rep8(float cut, float rep, F3 d, float r2) {
    const float rinv = rsqrtf(r2);
    const float rc = cut - r2 * rinv;
    const float rc3 = rc * rc * rc, rc7 = rc3 * rc3 * rc;
    const float s = 8.0f * rep * rc7 * rinv;
    return {d.x * s, d.y * s, d.z * s};
}

// ---- v1: one thread per lipid -------------------------------------------------------------------------------------------------
// lipid i gathers: LL over the r<6 stencil of its cell (lipid_lipid::rmax, compute_pairwise_fused.h:92), protein-lipid over the
// r<8 stencil (prote_lipid::rmax, :144) as the lipid side of protein_lipid_omp / lennard_jones_omp.
__global__ void __launch_bounds__(128) k_pair_lipid(PairArgs a) {
    const int i = a.range[0] + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.range[1]) return;
    const float4 xi4 = a.xl[i], ni4 = a.nl[i];
    const F3 xi = {xi4.x, xi4.y, xi4.z}, mi = {ni4.x, ni4.y, ni4.z};
    const int c = a.cell_l[i];
    const int cnt = a.stencil_cnt[c];
    const int n6 = cnt & 255, n8 = (cnt >> 8) & 255;
    const int *st = a.stencil + (size_t)c * kStencilStride;
    float fx = 0, fy = 0, fz = 0, tx = 0, ty = 0, tz = 0;
    const float cutsqll = c_ff.cutsqll;
    unsigned tests = 0, hits = 0;
    for (int k = 0; k < n8; ++k) {
        const int c2 = st[k];
        if (k < n6) {
            const int jb = a.cs_l[c2], je = a.cs_l[c2 + 1];
            tests += je - jb;
            for (int j = jb; j < je; ++j) {
                const float4 xj = a.xl[j];
                const F3 d = {xi.x - xj.x, xi.y - xj.y, xi.z - xj.z};
                const float r2 = dot3(d, d);
                if (r2 < cutsqll && r2 > 1e-5f) {
                    ++hits;
                    const float4 nj = a.nl[j];
                    F3 f, q1, q2;
                    poly48(c_ff.cutll, c_ff.attll, c_ff.repll, c_ff.alphall, d, r2, mi, {nj.x, nj.y, nj.z}, f, q1, q2);
                    fx += f.x; fy += f.y; fz += f.z; tx -= q1.x; ty -= q1.y; tz -= q1.z;
                }
            }
        }
        if (a.n_p) {
            const int jb = a.cs_p[c2], je = a.cs_p[c2 + 1];
            for (int j = jb; j < je; ++j) {
                const float4 xj = a.xp[j];
                const int type = __float_as_int(xj.w);
                const F3 d = {xj.x - xi.x, xj.y - xi.y, xj.z - xi.z};      // x_protein - x_lipid (compute_pairwise_fused.h:167)
                const float r2 = dot3(d, d);
                if (r2 < c_ff.cutsqlp[type] && r2 > 1e-5f) {
                    const float4 nj = a.np[j];
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
    float4 f = a.fl[i], t = a.tl[i];
    f.x += fx; f.y += fy; f.z += fz; t.x += tx; t.y += ty; t.z += tz;
    a.fl[i] = f; a.tl[i] = t;
    if (a.counters) { atomicAdd(a.counters + 1, (unsigned long long)tests); atomicAdd(a.counters + 2, (unsigned long long)hits); }   // one-sided counts
}

// ---- v1: one thread per protein: protein-protein over r<9 (prote_prote::rmax, :183), protein side of protein-lipid over r<8 ----
__global__ void __launch_bounds__(128) k_pair_protein(PairArgs a) {
    const int i = a.range[2] + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.range[3]) return;
    const float4 xi4 = a.xp[i], ni4 = a.np[i];
    const F3 xi = {xi4.x, xi4.y, xi4.z}, mi = {ni4.x, ni4.y, ni4.z};
    const int type1 = __float_as_int(xi4.w);
    const int c = a.cell_p[i];
    const int cnt = a.stencil_cnt[c];
    const int n8 = (cnt >> 8) & 255, n9 = cnt >> 16;
    const int *st = a.stencil + (size_t)c * kStencilStride;
    float fx = 0, fy = 0, fz = 0, tx = 0, ty = 0, tz = 0;
    const float cutsq = c_ff.cutsqlp[type1], ljcut = c_ff.lj_cutsq[type1];
    unsigned t_pl = 0, h_pl = 0, t_pp = 0, h_pp = 0;
    for (int k = 0; k < n9; ++k) {
        const int c2 = st[k];
        {
            const int jb = a.cs_p[c2], je = a.cs_p[c2 + 1];
            t_pp += je - jb;
            for (int j = jb; j < je; ++j) {
                const float4 xj = a.xp[j];
                const F3 d = {xi.x - xj.x, xi.y - xj.y, xi.z - xj.z};
                const float r2 = dot3(d, d);
                const int type12 = type1 + __float_as_int(xj.w) * kNType;
                if (r2 < c_ff.cutsqpp[type12] && r2 > 1e-5f) {
                    const F3 f = rep8(c_ff.cutpp[type12], c_ff.reppp[type12], d, r2);
                    fx += f.x; fy += f.y; fz += f.z; ++h_pp;
                } else if (r2 < c_ff.lj_cutsq[type12] && r2 > 1e-5f) {
                    const F3 f = lj126(c_ff.lj_lj1[type12], c_ff.lj_lj2[type12], d, r2);
                    fx += f.x; fy += f.y; fz += f.z; ++h_pp;
                }
            }
        }
        if (k < n8) {
            const int jb = a.cs_l[c2], je = a.cs_l[c2 + 1];
            t_pl += je - jb;
            for (int j = jb; j < je; ++j) {
                const float4 xj = a.xl[j];
                const F3 d = {xi.x - xj.x, xi.y - xj.y, xi.z - xj.z};
                const float r2 = dot3(d, d);
                if (r2 < cutsq && r2 > 1e-5f) {
                    const float4 nj = a.nl[j];
                    F3 f, q1, q2;
                    poly48(c_ff.cutlp[type1], c_ff.attlp[type1], c_ff.replp[type1], c_ff.alphalp[type1], d, r2, mi, {nj.x, nj.y, nj.z}, f, q1, q2);
                    fx += f.x; fy += f.y; fz += f.z; tx -= q1.x; ty -= q1.y; tz -= q1.z; ++h_pl;
                } else if (r2 < ljcut && r2 > 1e-5f) {
                    const F3 f = lj126(c_ff.lj_lj1[type1], c_ff.lj_lj2[type1], d, r2);
                    fx += f.x; fy += f.y; fz += f.z; ++h_pl;
                }
            }
        }
    }
    float4 f = a.fp[i], t = a.tp[i];
    f.x += fx; f.y += fy; f.z += fz; t.x += tx; t.y += ty; t.z += tz;
    a.fp[i] = f; a.tp[i] = t;
    if (a.counters) {
        atomicAdd(a.counters + 3, (unsigned long long)t_pl); atomicAdd(a.counters + 4, (unsigned long long)h_pl);
        atomicAdd(a.counters + 5, (unsigned long long)t_pp); atomicAdd(a.counters + 6, (unsigned long long)h_pp);   // PP one-sided
    }
}

// ---- compute_bonded.h:89-146: F = K (1 - r0 / |dx|) dx, +F on atom i, -F on atom j ----------------------------------------------
// Decomposed runs: every rank walks the whole bond list and applies the force to the atoms it owns (slots [p0, p1)); a bond
// that straddles two ranks is evaluated by both, one-sidedly, from the halo copy of the partner.
// `my_bonds` (decomposed): [0] = count, then the indices of the bonds with an owned atom (multi.cuh k_bond_mask).
__global__ void k_bonded(const int *__restrict__ bonds, size_t n_bonds, const int *__restrict__ tag2idx, const float4 *__restrict__ x, float4 *__restrict__ f,
                         const int *__restrict__ range, const int *__restrict__ my_bonds) {
    size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (my_bonds) { if (l >= (size_t)my_bonds[0]) return; l = (size_t)my_bonds[1 + l]; }
    else if (l >= n_bonds) return;
    const int type = bonds[3 * l], p1 = tag2idx[bonds[3 * l + 1]], p2 = tag2idx[bonds[3 * l + 2]];
    const int lo = range[2], hi = range[3];
    const bool own1 = p1 >= lo && p1 < hi, own2 = p2 >= lo && p2 < hi;
    if (!own1 && !own2) return;
    const float4 a = x[p1], b = x[p2];
    const float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
    const float rinv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);   // K (1 - r0/r) cancels near r0: keep the exact reciprocal root
    const float s = c_ff.K[type] * (1.0f - c_ff.r0[type] * rinv);
    if (own1) { atomicAdd(&f[p1].x, s * dx); atomicAdd(&f[p1].y, s * dy); atomicAdd(&f[p1].z, s * dz); }
    if (own2) { atomicAdd(&f[p2].x, -s * dx); atomicAdd(&f[p2].y, -s * dy); atomicAdd(&f[p2].z, -s * dz); }
}

} // namespace orbc
