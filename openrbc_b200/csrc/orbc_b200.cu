// orbc_b200.cu — the C ABI of include/orbc_b200.h over the sm_100a kernels in rebuild.cuh / pair.cuh / pair_tiled.cuh /
// integrate.cuh.  One context per GPU, every launch on the context's stream, no host synchronisation inside the
// per-step calls (only download / reductions returned to the host synchronise).
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"
#include "primitives.cuh"
#include "rebuild.cuh"
#include "pair.cuh"
#include "pair_queue.cuh"
#include "pair_tile.cuh"
#include "integrate.cuh"
#include "multi.cuh"

using namespace orbc;

namespace {

constexpr int kBlock = 256;

// ---- packing between the host's strided vect arrays and the device's float4 SoA ------------------------------------------------
__global__ void k_pack4(const float *__restrict__ src, size_t stride, size_t n, float4 *__restrict__ dst, const int *__restrict__ w) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = src + i * stride;
    dst[i] = make_float4(p[0], p[1], p[2], w ? __int_as_float(w[i]) : 0.f);
}
__global__ void k_zero4(float4 *__restrict__ dst, size_t n, const int *__restrict__ w) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = make_float4(0.f, 0.f, 0.f, w ? __int_as_float(w[i]) : 0.f);
}
__global__ void k_set3(const float *__restrict__ src, size_t stride, size_t n, float4 *__restrict__ dst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = src + i * stride;
    float4 d = dst[i]; d.x = p[0]; d.y = p[1]; d.z = p[2]; dst[i] = d;
}
__global__ void k_unpack3(const float4 *__restrict__ src, size_t n, size_t stride, float *__restrict__ dst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 s = src[i];
    float *p = dst + i * stride;
    p[0] = s.x; p[1] = s.y; p[2] = s.z;
    for (size_t d = 3; d < stride; ++d) p[d] = 0.f;
}
__global__ void k_unpack_w(const float4 *__restrict__ src, size_t n, int *__restrict__ dst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float_as_int(src[i].w);
}
__global__ void k_morton_keys_only(const float4 *__restrict__ pts, int n, uint32_t *__restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = morton_encode(pts[i]);
}

// ---- save_frame (trajectory.h:61-105): the frame in the .orbc byte layout, assembled on the device -------------------------------------
// FRAMEBEG nstep(int) NATOM n(size_t) IDENTITY (tag,type)* [POSITION xyz*] [VELOCITY xyz*] [ROTATION xyz*] [VORONOI cell*] [FORCE xyz*]
// FRAMEEND, titles NUL-padded to 8 bytes, lipids before proteins.  Section payloads start at 4-byte aligned offsets.
struct FrameLayout {
    size_t off_id, off_x, off_v, off_n, off_aff, off_f, off_end;   // payload offsets (0 = section absent); off_end = offset of "FRAMEEND"
    size_t n_l, n_p;
    int nstep, tag_base;
    int l0, l1, p0, p1;                                            // slots to write (all of them on one GPU, the owned ones on a rank)
};
__device__ __forceinline__ void put_title(unsigned char *dst, const char *t) {
    int k = 0;
    for (; k < 8 && t[k]; ++k) dst[k] = (unsigned char)t[k];
    for (; k < 8; ++k) dst[k] = 0;
}
__device__ __forceinline__ void put3(unsigned char *base, size_t off, size_t i, float4 v) {
    float *p = reinterpret_cast<float *>(base + off) + 3 * i;
    p[0] = v.x; p[1] = v.y; p[2] = v.z;
}
__global__ void __launch_bounds__(256) k_frame_pack(FrameLayout L, unsigned char *__restrict__ out,
                                                     const float4 *__restrict__ xl, const float4 *__restrict__ vl, const float4 *__restrict__ nl, const float4 *__restrict__ fl, const int *__restrict__ cl,
                                                     const float4 *__restrict__ xp, const float4 *__restrict__ vp, const float4 *__restrict__ np, const float4 *__restrict__ fp, const int *__restrict__ cp) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) {
        put_title(out, "FRAMEBEG");
        for (int b = 0; b < 4; ++b) out[8 + b] = (unsigned char)((unsigned)L.nstep >> (8 * b));
        put_title(out + 12, "NATOM");
        const unsigned long long n = L.n_l + L.n_p;
        for (int b = 0; b < 8; ++b) out[20 + b] = (unsigned char)(n >> (8 * b));
        put_title(out + 28, "IDENTITY");
        if (L.off_x) put_title(out + L.off_x - 8, "POSITION");
        if (L.off_v) put_title(out + L.off_v - 8, "VELOCITY");
        if (L.off_n) put_title(out + L.off_n - 8, "ROTATION");
        if (L.off_aff) put_title(out + L.off_aff - 8, "VORONOI");
        if (L.off_f) put_title(out + L.off_f - 8, "FORCE");
        put_title(out + L.off_end, "FRAMEEND");
    }
    if (k >= L.n_l + L.n_p) return;
    const bool prot = k >= L.n_l;
    const size_t i = prot ? k - L.n_l : k;
    if (prot ? ((int)i < L.p0 || (int)i >= L.p1) : ((int)i < L.l0 || (int)i >= L.l1)) return;
    const float4 x = prot ? xp[i] : xl[i], n = prot ? np[i] : nl[i];
    int *id = reinterpret_cast<int *>(out + L.off_id) + 2 * k;
    id[0] = prot ? __float_as_int(n.w) : L.tag_base + (int)i;      // container.h:122-130: lipid tag = base + i, type = 0
    id[1] = prot ? __float_as_int(x.w) : 0;
    if (L.off_x) put3(out, L.off_x, k, x);
    if (L.off_v) put3(out, L.off_v, k, prot ? vp[i] : vl[i]);
    if (L.off_n) put3(out, L.off_n, k, n);
    if (L.off_aff) reinterpret_cast<int *>(out + L.off_aff)[k] = prot ? cp[i] : cl[i];
    if (L.off_f) put3(out, L.off_f, k, prot ? fp[i] : fl[i]);
}

int ensure_stage(orbc_ctx *c, size_t floats) {
    if (c->stage_cap < floats) { ORBC_TRY(dev_alloc(&c->stage, floats)); c->stage_cap = floats; }
    return ORBC_OK;
}

int alloc_species(orbc_ctx *c, Species &s, size_t n) {
    if (n + 64 > s.cap) {                                        // the pair kernels read up to 3 elements past a cell's last slot: keep 64 spare
        const size_t cap = n + 64 + n / 16;
        for (int b = 0; b < 2; ++b) {
            ORBC_TRY(dev_alloc(&s.x[b], cap)); ORBC_TRY(dev_alloc(&s.nn[b], cap)); ORBC_TRY(dev_alloc(&s.v[b], cap)); ORBC_TRY(dev_alloc(&s.o[b], cap));
            ORBC_TRY(dev_alloc(&s.cellid[b], cap));
        }
        ORBC_TRY(dev_alloc(&s.f, cap)); ORBC_TRY(dev_alloc(&s.t, cap));
        ORBC_TRY(dev_alloc(&s.aff, cap + 1)); ORBC_TRY(dev_alloc(&s.li, cap + 1));
        ORBC_TRY(dev_alloc(&s.cells, cap)); ORBC_TRY(dev_alloc(&s.cells_tmp, cap));
        s.cap = cap;
    }
    s.n = n; s.cur = 0; s.cur_xn = 0; s.has_partition = false;
    c->nl_valid = false;
    const int sp = (int)(&s - c->sp);
    ORBC_LAUNCH(c, k_set_range_const, 1, 1, 0, c->d_range + 2 * sp, 0, (int)n);   // a single GPU computes every slot
    return ORBC_OK;
}

void free_species(Species &s) {
    for (int b = 0; b < 2; ++b) { dev_free(s.x[b]); dev_free(s.nn[b]); dev_free(s.v[b]); dev_free(s.o[b]); dev_free(s.cellid[b]); }
    dev_free(s.f); dev_free(s.t); dev_free(s.aff); dev_free(s.li); dev_free(s.cells); dev_free(s.cells_tmp); dev_free(s.cell_start);
    s = Species();
}

GridDev grid_dev(const orbc_ctx *c) {
    GridDev g;
    g.lox = c->grid.lo[0]; g.loy = c->grid.lo[1]; g.loz = c->grid.lo[2];
    g.h = c->grid.h; g.inv_h = 1.0f / c->grid.h;
    g.dx = c->grid.dim[0]; g.dy = c->grid.dim[1]; g.dz = c->grid.dim[2];
    g.bin_start = c->grid.bin_start; g.bin_items = c->grid.bin_items; g.sorted = c->grid.sorted;
    return g;
}

// device-side consistency flags -> status (synchronises)
int check_flags(orbc_ctx *c) {
    ORBC_CUDA(cudaMemcpyAsync(c->h_flags, c->d_flags, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ORBC_CUDA(cudaStreamSynchronize(c->stream));
    if (c->h_flags[0]) return fail(ORBC_ERR_STATE, "centroid stencil of cell %d holds more than %d cells (or more than 32 closer than 6)", c->h_flags[0] - 1, kStencilStride);
    if (c->h_flags[1]) return fail(ORBC_ERR_STATE, "particle %d has no nearest centroid (NaN position?)", c->h_flags[1] - 1);
    if (c->h_flags[2]) return fail(ORBC_ERR_STATE, "protein %d carries a tag outside the tag->index map", c->h_flags[2] - 1);
    if (c->h_flags[3] > 0) return fail(ORBC_ERR_STATE, "decomposed run: rank %d never reached a barrier (rank %d waited 20 s)", c->h_flags[3] - 1, c->mg.rank);
    if (c->h_flags[3] < 0) return fail(ORBC_ERR_STATE, "decomposed run: rank %d now owns %d particles of one container, more than its launch bound", c->mg.rank, -c->h_flags[3]);
    return ORBC_OK;
}

// grid + stencils from the current centroids (the part of VoronoiDiagram::update that replaces tree.build, voronoi.h:83)
int build_index(orbc_ctx *c, bool all_cells = true, bool renumbered = true) {
    const int nc = c->n_cells;
    Grid &g = c->grid;
    ORBC_CUDA(cudaMemsetAsync(g.bin_start, 0, sizeof(int) * ((size_t)g.nbins + 1), c->stream));
    const GridDev gd = grid_dev(c);
    ORBC_LAUNCH(c, k_bin_count, blocks_for(nc, kBlock), kBlock, 0, c->centroid, nc, gd, g.bin_start, g.bin_of, g.bin_slot);
    ORBC_TRY(scan_exclusive(c, g.bin_start, g.nbins));
    ORBC_LAUNCH(c, k_bin_fill, blocks_for(nc, kBlock), kBlock, 0, nc, g.bin_start, g.bin_of, g.bin_slot, g.bin_items, c->centroid, g.sorted);
    // decomposed run: only the owned cells' stencils are read (pair kernels, and the nearest-centroid search starts from the
    // particle's previous cell, which its owner owns) — except right after a Morton renumbering of the cells
    HaloOut halo;
    halo.dest_mask = c->mg.dest_mask; halo.need = c->mg.need; halo.need_epoch = ++c->mg.need_epoch; halo.own = cell_owners(c); halo.rank = c->mg.rank;
    if (!c->mg.need) halo.own.world = 1;                         // before orbc_mg_export: no halo bookkeeping yet
    const bool part = mg_active(c) && !all_cells;
    const int c0 = part ? c->mg.cb : 0, c1 = part ? c->mg.ce : nc;
    const WideOut wide = {c->wide, c->wide_cnt, c->cen_ref};
    if (c1 > c0) {
        if (c->wide_valid && !renumbered && c->wide_on) {
            // the cells kept their numbers since the last full search: re-classify the recorded neighbours (k_stencil_refresh); the full
            // search stands by behind a device flag in case a centroid has outrun the margin
            const Movers mv = {c->movers, c->wide_ok + 1, kMoversCap}, patch = {c->movers + kMoversCap, c->wide_ok + 2, kMoversCap};
            ORBC_CUDA(cudaMemsetAsync(c->wide_ok, 0xff, sizeof(int), c->stream));
            ORBC_CUDA(cudaMemsetAsync(c->wide_ok + 1, 0, 2 * sizeof(int), c->stream));
            ORBC_LAUNCH(c, k_centroid_disp, blocks_for(nc, kBlock), kBlock, 0, c->centroid, c->cen_ref, nc, c->wide_ok, mv);
            ORBC_LAUNCH(c, k_stencil_refresh, blocks_for(c1 - c0, kStencilWarps), kStencilWarps * 32, 0, c->centroid, c0, c1, c->wide, c->wide_cnt, c->stencil, c->stencil_cnt, c->d_flags, halo, c->wide_ok);
            ORBC_LAUNCH(c, k_stencil_movers<true>, 148, kStencilWarps * 32, 0, c->centroid, c->cen_ref, c0, c1, gd, c->stencil, c->stencil_cnt, c->d_flags, halo, c->wide, c->wide_cnt, mv, patch, c->wide_ok, c->d_counters);
            ORBC_LAUNCH(c, k_stencil_movers<false>, 148, kStencilWarps * 32, 0, c->centroid, c->cen_ref, c0, c1, gd, c->stencil, c->stencil_cnt, c->d_flags, halo, c->wide, c->wide_cnt, patch, patch, c->wide_ok, c->d_counters);
            ORBC_LAUNCH(c, k_stencil_build, 148 * 4, kStencilWarps * 32, 0, c->centroid, nc, c0, c1, gd, c->stencil, c->stencil_cnt, c->d_flags, halo, wide, c->wide_ok, c->d_counters);
        } else {
            ORBC_LAUNCH(c, k_stencil_build, blocks_for(c1 - c0, kStencilWarps), kStencilWarps * 32, 0, c->centroid, nc, c0, c1, gd, c->stencil, c->stencil_cnt, c->d_flags, halo, wide, (const int *)nullptr, c->d_counters);
            c->wide_valid = true;                                // (a rank's partial search records its own cells: all its refreshes need)
        }
    }
    c->stencil_valid = true;
    c->lruns_valid = false;
    return ORBC_OK;
}

// fit the uniform grid to the centroids' bounding box (host side, at upload); centroids that later drift outside are
// clamped into the boundary bins, which keeps every search exact
int fit_grid_box(orbc_ctx *c, float lo[3], float hi[3]);
int fit_grid(orbc_ctx *c, const float *centroids3, int nc) {
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = 0; i < nc; ++i) for (int d = 0; d < 3; ++d) {
        const float v = centroids3[3 * i + d];
        if (v == v) { lo[d] = std::min(lo[d], v); hi[d] = std::max(hi[d], v); }
    }
    return fit_grid_box(c, lo, hi);
}
int fit_grid_box(orbc_ctx *c, float lo[3], float hi[3]) {
    for (int d = 0; d < 3; ++d) if (lo[d] > hi[d]) { lo[d] = 0; hi[d] = 0; }
    Grid &g = c->grid;
    float h = kBin;
    const float margin = 4 * kBin;
    for (;;) {
        size_t nb = 1;
        for (int d = 0; d < 3; ++d) {
            g.lo[d] = lo[d] - margin;
            g.dim[d] = std::max(1, (int)std::ceil((hi[d] - lo[d] + 2 * margin) / h));
            nb *= (size_t)g.dim[d];
        }
        if (nb <= ((size_t)12 << 20)) { g.nbins = (int)nb; break; }
        h *= 1.25f;
    }
    g.h = h;
    if (g.cap_bins < (size_t)g.nbins + 1) { ORBC_TRY(dev_alloc(&g.bin_start, (size_t)g.nbins + 1)); g.cap_bins = (size_t)g.nbins + 1; }
    return ORBC_OK;
}

// forcefield_canonical.h:23-28
float pow_chain(float base, int expo) { return expo != 0 ? base * pow_chain(base, expo - 1) : 1.0f; }
float ff_rep(float cut, float req, float eps) { return eps / pow_chain(cut - req, 8); }
float ff_att(float cut, float req, float eps) { return (float)(-2.0 * eps / pow_chain(cut - req, 4)); }

// hit lists go with the default kernels; on a decomposed run the ranks exchange their displacement bounds (k_nl_share)
bool nl_active(const orbc_ctx *c) {
    if (!c->nl_on || c->pair_impl != 2 || c->ll_variant != 1) return false;
    if (!mg_active(c)) return true;
    // decomposed: measured on the RBC (profiles/r02_configs/r02_lists_*), the lists gain 3 % with 2 ranks and 1.4 % with 4; with 8 they lost
    // 3.5 % before the gate chose its skin (half of the recordings were then never walked) and were not measured again: a rank's share of the
    // work shrinks, the gate, the exchange of the bounds and the launches that return at once do not.  Automatic = up to kNlAutoWorld ranks
    return c->mg.connected && (c->nl_on == 2 || c->mg.world <= kNlAutoWorld);
}
int nl_share(orbc_ctx *c);
int prot_lanes(const orbc_ctx *c, size_t np) { return c->prot_lanes ? c->prot_lanes : (np <= 400000 ? 4 : 1); }
int nl_ensure(orbc_ctx *c) {
    Species &L = c->sp[0], &P = c->sp[1];
    if (!c->nl_state) { NlState *st = nullptr; ORBC_TRY(dev_alloc(&st, 1)); c->nl_state = st; ORBC_CUDA(cudaMemsetAsync(st, 0, sizeof(NlState), c->stream)); c->nl_valid = false; }
    if (c->ll_list_lipids < L.cap) {
        const size_t groups = (L.cap + 63) / 64;
        ORBC_TRY(dev_alloc(&c->ll_list, groups * 64 * (size_t)c->nl_cap_ll)); ORBC_TRY(dev_alloc(&c->ll_cnt, groups * 64));
        c->ll_list_lipids = L.cap; c->nl_valid = false;
    }
    const size_t prot_rows = (mg_active(c) ? owned_bound(c, ORBC_PROTEIN) : P.cap) * 4 + 64;   // one row per thread of the protein kernel: up to 4 lanes per protein
    if (P.n && c->pl_list_proteins < prot_rows) {
        const size_t groups = (prot_rows + 63) / 64;
        ORBC_TRY(dev_alloc(&c->pl_list, groups * 64 * (size_t)c->nl_cap_pl)); ORBC_TRY(dev_alloc(&c->pl_cnt, groups * 64));
        ORBC_TRY(dev_alloc(&c->pp_list, groups * 64 * (size_t)c->nl_cap_pp)); ORBC_TRY(dev_alloc(&c->pp_cnt, groups * 64));
        c->pl_list_proteins = prot_rows; c->nl_valid = false;
    }
    return ORBC_OK;
}

void fill_integ(IntegArgs &a, orbc_ctx *c, int sp, const orbc_step_params *p) {
    Species &s = c->sp[sp];
    a.x = s.X(); a.v = s.V(); a.f = s.f; a.nn = s.N(); a.o = s.O(); a.t = s.t;
    a.x_out = a.x; a.nn_out = a.nn;
    a.n = s.n; a.species = sp;
    a.dt = (float)p->dt; a.dt_d = p->dt;
    for (int i = 0; i < kNType; ++i) a.gamma[i] = a.sigma[i] = 0.f;
    a.zeta = p->zeta;
    a.dlo = p->box_lo; a.dhi = p->box_hi; a.lo = (float)p->box_lo; a.hi = (float)p->box_hi;
    a.dr_opt = p->dr_opt; a.dn_opt = p->dn_opt;
    a.seed = p->seed; a.step = (uint32_t)p->nstep;
    a.noise = nullptr; a.acc = c->d_acc; a.zeta_dev = nullptr;
    a.disp = nl_active(c) && c->nl_state ? ((NlState *)c->nl_state)->disp : nullptr;
    a.range = c->d_range + 2 * sp;
    a.clear = 1;
    a.push.world = 1; a.push.cell_mask = nullptr; a.push.pmask = nullptr; a.push.cellid = nullptr;
    if (mg_active(c)) {
        // decomposed: new x, n go to the OTHER buffer, here and on the ranks that read them as halo; peers may still be reading
        // the current one (their forces of this step), and nobody reads the other one until the barrier after this push
        a.push.world = c->mg.world; a.push.cell_mask = c->mg.dest_mask; a.push.pmask = sp == ORBC_PROTEIN ? c->mg.pmask : nullptr; a.push.cellid = s.C();
        const int nb = s.cur_xn ^ 1;
        a.x_out = s.x[nb]; a.nn_out = s.nn[nb];
        for (int r = 0; r < kMaxWorld; ++r) { a.push.x[r] = c->mg.peers.x[sp][nb][r]; a.push.nn[r] = c->mg.peers.nn[sp][nb][r]; }
    }
}


// integrate_langevin.h:110-114: gamma = 6 pi eta R (fp64 -> fp32), sigma = sqrtf(2 kBT gamma) * sqrt(3 / dt)
void langevin_coeffs(const orbc_ctx *c, IntegArgs &a, const orbc_step_params *p) {
    for (int i = 0; i < kNType; ++i) {
        a.gamma[i] = (float)(6.0 * M_PI * p->eta * c->host_ff.radius[i]);
        a.sigma[i] = (float)(std::sqrt((float)(2 * p->kBT * a.gamma[i])) * std::sqrt(3.0 / p->dt));
    }
}

// largest interaction range of every protein type against lipids and against the protein types present
CullTable cull_table(const orbc_ctx *c) {
    CullTable ct;
    for (int t = 0; t < kNType; ++t) {
        ct.cut_l[t] = std::sqrt(std::max(c->host_ff.cutsqlp[t], c->host_ff.lj_cutsq[t]));
        float m = 0.f;
        for (int u = 0; u < kNType; ++u) if (c->type_mask >> u & 1) m = std::max(m, std::max(c->host_ff.cutsqpp[t + kNType * u], c->host_ff.lj_cutsq[t + kNType * u]));
        ct.cut_p[t] = std::sqrt(m); ct.cutsq_p[t] = m;
    }
    return ct;
}

// thread -> protein map of k_pair_prot (heavy types first); rebuilt whenever the protein storage order changes
int build_porder(orbc_ctx *c) {
    Species &P = c->sp[1];
    if (!P.n) return ORBC_OK;
    const size_t cap = owned_bound(c, ORBC_PROTEIN);
    if (c->porder_cap < cap + 1) { ORBC_TRY(dev_alloc(&c->porder, cap + 1)); c->porder_cap = cap + 1; }
    const CullTable ct = cull_table(c);
    float lo = 1e30f, hi = 0.f;                                   // heavy = upper half of the ranges present
    for (int t = 0; t < kNType; ++t) if (c->type_mask >> t & 1) { lo = std::min(lo, ct.cut_l[t]); hi = std::max(hi, ct.cut_l[t]); }
    const float heavy_cut = hi > 1.5f * lo ? 0.5f * (lo + hi) : 2.f * hi + 1.f;   // homogeneous ranges: nobody is heavy
    int *flag = P.li;                                             // free between cell updates (n + 1 ints)
    ORBC_LAUNCH(c, k_porder_flag, blocks_for(cap, kBlock), kBlock, 0, P.X(), c->d_range, cap, ct, heavy_cut, flag);
    ORBC_TRY(scan_exclusive(c, flag, (int)cap));
    ORBC_LAUNCH(c, k_porder_scatter, blocks_for(cap, kBlock), kBlock, 0, flag, c->d_range, cap, P.X(), ct, heavy_cut, c->porder);
    c->porder_valid = true;
    return ORBC_OK;
}

int launch_pairwise(orbc_ctx *c, bool accumulate = true) {
    Species &L = c->sp[0], &P = c->sp[1];
    if (!c->n_cells || !L.has_partition) return fail(ORBC_ERR_ARG, "compute_pairwise_fused: no Voronoi partition (call orbc_voronoi_upload / orbc_rebuild first)");
    if (P.n && !P.has_partition) return fail(ORBC_ERR_ARG, "compute_pairwise_fused: proteins are not partitioned");
    PairArgs a;
    a.xl = L.X(); a.nl = L.N(); a.cs_l = L.cell_start; a.cell_l = L.C(); a.n_l = (int)L.n;
    a.xp = P.X(); a.np = P.N(); a.cs_p = P.cell_start; a.cell_p = P.C(); a.n_p = (int)P.n;
    a.stencil = c->stencil; a.stencil_cnt = c->stencil_cnt;
    a.fl = L.f; a.tl = L.t; a.fp = P.f; a.tp = P.t;
    a.range = c->d_range;
    const bool mg = mg_active(c);
    a.cb = mg ? c->mg.cb : 0; a.ce = mg ? c->mg.ce : c->n_cells; a.world = mg ? c->mg.world : 1;
    a.dest_mask = c->mg.dest_mask;
    a.accumulate = accumulate ? 1 : 0;
    a.counters = c->d_counters;
    const size_t nl_count = owned_bound(c, ORBC_LIPID), np = owned_bound(c, ORBC_PROTEIN);
    constexpr unsigned kSmallGrid = 148 * 16;                   // a gated launch that usually returns at once: grid-stride over few blocks
    if (c->pair_impl == 1) {
        if (mg) return fail(ORBC_ERR_ARG, "pair_impl 1 is a single-GPU cross-check");
        if (L.n) { ProfScope ps(c, ORBC_PROF_PAIR_LIPID); ORBC_LAUNCH(c, k_pair_lipid, blocks_for(nl_count, 128), 128, 0, a); }
        if (P.n) { ProfScope ps(c, ORBC_PROF_PAIR_PROTEIN); ORBC_LAUNCH(c, k_pair_protein, blocks_for(np, 128), 128, 0, a); }
        return ORBC_OK;
    }
    // hit lists: the evaluation after a rebuild records them, the following ones walk them (pair_queue.cuh)
    const bool nl = nl_active(c);
    bool host_build = true;
    NlState *nls = nullptr;
    if (nl) {
        ORBC_TRY(nl_ensure(c));
        nls = (NlState *)c->nl_state;
        host_build = !c->nl_valid;
        if (mg && c->nl_moves > 1) host_build = true;             // (the ranks exchange the bound of ONE step)
        if (c->nl_debug_mode > 0) host_build = true;
        ORBC_LAUNCH(c, k_nl_gate, 1, 32, 0, nls, host_build ? 1 : 0, c->nl_debug_mode, c->nl_moves, c->nl_skin, c->nl_skin_max,
                    mg ? (const unsigned *)(c->mg.flags + kMaxWorld * (1 + c->mg.disp_par)) : (const unsigned *)nullptr, mg ? c->mg.world : 1);
        c->nl_moves = 0; c->nl_valid = true;
    }
    {
        ProfScope ps(c, ORBC_PROF_PAIR_LIPID);
        // bounding spheres of the cells' current members: the protein kernel culls with them (on a list-walking step only the gated rebuild of the lists would)
        ORBC_LAUNCH(c, k_cell_bounds, blocks_for(c->n_cells, 128), 128, 0, c->centroid, c->n_cells, L.cell_start, L.X(), P.n ? P.cell_start : nullptr, P.X(), c->lbound, c->pbound,
                    mg ? c->mg.need : (const int *)nullptr, c->mg.need_epoch, nl ? &nls->need : (const int *)nullptr);
        if (L.n && a.ce > a.cb) {
            // candidate runs merged over Morton-adjacent stencil cells, rebuilt after every rebuild of the partition
            if (!c->lruns_valid) {
                ORBC_CUDA(cudaMemsetAsync(c->tile_overflow, 0, sizeof(int), c->stream));
                ORBC_LAUNCH(c, k_lipid_runs, blocks_for(a.ce - a.cb, 128), 128, 0, a.cb, a.ce, c->stencil, c->stencil_cnt, L.cell_start, c->lruns, c->lrun_cnt, c->tile_cap, c->tile_overflow);
                c->lruns_valid = true;
            }
            const orbc_forcefield &ff = c->host_ff;
            const LLConst kc = {ff.cutll, 8.0f * ff.repll, 4.0f * ff.attll, ff.alphall, ff.alphall * ff.attll, 1.0f - ff.alphall, ff.cutsqll};
            if (c->ll_variant == 0) {
                // warp-per-cell tile kernel; the thread-per-lipid kernel takes the step when a cell does not fit the tile (device flag)
                const unsigned warps = blocks_for((size_t)(a.ce - a.cb), kTileCells);
                ORBC_LAUNCH(c, k_pair_ll_t, blocks_for(warps, kTileWarps), kTileWarps * 32, kTileWarps * kTileBytes, a, kc, c->lruns, c->lrun_cnt, c->tile_overflow);
                ORBC_LAUNCH(c, (k_pair_ll_r<20, 4, false>), kSmallGrid, kLLBlock, 0, a, kc, c->lruns, c->lrun_cnt, c->tile_overflow, 1, LLList{}, (unsigned *)nullptr);
            } else if (!nl) {
                ORBC_LAUNCH(c, (k_pair_ll_r<20, 4, false>), blocks_for(nl_count, kLLBlock), kLLBlock, 0, a, kc, c->lruns, c->lrun_cnt, (const int *)nullptr, 0, LLList{}, (unsigned *)nullptr);
            } else {
                // hit lists: the device decides (k_nl_gate) whether this evaluation walks the lists, searches and records them, or just
                // searches; the candidates are launched over a grid that fills the GPU once (their blocks draw 64-lipid groups from a
                // ticket counter), all but one return at once
                const LLList ll = {c->ll_list, c->ll_cnt, c->nl_cap_ll, nls};
                const unsigned groups = blocks_for(nl_count, kLLBlock);
                const unsigned g16 = std::min(groups, 148u * 16u), g20 = std::min(groups, 148u * 20u);   // (blocks per SM: the kernels' launch bounds)
                // (what the host knows: after a change of the partition the gate never orders a walk, and otherwise never a recording)
                if (!host_build) {
                    ORBC_LAUNCH(c, k_pair_ll_list<16>, g16, kLLBlock, 0, a, kc, &nls->need, 0, ll, nls->work + 0);
                } else ORBC_LAUNCH(c, (k_pair_ll_r<16, 4, true>), g16, kLLBlock, 0, a, kc, c->lruns, c->lrun_cnt, &nls->need, 1, ll, nls->work + 1);
                ORBC_LAUNCH(c, (k_pair_ll_r<20, 4, false>), g20, kLLBlock, 0, a, kc, c->lruns, c->lrun_cnt, &nls->need, 2, LLList{}, nls->work + 2);
            }
        }
        // (decomposed: the lipid side of the protein-lipid pairs whose protein lives on another rank is the epilogue of the lipid kernels)
    }
    if (P.n) {
        const CullTable ct = cull_table(c);
        if (!c->porder_valid) ORBC_TRY(build_porder(c));
        ProfScope ps(c, ORBC_PROF_PAIR_PROTEIN);
        // few owned proteins (a rank of a decomposed run): more lanes per protein, shorter dependent-load chains, more warps
        const PLists pls = {c->pl_list, c->pl_cnt, c->pp_list, c->pp_cnt, c->nl_cap_pl, c->nl_cap_pp, nls};
        const int lanes = prot_lanes(c, np);
        if (nl) {
            const unsigned pieces = blocks_for(np * lanes, kPBlock);
#define ORBC_PROT3(LPP) do { \
                if (!host_build) ORBC_LAUNCH(c, k_pair_prot_list<LPP>, pieces, kPBlock, 0, a, c->porder, &nls->need, 0, pls); \
                else ORBC_LAUNCH(c, (k_pair_prot<LPP, true>), pieces, kPBlock, 0, a, c->lbound, c->pbound, ct, c->porder, &nls->need, 1, pls); \
                ORBC_LAUNCH(c, (k_pair_prot<LPP, false>), pieces, kPBlock, 0, a, c->lbound, c->pbound, ct, c->porder, &nls->need, 2, pls); } while (0)
            if (lanes == 4) ORBC_PROT3(4); else if (lanes == 2) ORBC_PROT3(2); else ORBC_PROT3(1);
#undef ORBC_PROT3
        } else {
            const unsigned grid = blocks_for(np * lanes, kPBlock);
            if (lanes == 4) ORBC_LAUNCH(c, (k_pair_prot<4, false>), grid, kPBlock, 0, a, c->lbound, c->pbound, ct, c->porder, (const int *)nullptr, 0, pls);
            else if (lanes == 2) ORBC_LAUNCH(c, (k_pair_prot<2, false>), grid, kPBlock, 0, a, c->lbound, c->pbound, ct, c->porder, (const int *)nullptr, 0, pls);
            else ORBC_LAUNCH(c, (k_pair_prot<1, false>), grid, kPBlock, 0, a, c->lbound, c->pbound, ct, c->porder, (const int *)nullptr, 0, pls);
        }
    }
    return ORBC_OK;
}

int launch_bonded(orbc_ctx *c) {
    Species &P = c->sp[1];
    if (!c->n_bonds) return ORBC_OK;
    ProfScope ps(c, ORBC_PROF_BONDED);
    if (mg_active(c)) ORBC_LAUNCH(c, k_bonded, blocks_for(c->mg.my_bonds_cap, kBlock), kBlock, 0, c->bonds, c->n_bonds, c->tag2idx, P.X(), P.f, c->d_range, c->mg.my_bonds);
    else ORBC_LAUNCH(c, k_bonded, blocks_for(c->n_bonds, kBlock), kBlock, 0, c->bonds, c->n_bonds, c->tag2idx, P.X(), P.f, c->d_range, (const int *)nullptr);
    return ORBC_OK;
}

int build_tag2idx(orbc_ctx *c) {
    Species &P = c->sp[1];
    if (!P.n || !c->tag2idx) return ORBC_OK;
    ORBC_LAUNCH(c, k_build_tag2idx, blocks_for(P.n, kBlock), kBlock, 0, P.N(), P.n, c->tag2idx, c->tag2idx_size, c->d_flags);
    return ORBC_OK;
}

int do_voronoi_update(orbc_ctx *c, int nstep, int freq_sort_ctrd) {
    Species &L = c->sp[0];
    if (!c->n_cells || !L.has_partition) return fail(ORBC_ERR_ARG, "voronoi_update: no previous partition (orbc_voronoi_upload with cell_start first)");
    const int nc = c->n_cells;
    const bool mg = mg_active(c);
    c->nl_valid = false;                                         // cells are about to be renumbered / repartitioned: the hit lists die with the old slots
    // centroids of the owned cells, published to every rank (voronoi.h:123-140)
    CentroidOut out; out.world = mg ? c->mg.world : 1;
    for (int r = 0; r < kMaxWorld; ++r) out.dst[r] = mg ? c->mg.peers.centroid[c->mg.cen_par ^ 1][r] : c->centroid_tmp;
    const int cb = mg ? c->mg.cb : 0, ce = mg ? c->mg.ce : nc;
    if (ce > cb) ORBC_LAUNCH(c, k_centroid_update, blocks_for(ce - cb, kBlock), kBlock, 0, L.cell_start, L.X(), cb, ce, out, (const int *)nullptr);
    ORBC_TRY(mg_barrier(c));
    const bool morton = freq_sort_ctrd > 0 && nstep % freq_sort_ctrd == 0;
    if (morton) {
        ORBC_LAUNCH(c, k_morton_keys, blocks_for(nc, kBlock), kBlock, 0, c->centroid_tmp, nc, c->keys, c->perm);
        ORBC_TRY(radix_sort_pairs(c, c->keys, c->perm, c->keys_tmp, c->perm_tmp, nc));
        ORBC_LAUNCH(c, k_permute_centroids, blocks_for(nc, kBlock), kBlock, 0, c->centroid_tmp, c->perm, nc, c->centroid, c->inv);
        // cells were renumbered: carry every particle's previous cell into the new numbering (it is only a search hint)
        for (int s = 0; s < 2; ++s) if (c->sp[s].n) ORBC_LAUNCH(c, k_remap_cellid, blocks_for(owned_bound(c, s), kBlock), kBlock, 0, c->sp[s].C(), c->d_range + 2 * s, c->inv);
    } else {
        std::swap(c->centroid, c->centroid_tmp);
        c->mg.cen_par ^= 1;
    }
    return build_index(c, morton, morton);
}

// VCellList::update in three phases, so that a decomposed rebuild can run both containers through each phase between two
// barriers.  `which`: the containers a phase works on (bit 0 lipids, bit 1 proteins); with both, every kernel of a phase is ONE
// launch whose leading blocks take the lipids and whose trailing blocks take the proteins (the protein launches are short and
// latency-bound -- 46 us of nearest-centroid search for a fifth of the particles -- and disappear behind the lipids' work).
// Phase A: nearest centroid of every owned particle + arrival counts (voronoi.h:179-216), counts published.
constexpr int kBothContainers = 3;
int cell_update_assign(orbc_ctx *c, int which, const int *keep = nullptr) {
    if (!c->n_cells || !c->stencil_valid) return fail(ORBC_ERR_ARG, "cell_update: no Voronoi diagram");
    const int nc = c->n_cells;
    const bool mg = mg_active(c);
    AssignArgs a[2]; unsigned blocks[2] = {0, 0}; int m = 0;
    ShareArgs sh[2]; int ms = 0;
    for (int sp = 0; sp < 2; ++sp) {
        if (!((which >> sp) & 1)) continue;
        Species &S = c->sp[sp];
        if (!S.cell_start) ORBC_TRY(dev_alloc(&S.cell_start, (size_t)nc + 1));
        int *cnt = mg ? c->mg.cnt_all[sp] + (size_t)c->mg.rank * (nc + 1) : S.cell_start;
        ORBC_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)nc + 1), c->stream));
        if (S.n) {
            a[m] = AssignArgs{S.X(), S.has_partition ? S.C() : nullptr, c->d_range + 2 * sp, S.aff, S.li, cnt, sp == ORBC_LIPID ? keep : nullptr};
            blocks[m++] = blocks_for(owned_bound(c, sp), kBlock);
        }
        if (mg) { sh[ms].cnt_me = cnt; sh[ms].prev = c->mg.cnt_prev[sp]; for (int r = 0; r < kMaxWorld; ++r) sh[ms].dst[r] = c->mg.peers.cnt_all[sp][r]; ++ms; }
    }
    const AssignCommon common = {c->centroid, nc, c->stencil, c->stencil_cnt, grid_dev(c), c->d_counters, c->d_flags};
    if (m == 2) ORBC_LAUNCH(c, k_assign_nearest2, blocks[0] + blocks[1], kBlock, 0, a[0], a[1], blocks[0], common);
    else if (m == 1) ORBC_LAUNCH(c, k_assign_nearest, blocks[0], kBlock, 0, a[0], common);
    if (ms) {
        const unsigned nb = blocks_for((size_t)nc + 1, kBlock);
        ORBC_LAUNCH(c, k_share_counts, nb * ms, kBlock, 0, sh[0], sh[ms - 1], nb, nc + 1, c->mg.rank, c->mg.world);   // (ms sets of nb blocks)
    }
    return ORBC_OK;
}
// Phase B: global cell_start (voronoi.h:217-227), then every particle moves to its new slot on its new owner (voronoi.h:228-231
// + reorder.h:73-149 as one scatter; the migration of a decomposed run is the same store, into a peer's memory)
int cell_update_move(orbc_ctx *c, int which) {
    const int nc = c->n_cells;
    const bool mg = mg_active(c);
    ScatterArgs sc[2]; MoveArgs mv[2]; unsigned blocks[2] = {0, 0}; int m = 0;
    TotalsArgs tt[2]; int mt = 0;
    for (int sp = 0; sp < 2; ++sp) {
        if (!((which >> sp) & 1)) continue;
        Species &S = c->sp[sp];
        if (mg) tt[mt++] = TotalsArgs{c->mg.cnt_all[sp], S.cell_start, c->mg.off_me[sp]};
    }
    if (mt) {
        const unsigned nb = blocks_for(nc, kBlock);
        ORBC_LAUNCH(c, k_cell_totals, nb * mt, kBlock, 0, tt[0], tt[mt - 1], nb, nc, c->mg.rank, c->mg.world);
    }
    if (which == kBothContainers) ORBC_TRY(scan_exclusive(c, c->sp[0].cell_start, nc, c->sp[1].cell_start));   // both cell_start arrays in one launch
    else ORBC_TRY(scan_exclusive(c, c->sp[which >> 1].cell_start, nc));
    for (int sp = 0; sp < 2; ++sp) {
        if (!((which >> sp) & 1)) continue;
        Species &S = c->sp[sp];
        const int *cnt_me = nullptr, *off_me = nullptr;
        if (mg) { cnt_me = c->mg.cnt_all[sp] + (size_t)c->mg.rank * (nc + 1); off_me = c->mg.off_me[sp]; }
        if (S.n) {
            const int nx = S.cur ^ 1, nxn = S.cur_xn ^ 1;
            sc[m] = ScatterArgs{S.aff, S.li, c->d_range + 2 * sp, S.cell_start, off_me, S.cells_tmp};
            MoveArgs &v = mv[m];
            v.aff = S.aff; v.range = c->d_range + 2 * sp; v.cell_start = S.cell_start; v.cnt_me = cnt_me; v.off_me = off_me; v.cells_tmp = S.cells_tmp; v.cells = S.cells;
            v.x0 = S.X(); v.n0 = S.N(); v.v0 = S.V(); v.o0 = S.O();
            v.announce_tags = (mg && sp == ORBC_PROTEIN) ? 1 : 0;
            v.d.own = cell_owners(c);
            for (int r = 0; r < kMaxWorld; ++r) {
                v.d.x[r] = mg ? c->mg.peers.x[sp][nxn][r] : S.x[nxn]; v.d.nn[r] = mg ? c->mg.peers.nn[sp][nxn][r] : S.nn[nxn];
                v.d.v[r] = mg ? c->mg.peers.v[sp][nx][r] : S.v[nx]; v.d.o[r] = mg ? c->mg.peers.o[sp][nx][r] : S.o[nx];
                v.d.cellid[r] = mg ? c->mg.peers.cellid[sp][nx][r] : S.cellid[nx];
                v.d.tag2idx[r] = mg ? c->mg.peers.tag2idx[r] : c->tag2idx;
            }
            blocks[m++] = blocks_for(owned_bound(c, sp), kBlock);
            S.cur = nx; S.cur_xn = nxn;
        }
        S.has_partition = true;
        if (sp == ORBC_LIPID) c->lruns_valid = false;
    }
    if (m == 2) {
        ORBC_LAUNCH(c, k_cell_scatter2, blocks[0] + blocks[1], kBlock, 0, sc[0], sc[1], blocks[0]);
        ORBC_LAUNCH(c, k_rank_and_move2, blocks[0] + blocks[1], kBlock, 0, mv[0], mv[1], blocks[0]);
    } else if (m == 1) {
        ORBC_LAUNCH(c, k_cell_scatter, blocks[0], kBlock, 0, sc[0]);
        ORBC_LAUNCH(c, k_rank_and_move, blocks[0], kBlock, 0, mv[0]);
    }
    c->nl_valid = false;
    return ORBC_OK;
}
// Phase C: tag -> index map (container.h:39-58); decomposed: new owned range, bonded-partner masks, halo copies of the moved particles
int cell_update_finish(orbc_ctx *c, int which) {
    const bool mg = mg_active(c);
    if (which & 2) c->porder_valid = false;
    if (!mg) { if (which & 2) ORBC_TRY(build_tag2idx(c)); return ORBC_OK; }
    RangeArgs rg[2]; HaloArgs ha[2]; unsigned blocks[2] = {0, 0}; int mr = 0, m = 0;
    for (int sp = 0; sp < 2; ++sp) {
        if (!((which >> sp) & 1)) continue;
        Species &S = c->sp[sp];
        rg[mr++] = RangeArgs{S.cell_start, c->d_range + 2 * sp, (int)c->mg.own_cap[sp]};
    }
    ORBC_LAUNCH(c, k_set_range, 1, 32, 0, rg[0], rg[mr - 1], mr, c->mg.cb, c->mg.ce, c->d_flags);
    Species &P = c->sp[1];
    if ((which & 2) && P.n && c->n_bonds) {
        ORBC_CUDA(cudaMemsetAsync(c->mg.pmask, 0, (P.n + 3) / 4 * 4, c->stream));
        ORBC_CUDA(cudaMemsetAsync(c->mg.my_bonds, 0, sizeof(int), c->stream));
        ORBC_LAUNCH(c, k_bond_mask, blocks_for(c->n_bonds, kBlock), kBlock, 0, c->bonds, c->n_bonds, c->tag2idx, c->d_range, P.cell_start, cell_owners(c), (unsigned *)c->mg.pmask,
                    c->mg.my_bonds, c->mg.my_bonds_cap, c->d_flags);
    }
    for (int sp = 0; sp < 2; ++sp) {
        if (!((which >> sp) & 1)) continue;
        Species &S = c->sp[sp];
        if (!S.n) continue;
        HaloArgs &h = ha[m];
        h.range2 = c->d_range + 2 * sp; h.cell_mask = c->mg.dest_mask; h.pmask = sp == ORBC_PROTEIN ? c->mg.pmask : (const unsigned char *)nullptr;
        h.cellid = S.C(); h.x = S.X(); h.nn = S.N();
        for (int r = 0; r < kMaxWorld; ++r) { h.d.x[r] = c->mg.peers.x[sp][S.cur_xn][r]; h.d.nn[r] = c->mg.peers.nn[sp][S.cur_xn][r]; }
        blocks[m++] = blocks_for(owned_bound(c, sp), kBlock);
    }
    if (m) ORBC_LAUNCH(c, k_halo_push, blocks[0] + (m == 2 ? blocks[1] : 0), kBlock, 0, ha[0], ha[m - 1], blocks[0]);
    return ORBC_OK;
}

int do_cell_update(orbc_ctx *c, int sp) {
    ORBC_TRY(cell_update_assign(c, 1 << sp)); ORBC_TRY(mg_barrier(c));
    ORBC_TRY(cell_update_move(c, 1 << sp));   ORBC_TRY(mg_barrier(c));
    ORBC_TRY(cell_update_finish(c, 1 << sp)); return mg_barrier(c);
}

// `rebuild_follows`: the caller rebuilds next; the first barrier of the rebuild then also covers the arrival of this push
int do_integrate_langevin(orbc_ctx *c, const orbc_step_params *p, bool rebuild_follows = false, bool clear = true) {
    // decomposed: no barrier before the push — it goes to the x, n buffers nobody is reading (fill_integ)
    ++c->nl_moves;                                               // one tracked integration step (hit lists: displacement bound)
    {
        ProfScope ps(c, ORBC_PROF_INTEGRATE);
        // both containers in one launch (leading blocks the lipids, trailing blocks the proteins)
        IntegArgs a[2]; unsigned blocks[2] = {0, 0}; int m = 0;
        for (int sp = 0; sp < 2; ++sp) {
            Species &S = c->sp[sp];
            if (!S.n) continue;
            fill_integ(a[m], c, sp, p); langevin_coeffs(c, a[m], p);
            a[m].clear = clear ? 1 : 0;
            const float *hn = sp == 0 ? p->noise_lipid : p->noise_protein;
            if (hn) {
                if (c->noise_cap[sp] < 3 * S.n) { ORBC_TRY(dev_alloc(&c->noise[sp], 3 * S.n)); c->noise_cap[sp] = 3 * S.n; }
                ORBC_CUDA(cudaMemcpyAsync(c->noise[sp], hn, sizeof(float) * 3 * S.n, cudaMemcpyHostToDevice, c->stream));
                a[m].noise = c->noise[sp];
            }
            blocks[m++] = blocks_for(owned_bound(c, sp), 256);
            if (mg_active(c)) S.cur_xn ^= 1;
        }
        if (m == 2) ORBC_LAUNCH(c, k_verlet_langevin2, blocks[0] + blocks[1], 256, 0, a[0], a[1], blocks[0]);
        else if (m == 1) ORBC_LAUNCH(c, k_verlet_langevin, blocks[0], 256, 0, a[0]);
        ORBC_TRY(nl_share(c));
    }
    return rebuild_follows ? ORBC_OK : mg_barrier(c);            // the pushed halo has landed everywhere
}

// CUDA loads kernels lazily, and loading one may wait for the device to go idle — a rank spinning in k_mg_barrier while its
// peer loads a kernel for the first time would never be released.  A decomposed context therefore loads every kernel up front.
// decomposed run: publish this rank's displacement bound of the step (the barrier behind the halo push covers it)
int nl_share(orbc_ctx *c) {
    if (!mg_active(c) || !nl_active(c) || !c->nl_state) return ORBC_OK;
    c->mg.disp_par ^= 1;
    NlShare d; for (int r = 0; r < kMaxWorld; ++r) d.dst[r] = c->mg.peers.flags[r] ? c->mg.peers.flags[r] + kMaxWorld * (1 + c->mg.disp_par) : nullptr;
    ORBC_LAUNCH(c, k_nl_share, 1, 32, 0, (NlState *)c->nl_state, c->mg.rank, c->mg.world, d);
    return ORBC_OK;
}

int preload_kernels() {
    cudaFuncAttributes fa;
#define ORBC_PRELOAD(k) ORBC_CUDA(cudaFuncGetAttributes(&fa, (const void *)(k)))
    ORBC_PRELOAD(k_assign_nearest); ORBC_PRELOAD(k_bin_count); ORBC_PRELOAD(k_bin_fill); ORBC_PRELOAD(k_bond_mask); ORBC_PRELOAD(k_bonded);
    ORBC_PRELOAD(k_bounce_back); ORBC_PRELOAD(k_build_tag2idx); ORBC_PRELOAD(k_cell_bounds); ORBC_PRELOAD(k_cell_scatter); ORBC_PRELOAD(k_cell_totals);
    ORBC_PRELOAD(k_centroid_update); ORBC_PRELOAD(k_check_ids); ORBC_PRELOAD(k_check_bonds); ORBC_PRELOAD(k_clear_force); ORBC_PRELOAD(k_compact); ORBC_PRELOAD(k_count_strays); ORBC_PRELOAD(k_cv_apply); ORBC_PRELOAD(k_cv_center); ORBC_PRELOAD(k_cv_share); ORBC_PRELOAD(k_sum_partials); ORBC_PRELOAD(k_opt_fused); ORBC_PRELOAD(k_frame_pack);
    ORBC_PRELOAD(k_cv_normal_volume); ORBC_PRELOAD(k_fill_cellid); ORBC_PRELOAD(k_fill_int); ORBC_PRELOAD(k_halo_push); ORBC_PRELOAD(k_kinetic);
    ORBC_PRELOAD(k_mg_barrier); ORBC_PRELOAD(k_morton_keys); ORBC_PRELOAD(k_morton_keys_only); ORBC_PRELOAD(k_nh_final); ORBC_PRELOAD(k_nh_final_fused);
    ORBC_PRELOAD(k_nh_initial_fused); ORBC_PRELOAD(k_nh_zeta_update); ORBC_PRELOAD(k_share_ke); ORBC_PRELOAD(k_sum_ke); ORBC_PRELOAD(k_noise); ORBC_PRELOAD(k_opt_move); ORBC_PRELOAD(k_pack4);
    ORBC_PRELOAD(k_pair_lipid); ORBC_PRELOAD(k_lipid_runs); ORBC_PRELOAD(k_rank_only); ORBC_PRELOAD(k_init_centroids); ORBC_PRELOAD(k_bbox); ORBC_PRELOAD((k_pair_ll_r<20, 4, false>)); ORBC_PRELOAD(k_pair_ll_t); ORBC_PRELOAD((k_pair_prot<1, false>)); ORBC_PRELOAD((k_pair_prot<2, false>)); ORBC_PRELOAD((k_pair_prot<4, false>)); ORBC_PRELOAD((k_pair_prot<1, true>)); ORBC_PRELOAD((k_pair_prot<2, true>)); ORBC_PRELOAD((k_pair_prot<4, true>)); ORBC_PRELOAD(k_pair_prot_list<1>); ORBC_PRELOAD(k_pair_prot_list<2>); ORBC_PRELOAD(k_pair_prot_list<4>); ORBC_PRELOAD((k_pair_ll_r<16, 4, true>)); ORBC_PRELOAD(k_pair_ll_list<16>); ORBC_PRELOAD(k_nl_gate); ORBC_PRELOAD(k_nl_share); ORBC_PRELOAD(k_pair_protein);
    ORBC_PRELOAD(k_permute_centroids); ORBC_PRELOAD(k_porder_flag); ORBC_PRELOAD(k_porder_scatter); ORBC_PRELOAD(k_post_torque); ORBC_PRELOAD(k_radix_hist);
    ORBC_PRELOAD(k_radix_scatter); ORBC_PRELOAD(k_rank_and_move); ORBC_PRELOAD(k_remap_cellid); ORBC_PRELOAD(k_scan_onepass); ORBC_PRELOAD(k_set3); ORBC_PRELOAD(k_set_range); ORBC_PRELOAD(k_set_range_const); ORBC_PRELOAD(k_share_counts);
    ORBC_PRELOAD(k_stencil_build); ORBC_PRELOAD(k_stencil_refresh); ORBC_PRELOAD(k_stencil_movers<true>); ORBC_PRELOAD(k_stencil_movers<false>); ORBC_PRELOAD(k_centroid_disp); ORBC_PRELOAD(k_stray_mask); ORBC_PRELOAD(k_unpack3); ORBC_PRELOAD(k_unpack_w); ORBC_PRELOAD(k_verlet_langevin); ORBC_PRELOAD(k_verlet_langevin2); ORBC_PRELOAD(k_nh_initial_fused2); ORBC_PRELOAD(k_nh_final_fused2); ORBC_PRELOAD(k_assign_nearest2); ORBC_PRELOAD(k_cell_scatter2); ORBC_PRELOAD(k_rank_and_move2); ORBC_PRELOAD(k_zero4);
#undef ORBC_PRELOAD
    return ORBC_OK;
}

// constrain_volume.h:26-83.  Decomposed: every rank holds all centroids (same centre everywhere), computes the normals and the
// volume share of its own cells, the shares are exchanged (peer stores + one barrier) and summed in rank order.
int do_constrain_volume(orbc_ctx *c, float target, float strength) {
    Species &L = c->sp[0], &P = c->sp[1];
    if (!L.has_partition) return fail(ORBC_ERR_ARG, "constrain_volume: lipids are not partitioned");
    const int nc = c->n_cells;
    const bool mg = mg_active(c);
    const int cb = mg ? c->mg.cb : 0, ce = mg ? c->mg.ce : nc;
    ORBC_CUDA(cudaMemsetAsync(c->d_acc + 4, 0, sizeof(double), c->stream));
    ORBC_LAUNCH(c, k_cv_center, 1, 1024, 0, c->centroid, nc, c->d_acc);
    if (ce > cb) ORBC_LAUNCH(c, k_cv_normal_volume, blocks_for(ce - cb, 256), 256, 0, c->centroid, nc, cb, ce, L.cell_start, L.N(), c->cell_normal, c->d_acc);
    const double *vol_all = nullptr; const int *ptype = nullptr;
    if (mg) {
        // two sets of slots used alternately: a fast rank may publish its next share while a slow one still reads this one
        const int half = (c->mg.cv_par ^= 1) * kMaxWorld;
        CvShare d; for (int r = 0; r < kMaxWorld; ++r) { d.vol[r] = c->mg.peers.vol_all[r] + half; d.ptype[r] = c->mg.peers.cv_ptype[r]; }
        const size_t work = std::max<size_t>(kMaxWorld, std::min<size_t>(owned_bound(c, ORBC_PROTEIN), (size_t)nc));
        ORBC_LAUNCH(c, k_cv_share, blocks_for(work, kBlock), kBlock, 0, c->d_acc, P.X(), c->d_range, nc, c->mg.rank, c->mg.world, d);
        ORBC_TRY(mg_barrier(c));
        vol_all = c->mg.vol_all + half; ptype = c->mg.cv_ptype;
        ORBC_LAUNCH(c, k_sum_partials, 1, 32, 0, c->d_acc + 4, vol_all, c->mg.world);   // acc[4] = the whole volume, for the caller
    }
    const int world = mg ? c->mg.world : 1;
    if (L.n) ORBC_LAUNCH(c, k_cv_apply, blocks_for(owned_bound(c, ORBC_LIPID), kBlock), kBlock, 0, L.C(), c->d_range, c->cell_normal, (const float4 *)nullptr, (size_t)0,
                         (const int *)nullptr, 0, target, strength, c->d_acc, vol_all, world, L.f);
    if (P.n && P.has_partition) ORBC_LAUNCH(c, k_cv_apply, blocks_for(owned_bound(c, ORBC_PROTEIN), kBlock), kBlock, 0, P.C(), c->d_range + 2, c->cell_normal, P.X(), P.n,
                                            ptype, 1, target, strength, c->d_acc, vol_all, world, P.f);
    return ORBC_OK;
}

int alloc_voronoi(orbc_ctx *c, int nc) {
    if (nc == c->n_cells) return ORBC_OK;
    ORBC_TRY(dev_alloc(&c->centroid, nc)); ORBC_TRY(dev_alloc(&c->centroid_tmp, nc));
    ORBC_TRY(dev_alloc(&c->keys, nc)); ORBC_TRY(dev_alloc(&c->keys_tmp, nc)); ORBC_TRY(dev_alloc(&c->perm, nc)); ORBC_TRY(dev_alloc(&c->perm_tmp, nc)); ORBC_TRY(dev_alloc(&c->inv, nc));
    ORBC_TRY(dev_alloc(&c->grid.bin_items, nc)); ORBC_TRY(dev_alloc(&c->grid.bin_of, nc)); ORBC_TRY(dev_alloc(&c->grid.bin_slot, nc)); ORBC_TRY(dev_alloc(&c->grid.sorted, nc));
    ORBC_TRY(dev_alloc(&c->stencil, (size_t)nc * kStencilStride)); ORBC_TRY(dev_alloc(&c->stencil_cnt, nc));
    ORBC_TRY(dev_alloc(&c->wide, (size_t)nc * kStencilStride)); ORBC_TRY(dev_alloc(&c->wide_cnt, nc)); ORBC_TRY(dev_alloc(&c->cen_ref, nc)); ORBC_TRY(dev_alloc(&c->wide_ok, 3)); ORBC_TRY(dev_alloc(&c->movers, 2 * kMoversCap));
    c->wide_valid = false;
    ORBC_TRY(dev_alloc(&c->cell_normal, nc)); ORBC_TRY(dev_alloc(&c->lbound, nc)); ORBC_TRY(dev_alloc(&c->pbound, nc));
    ORBC_TRY(dev_alloc(&c->lruns, (size_t)nc * kRunStride)); ORBC_TRY(dev_alloc(&c->lrun_cnt, (size_t)nc)); c->lruns_cells = (size_t)nc;
    for (int s = 0; s < 2; ++s) ORBC_TRY(dev_alloc(&c->sp[s].cell_start, (size_t)nc + 1));
    c->n_cells = nc;
    return ORBC_OK;
}

int single_gpu_only(orbc_ctx *c, const char *what) {
    return mg_active(c) ? fail(ORBC_ERR_ARG, "%s is not available on a decomposed run yet", what) : ORBC_OK;
}

} // namespace

extern "C" {

const char *orbc_last_error(void) { return err_buf(); }

int orbc_forcefield_canonical(orbc_forcefield *ff) {
    // forcefield_canonical.h:30-156: the same constants and the same derivations (rep in fp32, att and lj_* in fp64)
    if (!ff) return fail(ORBC_ERR_ARG, "null forcefield");
    memset(ff, 0, sizeof(*ff));
    const float mass[6] = {1, 4, 4, 1, 10, 10}, radius[6] = {0.56125f, 1.12375f, 1.12375f, 0.56125f, 1.5f, 0.5f};
    const float cutlp[6] = {2.6f, 2.6f, 2.6f, 2.6f, 0, 0}, reqlp[6] = {1.1225f, 1.685f, 1.685f, 1.1225f, 0, 0};
    const float epslp[6] = {1.2f, 1.4f, 2.8f, 2.8f, 0, 0}, alphalp[6] = {1.55f, 5, 5, 5, 0, 0};
    for (int i = 0; i < 6; ++i) { ff->mass[i] = mass[i]; ff->radius[i] = radius[i]; ff->cutlp[i] = cutlp[i]; ff->alphalp[i] = alphalp[i]; }
    for (int i = 0; i < 4; ++i) {
        ff->cutsqlp[i] = cutlp[i] * cutlp[i];
        ff->replp[i] = ff_rep(cutlp[i], reqlp[i], epslp[i]);
        ff->attlp[i] = ff_att(cutlp[i], reqlp[i], epslp[i]);
    }
    ff->cutll = cutlp[0]; ff->cutsqll = ff->cutsqlp[0]; ff->repll = ff->replp[0]; ff->attll = ff->attlp[0]; ff->alphall = alphalp[0];
    // protein-protein: row/column 0 repeat the lipid-protein entries, the 3x3 block {1,2,3}^2 has its own req, eps = 1
    const float reqpp_in[3][3] = {{2.245f, 2.245f, 1.685f}, {2.245f, 2.245f, 1.685f}, {1.685f, 1.685f, 1.1225f}};
    for (int r = 0; r < 4; ++r) for (int q = 0; q < 4; ++q) {
        const int k = 6 * r + q;
        if (r == 0 || q == 0) { const int t = r + q; ff->cutpp[k] = cutlp[t]; ff->cutsqpp[k] = ff->cutsqlp[t]; ff->reppp[k] = ff->replp[t]; }
        else { ff->cutpp[k] = 2.6f; ff->cutsqpp[k] = 2.6f * 2.6f; ff->reppp[k] = ff_rep(2.6f, reqpp_in[r - 1][q - 1], 1.0f); }
    }
    // 12-6 LJ between {lipid, band-3} and {actin, spectrin}
    const float c1 = 1.1225f, c34 = (float)(3.4 * 1.1225);
    const float eps[36] = {0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 1, 1, 1, 0, 0, 0, 1, 1, 1, 0, 0, 0};
    const float sig[36] = {0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 3.4f, 3.4f, 0, 0, 0, 0, 3.4f, 1, 0, 0, 0, 0, 0, 0, 1, 3.4f, 3.4f, 0, 3, 1.8f, 1, 3.4f, 1, 0, 1.8f, 1};
    const float cut[36] = {0, 0, 0, 0, c1, c1, 0, 0, 0, 0, c34, c34, 0, 0, 0, 0, c34, c1, 0, 0, 0, 0, 0, 0, c1, c34, c34, 0, 0, 0, c1, c34, c1, 0, 0, 0};
    for (int i = 0; i < 36; ++i) {
        ff->lj_cutsq[i] = cut[i] * cut[i];
        ff->lj_lj1[i] = (float)(48.0 * eps[i] * std::pow((double)sig[i], 12.0));
        ff->lj_lj2[i] = (float)(24.0 * eps[i] * std::pow((double)sig[i], 6.0));
    }
    const float r0[4] = {2.25f, 1.1225f, 2.25f, 2.24f};
    for (int i = 0; i < 4; ++i) { ff->r0[i] = r0[i]; ff->K[i] = 57.f; }
    return ORBC_OK;
}

int orbc_set_forcefield(orbc_ctx *c, const orbc_forcefield *ff) { if (c) cudaSetDevice(c->device);
    if (!c || !ff) return fail(ORBC_ERR_ARG, "null argument");
    ORBC_CUDA(cudaMemcpyToSymbolAsync(c_ff, ff, sizeof(*ff), 0, cudaMemcpyHostToDevice, c->stream));
    ORBC_CUDA(cudaStreamSynchronize(c->stream));
    c->host_ff = *ff;
    c->ff_set = true;
    c->porder_valid = false;
    return ORBC_OK;
}

int orbc_create(orbc_ctx **out, int device) {
    if (!out) return fail(ORBC_ERR_ARG, "null ctx pointer");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail(ORBC_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
    if (device < 0 || device >= n_dev) return fail(ORBC_ERR_ARG, "device %d out of range (%d devices)", device, n_dev);
    ORBC_CUDA(cudaSetDevice(device));
    orbc_ctx *c = new orbc_ctx();
    c->device = device;
    ORBC_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    for (auto &e : c->ev) ORBC_CUDA(cudaEventCreate(&e));
    ORBC_TRY(dev_alloc(&c->d_acc, 8)); ORBC_TRY(dev_alloc(&c->d_counters, 8)); ORBC_TRY(dev_alloc(&c->d_flags, 4)); ORBC_TRY(dev_alloc(&c->d_check, 8)); ORBC_TRY(dev_alloc(&c->d_nh, 2));
    ORBC_TRY(dev_alloc(&c->d_range, 4)); ORBC_CUDA(cudaMemset(c->d_range, 0, 4 * sizeof(int)));
    ORBC_TRY(dev_alloc(&c->tile_overflow, 1)); ORBC_CUDA(cudaMemset(c->tile_overflow, 0, sizeof(int)));
    // the tile kernel wants the whole shared-memory carve-out: five blocks of four 11 KB warp tiles per SM
    ORBC_CUDA(cudaFuncSetAttribute((const void *)k_pair_ll_t, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    ORBC_CUDA(cudaMemset(c->d_acc, 0, 8 * sizeof(double))); ORBC_CUDA(cudaMemset(c->d_counters, 0, 8 * sizeof(unsigned long long)));
    ORBC_CUDA(cudaMemset(c->d_flags, 0, 4 * sizeof(int))); ORBC_CUDA(cudaMemset(c->d_nh, 0, 2 * sizeof(float)));
    ORBC_CUDA(cudaMallocHost((void **)&c->h_acc, 8 * sizeof(double))); ORBC_CUDA(cudaMallocHost((void **)&c->h_flags, 4 * sizeof(int))); ORBC_CUDA(cudaMallocHost((void **)&c->h_check, 8 * sizeof(int)));
    ORBC_CUDA(cudaMallocHost((void **)&c->h_counters, 8 * sizeof(unsigned long long))); ORBC_CUDA(cudaMallocHost((void **)&c->h_nh, 2 * sizeof(float)));
    orbc_forcefield ff; orbc_forcefield_canonical(&ff);
    *out = c;
    return orbc_set_forcefield(c, &ff);
}

void orbc_destroy(orbc_ctx *c) { if (c) cudaSetDevice(c->device);
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_species(c->sp[0]); free_species(c->sp[1]);
    dev_free(c->centroid); dev_free(c->centroid_tmp); dev_free(c->keys); dev_free(c->keys_tmp); dev_free(c->perm); dev_free(c->perm_tmp); dev_free(c->inv);
    dev_free(c->grid.bin_start); dev_free(c->grid.bin_items); dev_free(c->grid.bin_of); dev_free(c->grid.bin_slot); dev_free(c->grid.sorted);
    dev_free(c->stencil); dev_free(c->stencil_cnt); dev_free(c->wide); dev_free(c->wide_cnt); dev_free(c->cen_ref); dev_free(c->wide_ok); dev_free(c->movers); dev_free(c->cell_normal); dev_free(c->lbound); dev_free(c->pbound); dev_free(c->porder); dev_free(c->lruns); dev_free(c->lrun_cnt); dev_free(c->bonds); dev_free(c->tag2idx);
    dev_free(c->scan_tmp); dev_free(c->radix_hist); dev_free(c->stage); dev_free(c->d_acc); dev_free(c->d_counters); dev_free(c->d_flags); dev_free(c->d_check); dev_free(c->d_nh);
    dev_free(c->noise[0]); dev_free(c->noise[1]); dev_free(c->d_range); dev_free(c->tile_overflow);
    { NlState *st = (NlState *)c->nl_state; dev_free(st); } dev_free(c->ll_list); dev_free(c->ll_cnt); dev_free(c->pl_list); dev_free(c->pl_cnt); dev_free(c->pp_list); dev_free(c->pp_cnt);
    for (void *m : c->mg.opened) cudaIpcCloseMemHandle(m);
    dev_free(c->mg.my_bonds); dev_free(c->mg.keep); dev_free(c->mg.ke_all); dev_free(c->mg.vol_all); dev_free(c->mg.cv_ptype); dev_free(c->mg.flags); dev_free(c->mg.dest_mask); dev_free(c->mg.pmask); dev_free(c->mg.need);
    for (int s = 0; s < 2; ++s) { dev_free(c->mg.cnt_all[s]); dev_free(c->mg.off_me[s]); dev_free(c->mg.cnt_prev[s]); }
    if (c->h_acc) cudaFreeHost(c->h_acc); if (c->h_flags) cudaFreeHost(c->h_flags); if (c->h_check) cudaFreeHost(c->h_check); if (c->h_counters) cudaFreeHost(c->h_counters); if (c->h_nh) cudaFreeHost(c->h_nh);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    for (auto &v : c->prof_ev) for (auto &e : v) cudaEventDestroy(e);
    for (auto &e : c->kprof_ev) cudaEventDestroy(e);
    for (auto &e : c->xfer_ev) if (e) cudaEventDestroy(e);
    for (int k = 0; k < 2; ++k) {
        dev_free(c->frame_dev[k]); if (c->frame_host[k]) cudaFreeHost(c->frame_host[k]);
        if (c->frame_packed[k]) cudaEventDestroy(c->frame_packed[k]); if (c->frame_copied[k]) cudaEventDestroy(c->frame_copied[k]);
    }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int orbc_set_option(orbc_ctx *c, const char *name, double value) { if (c) cudaSetDevice(c->device);
    if (!c || !name) return fail(ORBC_ERR_ARG, "null argument");
    if (!strcmp(name, "pair_impl")) { if (value != 1 && value != 2) return fail(ORBC_ERR_ARG, "pair_impl must be 1 or 2"); c->pair_impl = (int)value; c->nl_valid = false; return ORBC_OK; }
    if (!strcmp(name, "prot_lanes")) { if (value != 0 && value != 1 && value != 2 && value != 4) return fail(ORBC_ERR_ARG, "prot_lanes must be 0 (automatic), 1, 2 or 4"); c->prot_lanes = (int)value; c->nl_valid = false; return ORBC_OK; }
    if (!strcmp(name, "ll_variant")) {                           // 0: warp-per-cell tile kernel k_pair_ll_t (default); 1: thread-per-lipid run-list kernel k_pair_ll_r
        if (value != 0 && value != 1) return fail(ORBC_ERR_ARG, "ll_variant must be 1 (run-list kernel + hit lists) or 0 (tile kernel)");
        c->ll_variant = (int)value; c->nl_valid = false; return ORBC_OK;
    }
    if (!strcmp(name, "nl_reuse")) {                             // hit lists between rebuilds: 0 off, 1 automatic (default), 2 on
        if (value != 0 && value != 1 && value != 2) return fail(ORBC_ERR_ARG, "nl_reuse must be 0, 1 or 2");
        c->nl_on = (int)value; c->nl_valid = false;
        if (mg_active(c) && c->mg.connected && nl_active(c)) ORBC_TRY(nl_ensure(c));   // (not at the first launch, while peers may wait in a barrier)
        return ORBC_OK;
    }
    if (!strcmp(name, "nl_skin_max")) {
        if (!(value >= 0.0 && value <= 1.0)) return fail(ORBC_ERR_ARG, "nl_skin_max must be in [0, 1]");
        c->nl_skin_max = (float)value; c->nl_valid = false; return ORBC_OK;
    }
    if (!strcmp(name, "nl_skin")) {
        if (!(value >= 0.0 && value <= 1.0)) return fail(ORBC_ERR_ARG, "nl_skin must be in [0, 1]");
        c->nl_skin = (float)value; c->nl_valid = false; return ORBC_OK;
    }
    if (!strcmp(name, "debug_tile_cap")) {                       // test aid: a smaller tile capacity, so that small systems reach the overflow path
        if (!(value >= 1 && value <= kTileCap)) return fail(ORBC_ERR_ARG, "debug_tile_cap must be in [1, %d]", kTileCap);
        c->tile_cap = (int)value; c->lruns_valid = false; return ORBC_OK;
    }
    if (!strcmp(name, "stencil_refresh")) { c->wide_on = value != 0; return ORBC_OK; }   // 0: every rebuild searches the centroid grid in full
    if (!strcmp(name, "debug_nl_cap")) {                         // test aid: rows of the hit lists this short (before the lists are first used), so that they overflow
        if (!(value >= 1 && value <= 96)) return fail(ORBC_ERR_ARG, "debug_nl_cap must be in [1, 96]");
        if (c->ll_list || c->pl_list) return fail(ORBC_ERR_STATE, "debug_nl_cap: the lists are allocated already");
        c->nl_cap_ll = (int)value; c->nl_cap_pl = std::min(c->nl_cap_pl, (int)value); c->nl_cap_pp = std::min(c->nl_cap_pp, (int)value);
        return ORBC_OK;
    }
    if (!strcmp(name, "debug_nl_mode")) {                        // measurement aid: -1 the gate decides, 1 every evaluation records the hit lists, 2 every evaluation searches
        if (value != -1 && value != 1 && value != 2) return fail(ORBC_ERR_ARG, "debug_nl_mode must be -1, 1 or 2");
        c->nl_debug_mode = (int)value; c->nl_valid = false; return ORBC_OK;
    }
    if (!strcmp(name, "debug_own_slack")) {                      // test aid: slack of the owned-particle launch bounds (before orbc_mg_export), so that
        if (!(value >= 0)) return fail(ORBC_ERR_ARG, "debug_own_slack must be >= 0");   // small systems get launch bounds below their size
        c->mg_own_slack = (int)value; return ORBC_OK;
    }
    if (!strcmp(name, "debug_barriers")) {                       // profiling aid: `value` back-to-back barriers of a decomposed run
        for (int k = 0; k < (int)value; ++k) ORBC_TRY(mg_barrier(c));
        return ORBC_OK;
    }
    if (!strcmp(name, "debug_owned_fraction")) {
        // test / profiling aid: compute only the first `value` of the slots of both containers, as one rank of a decomposed run
        // would (forces and integration of the other slots are skipped; results are partial by construction)
        if (!(value > 0.0 && value <= 1.0)) return fail(ORBC_ERR_ARG, "debug_owned_fraction must be in (0, 1]");
        for (int sp = 0; sp < 2; ++sp) ORBC_LAUNCH(c, k_set_range_const, 1, 1, 0, c->d_range + 2 * sp, 0, (int)(c->sp[sp].n * value));
        c->porder_valid = false; c->nl_valid = false;
        return ORBC_OK;
    }
    return fail(ORBC_ERR_ARG, "unknown option '%s'", name);
}

int orbc_synchronize(orbc_ctx *c) { if (c) cudaSetDevice(c->device); ORBC_CUDA(cudaStreamSynchronize(c->stream)); return check_flags(c); }
int orbc_set_stream(orbc_ctx *c, void *s) { if (c) cudaSetDevice(c->device); ORBC_CUDA(cudaStreamSynchronize(c->stream)); c->stream = s ? (cudaStream_t)s : c->own_stream; return ORBC_OK; }

// second stream for host <-> device copies that overlap the packing kernels (also used by the asynchronous frames)
static int ensure_copy_stream(orbc_ctx *c) {
    if (!c->copy_stream) {
        ORBC_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) { ORBC_CUDA(cudaEventCreateWithFlags(&c->frame_packed[k], cudaEventDisableTiming)); ORBC_CUDA(cudaEventCreateWithFlags(&c->frame_copied[k], cudaEventDisableTiming)); }
        for (auto &e : c->xfer_ev) ORBC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    return ORBC_OK;
}

// the upload proper: rows [first, first + count) of the host's strided arrays, type / tag of EVERY slot.  All host-to-device
// copies are queued back to back on the copy stream into one staging area (the link never idles); the packing kernels follow
// each copy on the compute stream.
static int upload_rows(orbc_ctx *c, int sp, size_t n, size_t first, size_t count, size_t stride,
                       const float *x, const float *v, const float *n_, const float *o, const int *type, const int *tag) {
    Species &S = c->sp[sp];
    ORBC_TRY(alloc_species(c, S, n));
    if (!n) return ORBC_OK;
    ORBC_TRY(ensure_copy_stream(c));
    const size_t rows = count * stride;
    ORBC_TRY(ensure_stage(c, 4 * rows + 2 * n + 16));
    float *stage[4] = {c->stage, c->stage + rows, c->stage + 2 * rows, c->stage + 3 * rows};
    int *w_type = (int *)(c->stage + 4 * rows), *w_tag = w_type + n;
    const float *src[4] = {x, n_, v, o};
    float4 *dst[4] = {S.X(), S.N(), S.V(), S.O()};
    const int *wsrc[4] = {type, tag, nullptr, nullptr};
    int *wdev[4] = {w_type, w_tag, nullptr, nullptr};
    ORBC_CUDA(cudaEventRecord(c->xfer_ev[6], c->stream));
    ORBC_CUDA(cudaStreamWaitEvent(c->copy_stream, c->xfer_ev[6], 0));   // the staging area may still be read by earlier kernels
    for (int k = 0; k < 4; ++k) {
        if (wsrc[k]) ORBC_CUDA(cudaMemcpyAsync(wdev[k], wsrc[k], sizeof(int) * n, cudaMemcpyHostToDevice, c->copy_stream));
        if (src[k] && count) ORBC_CUDA(cudaMemcpyAsync(stage[k], src[k] + first * stride, sizeof(float) * rows, cudaMemcpyHostToDevice, c->copy_stream));
        ORBC_CUDA(cudaEventRecord(c->xfer_ev[k], c->copy_stream));
    }
    const unsigned nb = blocks_for(n, kBlock), nbr = blocks_for(std::max<size_t>(count, 1), kBlock);
    for (int k = 0; k < 4; ++k) {
        ORBC_CUDA(cudaStreamWaitEvent(c->stream, c->xfer_ev[k], 0));
        // type / tag go into .w of every slot; the vectors into the uploaded rows (zeros when the host passes no array)
        if (k < 2 && count < n) ORBC_LAUNCH(c, k_zero4, nb, kBlock, 0, dst[k], n, (const int *)(wsrc[k] ? wdev[k] : nullptr));
        if (count) {
            if (src[k]) ORBC_LAUNCH(c, k_pack4, nbr, kBlock, 0, stage[k], stride, count, dst[k] + first, (const int *)(wsrc[k] ? wdev[k] + first : nullptr));
            else ORBC_LAUNCH(c, k_zero4, nbr, kBlock, 0, dst[k] + first, count, (const int *)(wsrc[k] ? wdev[k] + first : nullptr));
        }
    }
    ORBC_LAUNCH(c, k_zero4, nb, kBlock, 0, S.f, n, (const int *)nullptr);
    ORBC_LAUNCH(c, k_zero4, nb, kBlock, 0, S.t, n, (const int *)nullptr);
    ORBC_LAUNCH(c, k_fill_int, nb, kBlock, 0, S.C(), n, -1);
    // type and tag are checked where they now are (k_check_ids); the call returns only when the host arrays are free again
    if (sp == ORBC_PROTEIN) {
        c->porder_valid = false;
        ORBC_CUDA(cudaMemsetAsync(c->d_check, 0, 4 * sizeof(int), c->stream));
        ORBC_LAUNCH(c, k_check_ids, nb, kBlock, 0, (const int *)(type ? w_type : nullptr), (const int *)(tag ? w_tag : nullptr), n, kNType, c->d_check);
        ORBC_CUDA(cudaMemcpyAsync(c->h_check, c->d_check, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    }
    ORBC_CUDA(cudaStreamSynchronize(c->stream));
    int mx = 0;
    if (sp == ORBC_PROTEIN) {
        const int *h = c->h_check;
        if (h[0]) { const size_t i = n - (size_t)h[0]; return fail(ORBC_ERR_ARG, "protein %zu has type %d outside [0,%d)", i, type[i], kNType); }
        c->type_mask = type ? (unsigned)h[1] : 1u;
        if (h[2]) return fail(ORBC_ERR_ARG, "negative protein tag");
        mx = h[3];
    }
    if (sp == ORBC_PROTEIN && tag) {
        if (c->tag2idx_size < (size_t)mx + 1) { ORBC_TRY(dev_alloc(&c->tag2idx, (size_t)mx + 1)); c->tag2idx_size = (size_t)mx + 1; }
        ORBC_CUDA(cudaMemsetAsync(c->tag2idx, 0xff, sizeof(int) * c->tag2idx_size, c->stream));
        ORBC_TRY(build_tag2idx(c));
    }
    return ORBC_OK;
}

int orbc_upload(orbc_ctx *c, int sp, size_t n, size_t stride, const float *x, const float *v, const float *n_, const float *o, const int *type, const int *tag) { if (c) cudaSetDevice(c->device);
    if (!c || sp < 0 || sp > 1 || stride < 3) return fail(ORBC_ERR_ARG, "orbc_upload: bad argument");
    if (n && (!x || !n_)) return fail(ORBC_ERR_ARG, "orbc_upload: x and n are required");
    if (n >= (size_t)1 << 31) return fail(ORBC_ERR_ARG, "orbc_upload: more than 2^31 particles per container");
    c->mg.partial[sp] = false;
    return upload_rows(c, sp, n, 0, n, stride, x, v, n_, o, type, tag);
}

int orbc_upload_range(orbc_ctx *c, int sp, size_t n, size_t first, size_t count, size_t stride,
                      const float *x, const float *v, const float *n_, const float *o, const int *type, const int *tag) { if (c) cudaSetDevice(c->device);
    if (!c || sp < 0 || sp > 1 || stride < 3 || first + count > n) return fail(ORBC_ERR_ARG, "orbc_upload_range: bad argument");
    if (count && (!x || !n_)) return fail(ORBC_ERR_ARG, "orbc_upload_range: x and n are required");
    if (n >= (size_t)1 << 31) return fail(ORBC_ERR_ARG, "orbc_upload_range: more than 2^31 particles per container");
    if (!c->mg.connected) return fail(ORBC_ERR_ARG, "orbc_upload_range: only on the connected ranks of a decomposed run (the first upload is a whole one)");
    if (!c->sp[sp].cap || n + 64 > c->sp[sp].cap) return fail(ORBC_ERR_ARG, "orbc_upload_range: the container would have to grow (peer mappings are fixed): upload the whole system once first");
    c->mg.partial[sp] = count < n;
    return upload_rows(c, sp, n, first, count, stride, x, v, n_, o, type, tag);
}

int orbc_upload_bonds(orbc_ctx *c, size_t n_bonds, const int *tij) { if (c) cudaSetDevice(c->device);
    if (!c || (n_bonds && !tij)) return fail(ORBC_ERR_ARG, "orbc_upload_bonds: bad argument");
    if (!c->bonds || c->bonds_cap < 3 * n_bonds) { ORBC_TRY(dev_alloc(&c->bonds, 3 * n_bonds)); c->bonds_cap = 3 * n_bonds; }   // a re-upload keeps the allocation
    c->n_bonds = 0;                                              // (until the new list has been checked)
    if (n_bonds) ORBC_CUDA(cudaMemcpyAsync(c->bonds, tij, sizeof(int) * 3 * n_bonds, cudaMemcpyHostToDevice, c->stream));
    // checked on the device (k_check_bonds): the host's sweep over the 3 n integers cost ten times the copy
    ORBC_CUDA(cudaMemsetAsync(c->d_check + 4, 0, sizeof(int), c->stream));
    if (n_bonds) ORBC_LAUNCH(c, k_check_bonds, blocks_for(n_bonds, kBlock), kBlock, 0, c->bonds, n_bonds, c->tag2idx_size, c->d_check);
    ORBC_CUDA(cudaMemcpyAsync(c->h_check + 4, c->d_check + 4, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ORBC_CUDA(cudaStreamSynchronize(c->stream));
    if (c->h_check[4]) {
        const size_t b = n_bonds - (size_t)c->h_check[4];
        if (tij[3 * b] < 0 || tij[3 * b] >= 4) return fail(ORBC_ERR_ARG, "bond %zu has type %d outside [0,4)", b, tij[3 * b]);
        return fail(ORBC_ERR_ARG, "bond %zu refers to a tag that no uploaded protein carries (upload proteins first)", b);
    }
    c->n_bonds = n_bonds;
    return ORBC_OK;
}

int orbc_voronoi_upload(orbc_ctx *c, int nc, const float *centroids3, const int *csl, const int *csp) { if (c) cudaSetDevice(c->device);
    if (!c || nc <= 0 || !centroids3) return fail(ORBC_ERR_ARG, "orbc_voronoi_upload: bad argument");
    if (nc >= (1 << 28)) return fail(ORBC_ERR_ARG, "more than 2^28 Voronoi cells");
    ORBC_TRY(alloc_voronoi(c, nc));
    c->nl_valid = false;
    ORBC_CUDA(cudaMemsetAsync(c->cell_normal, 0, sizeof(float4) * nc, c->stream));
    ORBC_TRY(ensure_stage(c, (size_t)3 * nc));
    ORBC_CUDA(cudaMemcpyAsync(c->stage, centroids3, sizeof(float) * 3 * nc, cudaMemcpyHostToDevice, c->stream));
    ORBC_LAUNCH(c, k_pack4, blocks_for(nc, kBlock), kBlock, 0, c->stage, (size_t)3, (size_t)nc, c->centroid, (const int *)nullptr);
    ORBC_TRY(fit_grid(c, centroids3, nc));
    for (int s = 0; s < 2; ++s) {
        Species &S = c->sp[s];
        const int *cs = s == 0 ? csl : csp;
        S.has_partition = false;
        if (cs) {
            if ((size_t)cs[nc] != S.n || cs[0] != 0) return fail(ORBC_ERR_ARG, "cell_start of species %d does not cover its %zu particles", s, S.n);
            ORBC_CUDA(cudaMemcpyAsync(S.cell_start, cs, sizeof(int) * ((size_t)nc + 1), cudaMemcpyHostToDevice, c->stream));
            if (S.n) ORBC_LAUNCH(c, k_fill_cellid, blocks_for(nc, kBlock), kBlock, 0, S.cell_start, nc, S.C());
            S.has_partition = true;
        }
    }
    ORBC_TRY(build_index(c));
    ORBC_CUDA(cudaStreamSynchronize(c->stream));   // host arrays are borrowed only for the duration of the call
    return ORBC_OK;
}

int orbc_voronoi_init(orbc_ctx *c, int nc, int n_iterate) { if (c) cudaSetDevice(c->device);
    // VoronoiDiagram::init (voronoi.h:54-75): initial guess = every delta-th lipid; n_iterate rounds of { Morton sort of the centroids,
    // index rebuild (tree.build), partition, centroid update through VCellList::cells }, the container being reordered only at
    // the rounds k = 0, 1, 2, 4, 8, ... and once more at the end.
    if (!c || nc <= 0 || n_iterate <= 0) return fail(ORBC_ERR_ARG, "orbc_voronoi_init: bad argument");
    ORBC_TRY(single_gpu_only(c, "orbc_voronoi_init"));
    Species &L = c->sp[0];
    if (!L.n) return fail(ORBC_ERR_ARG, "orbc_voronoi_init: upload the lipids first");
    if (nc >= (1 << 28)) return fail(ORBC_ERR_ARG, "more than 2^28 Voronoi cells");
    const int last = n_iterate - 1;
    if ((last & (~last + 1)) == last)
        return fail(ORBC_ERR_ARG, "orbc_voronoi_init: with n_iterate - 1 = %d a power of two (or zero) the reference applies its last permutation twice (voronoi.h:70,74) and leaves the "
                    "container unsorted; use another count (the driver uses 64)", last);
    ORBC_TRY(alloc_voronoi(c, nc));
    ORBC_CUDA(cudaMemsetAsync(c->cell_normal, 0, sizeof(float4) * nc, c->stream));
    // grid over the bounding box of the lipids: every centroid is a mean of lipid positions and stays inside it
    int *bb_dev = nullptr; ORBC_TRY(dev_alloc(&bb_dev, 6));
    const int init_bb[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
    ORBC_CUDA(cudaMemcpyAsync(bb_dev, init_bb, sizeof(init_bb), cudaMemcpyHostToDevice, c->stream));
    ORBC_LAUNCH(c, k_bbox, blocks_for(L.n, kBlock), kBlock, 0, L.X(), L.n, bb_dev);
    int hb[6];
    ORBC_CUDA(cudaMemcpyAsync(hb, bb_dev, sizeof(hb), cudaMemcpyDeviceToHost, c->stream));
    ORBC_CUDA(cudaStreamSynchronize(c->stream));
    dev_free(bb_dev);
    float lo[3], hi[3];
    for (int d = 0; d < 6; ++d) {
        const int k = hb[d], bits = k >= 0 ? k : k ^ 0x7fffffff;
        float v; memcpy(&v, &bits, sizeof(v));
        (d < 3 ? lo[d] : hi[d - 3]) = v;
    }
    ORBC_TRY(fit_grid_box(c, lo, hi));
    ORBC_LAUNCH(c, k_init_centroids, blocks_for(nc, kBlock), kBlock, 0, L.X(), L.n, nc, L.n / (size_t)nc, c->centroid);
    ORBC_LAUNCH(c, k_fill_int, blocks_for(L.n, kBlock), kBlock, 0, L.C(), L.n, -1);
    L.has_partition = false;
    CentroidOut out; out.world = 1;
    for (int k = 0; k < n_iterate; ++k) {
        // reorder_morton(centroids, param); tree.build(centroids)
        ORBC_LAUNCH(c, k_morton_keys, blocks_for(nc, kBlock), kBlock, 0, c->centroid, nc, c->keys, c->perm);
        ORBC_TRY(radix_sort_pairs(c, c->keys, c->perm, c->keys_tmp, c->perm_tmp, nc));
        ORBC_LAUNCH(c, k_permute_centroids, blocks_for(nc, kBlock), kBlock, 0, c->centroid, c->perm, nc, c->centroid_tmp, c->inv);
        std::swap(c->centroid, c->centroid_tmp);
        if (k > 0) ORBC_LAUNCH(c, k_remap_cellid, blocks_for(L.n, kBlock), kBlock, 0, L.C(), c->d_range, c->inv);   // last round's cell = this round's search hint
        ORBC_TRY(build_index(c));
        // cell_list.partition(cont, *this)
        ORBC_TRY(cell_update_assign(c, 1));
        const bool reorder = (k & (~k + 1)) == k || k == last;
        if (reorder) {
            ORBC_TRY(cell_update_move(c, 1));                     // scan, arrival lists, rank + move (reorder.h:73-149)
            for (int r = 0; r < kMaxWorld; ++r) out.dst[r] = c->centroid;
            ORBC_LAUNCH(c, k_centroid_update, blocks_for(nc, kBlock), kBlock, 0, L.cell_start, L.X(), 0, nc, out, (const int *)nullptr);
        } else {
            ORBC_TRY(scan_exclusive(c, L.cell_start, nc));
            ORBC_LAUNCH(c, k_cell_scatter, blocks_for(L.n, kBlock), kBlock, 0, ScatterArgs{L.aff, L.li, c->d_range, L.cell_start, (const int *)nullptr, L.cells_tmp});
            ORBC_LAUNCH(c, k_rank_only, blocks_for(L.n, kBlock), kBlock, 0, L.aff, c->d_range, L.cell_start, L.cells_tmp, L.cells);
            for (int r = 0; r < kMaxWorld; ++r) out.dst[r] = c->centroid;
            ORBC_LAUNCH(c, k_centroid_update, blocks_for(nc, kBlock), kBlock, 0, L.cell_start, L.X(), 0, nc, out, L.cells);
            ORBC_CUDA(cudaMemcpyAsync(L.C(), L.aff, sizeof(int) * L.n, cudaMemcpyDeviceToDevice, c->stream));
            L.has_partition = true;                              // (as a search hint only: the container is not sorted by cell here)
        }
    }
    // the centroids moved once more after the last partition: the index follows them (the reference leaves its tree one update behind)
    ORBC_TRY(build_index(c));
    c->sp[1].has_partition = false;
    ORBC_CUDA(cudaStreamSynchronize(c->stream));
    return check_flags(c);
}

int orbc_set_field(orbc_ctx *c, int sp, char field, size_t stride, const float *src) { if (c) cudaSetDevice(c->device);
    if (!c || sp < 0 || sp > 1 || stride < 3 || !src) return fail(ORBC_ERR_ARG, "orbc_set_field: bad argument");
    Species &S = c->sp[sp];
    float4 *dst = field == 'f' ? S.f : field == 't' ? S.t : field == 'v' ? S.V() : field == 'x' ? S.X() : field == 'n' ? S.N() : field == 'o' ? S.O() : nullptr;
    if (!strchr("ftvxno", field)) return fail(ORBC_ERR_ARG, "orbc_set_field: unknown field '%c'", field);
    if (!S.n) return ORBC_OK;
    if (field == 'x') c->nl_valid = false;
    ORBC_TRY(ensure_stage(c, S.n * stride));
    ORBC_CUDA(cudaMemcpyAsync(c->stage, src, sizeof(float) * S.n * stride, cudaMemcpyHostToDevice, c->stream));
    ORBC_LAUNCH(c, k_set3, blocks_for(S.n, kBlock), kBlock, 0, c->stage, stride, S.n, dst);
    ORBC_CUDA(cudaStreamSynchronize(c->stream));
    return ORBC_OK;
}

int orbc_voronoi_update(orbc_ctx *c, int nstep, int freq_sort_ctrd) { if (c) cudaSetDevice(c->device); if (!c) return fail(ORBC_ERR_ARG, "null ctx"); return do_voronoi_update(c, nstep, freq_sort_ctrd); }

int orbc_cell_update(orbc_ctx *c, int sp, int nstep, int freq_sort_bond) { if (c) cudaSetDevice(c->device);
    (void)nstep; (void)freq_sort_bond;   // reorder_bond (reorder.h:33-68) only permutes bond storage for CPU cache locality
    if (!c || sp < 0 || sp > 1) return fail(ORBC_ERR_ARG, "orbc_cell_update: bad argument");
    return do_cell_update(c, sp);
}

int do_rebuild(orbc_ctx *c, int nstep, int freq_sort_ctrd) {
    ProfScope ps(c, ORBC_PROF_REBUILD);
    ORBC_TRY(do_voronoi_update(c, nstep, freq_sort_ctrd));
    ORBC_TRY(cell_update_assign(c, kBothContainers));
    ORBC_TRY(mg_barrier(c));
    ORBC_TRY(cell_update_move(c, kBothContainers));
    ORBC_TRY(mg_barrier(c));
    ORBC_TRY(cell_update_finish(c, kBothContainers));
    return mg_barrier(c);
}

int orbc_rebuild(orbc_ctx *c, int nstep, int freq_sort_ctrd, int freq_sort_bond) { if (c) cudaSetDevice(c->device);
    (void)freq_sort_bond;
    if (!c) return fail(ORBC_ERR_ARG, "null ctx");
    return do_rebuild(c, nstep, freq_sort_ctrd);
}

int orbc_delete_lipid(orbc_ctx *c, float tol, size_t *n_out) { if (c) cudaSetDevice(c->device);
    if (!c) return fail(ORBC_ERR_ARG, "null ctx");
    Species &L = c->sp[0];
    if (!L.has_partition) return fail(ORBC_ERR_ARG, "delete_lipid: lipids are not partitioned");
    const int nc = c->n_cells;
    const size_t n = L.n;
    if (mg_active(c)) {
        // decomposed: every rank masks the strays of its own cells, then the lipid cell update runs with the mask — a deleted
        // lipid is not counted and not moved, so the survivors close ranks in the global slot order (order-preserving
        // compaction + cell_lipid.update of cleanup.h:62-85 in one pass).  All ranks learn the new size from the same cell_start.
        if (c->mg.ce > c->mg.cb)
            ORBC_LAUNCH(c, k_stray_mask, blocks_for((size_t)(c->mg.ce - c->mg.cb) * 32, kBlock), kBlock, 0, L.cell_start, c->mg.cb, c->mg.ce, c->centroid, L.X(), tol, c->mg.keep);
        // nothing to delete anywhere: the reference leaves the partition alone (cleanup.h:62 `if ( size_new < n )`), so must every rank —
        // a re-partition here would change the membership the next voronoi.update averages over.  The ranks add up their counts.
        ORBC_CUDA(cudaMemsetAsync(c->d_acc, 0, sizeof(double), c->stream));
        ORBC_LAUNCH(c, k_count_strays, blocks_for(owned_bound(c, ORBC_LIPID), kBlock), kBlock, 0, c->mg.keep, c->d_range, c->d_acc);
        ORBC_TRY(mg_share_ke(c)); ORBC_TRY(mg_barrier(c));
        ORBC_LAUNCH(c, k_sum_ke, 1, 32, 0, c->d_acc, mg_ke_slots(c), c->mg.world);
        ORBC_CUDA(cudaMemcpyAsync(c->h_acc, c->d_acc, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        ORBC_CUDA(cudaStreamSynchronize(c->stream));
        if (c->h_acc[0] == 0.0) { if (n_out) *n_out = L.n; return check_flags(c); }
        ORBC_TRY(cell_update_assign(c, 1, c->mg.keep)); ORBC_TRY(mg_barrier(c));
        ORBC_TRY(cell_update_move(c, 1));               ORBC_TRY(mg_barrier(c));
        ORBC_TRY(cell_update_finish(c, 1));             ORBC_TRY(mg_barrier(c));
        int total = 0;
        ORBC_CUDA(cudaMemcpyAsync(&total, L.cell_start + nc, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        ORBC_CUDA(cudaStreamSynchronize(c->stream));
        L.n = (size_t)total;
        if (n_out) *n_out = L.n;
        return check_flags(c);
    }
    int *keep = L.aff, *newpos = L.li;
    ORBC_LAUNCH(c, k_stray_mask, blocks_for((size_t)nc * 32, kBlock), kBlock, 0, L.cell_start, 0, nc, c->centroid, L.X(), tol, keep);
    ORBC_CUDA(cudaMemcpyAsync(newpos, keep, sizeof(int) * n, cudaMemcpyDeviceToDevice, c->stream));
    ORBC_TRY(scan_exclusive(c, newpos, (int)n));
    int total = 0;
    ORBC_CUDA(cudaMemcpyAsync(&total, newpos + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ORBC_CUDA(cudaStreamSynchronize(c->stream));
    if ((size_t)total < n) {
        const int nx = L.cur ^ 1, nxn = L.cur_xn ^ 1;
        ORBC_LAUNCH(c, k_compact, blocks_for(n, kBlock), kBlock, 0, keep, newpos, n, L.X(), L.N(), L.V(), L.O(), L.C(),
                    L.x[nxn], L.nn[nxn], L.v[nx], L.o[nx], L.cellid[nx]);
        L.cur = nx; L.cur_xn = nxn; L.n = (size_t)total;
        ORBC_LAUNCH(c, k_set_range_const, 1, 1, 0, c->d_range, 0, total);
        // f and t are zero at this point of the loop (cleared by the integrator); keep them so for the survivors
        ORBC_LAUNCH(c, k_zero4, blocks_for(n, kBlock), kBlock, 0, L.f, n, (const int *)nullptr);
        ORBC_LAUNCH(c, k_zero4, blocks_for(n, kBlock), kBlock, 0, L.t, n, (const int *)nullptr);
        ORBC_TRY(do_cell_update(c, ORBC_LIPID));   // cleanup.h:85
    }
    if (n_out) *n_out = L.n;
    return ORBC_OK;
}

int orbc_compute_pairwise_fused(orbc_ctx *c) { if (c) cudaSetDevice(c->device); if (!c) return fail(ORBC_ERR_ARG, "null ctx"); return launch_pairwise(c); }
int orbc_compute_bonded(orbc_ctx *c) { if (c) cudaSetDevice(c->device); if (!c) return fail(ORBC_ERR_ARG, "null ctx"); return launch_bonded(c); }

int orbc_constrain_volume(orbc_ctx *c, float target, float strength, float *volume_out) { if (c) cudaSetDevice(c->device);
    if (!c) return fail(ORBC_ERR_ARG, "null ctx");
    ORBC_TRY(do_constrain_volume(c, target, strength));
    if (volume_out) {
        ORBC_CUDA(cudaMemcpyAsync(c->h_acc, c->d_acc, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        ORBC_CUDA(cudaStreamSynchronize(c->stream));
        *volume_out = (float)c->h_acc[4];
    }
    return ORBC_OK;
}

int orbc_set_volume_constraint(orbc_ctx *c, int on, float target, float strength) {
    if (!c) return fail(ORBC_ERR_ARG, "null ctx");
    if (on && !(target != 0.f)) return fail(ORBC_ERR_ARG, "orbc_set_volume_constraint: target volume must not be zero");
    c->cv_on = on != 0; c->cv_target = target; c->cv_strength = strength;
    return ORBC_OK;
}

float orbc_nh_zeta_update(float zeta, float *Q, double dt, float kBT, double ke, long n) {
    // destructor of verlet_initial_bounce_clearforce_update / post_toque_final_update (integrate_nh.h:181-185, 240-244)
    if (!*Q) *Q = (float)(n * 0.01);
    zeta += 0.5 * dt / *Q * (ke - 0.5 * 3.0 * n * kBT);
    return zeta;
}

float orbc_nh_zeta_update_unfused(float zeta, float *Q, double dt, float kBT, double ke, long n) {
    // destructor of the unfused verlet_nh_update (integrate_nh.h:72-76): the target kinetic energy is a product of floats there
    // (constant::onehalf * n * parameter.kBT with real kBT), not the fused kernels' double expression
    if (!*Q) *Q = (float)(n * 0.01);
    const float target = 1.5f * (float)n * kBT;
    zeta += 0.5f * dt / *Q * (ke - target);
    return zeta;
}

int orbc_integrate(orbc_ctx *c, int kernel, const orbc_step_params *p, orbc_step_result *res) { if (c) cudaSetDevice(c->device);
    if (!c) return fail(ORBC_ERR_ARG, "null ctx");
    const bool needs_p = kernel != ORBC_CLEAR_FORCE && kernel != ORBC_POST_TORQUE;
    if (needs_p && !p) return fail(ORBC_ERR_ARG, "orbc_integrate: this kernel needs step parameters");
    const bool reduces = kernel == ORBC_NH_INITIAL_FUSED || kernel == ORBC_NH_FINAL_FUSED || kernel == ORBC_NH_UPDATE;
    if (reduces) ORBC_CUDA(cudaMemsetAsync(c->d_acc, 0, sizeof(double), c->stream));
    const bool mg_ok = kernel == ORBC_VERLET_LANGEVIN || kernel == ORBC_CLEAR_FORCE || kernel == ORBC_NH_INITIAL_FUSED || kernel == ORBC_NH_FINAL_FUSED || kernel == ORBC_OPT_FUSED;
    if (!mg_ok) ORBC_TRY(single_gpu_only(c, "this integrate() kernel"));
    if (kernel == ORBC_NH_INITIAL_FUSED) ++c->nl_moves;          // tracked by the kernel (hit lists)
    if (kernel == ORBC_BOUNCE_BACK || kernel == ORBC_OPT_MOVE || kernel == ORBC_OPT_FUSED) c->nl_valid = false;   // untracked moves
    if (kernel == ORBC_VERLET_LANGEVIN) ORBC_TRY(do_integrate_langevin(c, p));
    else for (int sp = 0; sp < 2; ++sp) {
        Species &S = c->sp[sp];
        if (!S.n) continue;
        const unsigned nb = blocks_for(owned_bound(c, sp), 256);
        IntegArgs a;
        if (p) fill_integ(a, c, sp, p);
        switch (kernel) {
        // the whole array, not the owned range: slots a rank stops owning must not keep stale sums for the accumulating force kernels
        case ORBC_CLEAR_FORCE: ORBC_LAUNCH(c, k_clear_force, blocks_for(S.n, 256), 256, 0, S.f, S.t, S.n); break;
        case ORBC_POST_TORQUE: ORBC_LAUNCH(c, k_post_torque, nb, 256, 0, S.N(), S.t, S.n); break;
        case ORBC_BOUNCE_BACK: ORBC_LAUNCH(c, k_bounce_back, nb, 256, 0, a); break;
        case ORBC_NH_INITIAL_FUSED: ORBC_LAUNCH(c, k_nh_initial_fused, nb, 256, 0, a); if (mg_active(c)) S.cur_xn ^= 1; if (sp == 1 || !c->sp[1].n) ORBC_TRY(nl_share(c)); break;
        case ORBC_NH_FINAL_FUSED: ORBC_LAUNCH(c, k_nh_final_fused, nb, 256, 0, a); break;
        case ORBC_NH_FINAL: ORBC_LAUNCH(c, k_nh_final, nb, 256, 0, a); break;
        case ORBC_NH_UPDATE: ORBC_LAUNCH(c, k_kinetic, nb, 256, 0, S.X(), S.V(), c->d_range + 2 * sp, 0.5f, c->d_acc); break;
        case ORBC_OPT_MOVE: ORBC_LAUNCH(c, k_opt_move, nb, 256, 0, a); break;
        case ORBC_OPT_FUSED: a.clear = 0; ORBC_LAUNCH(c, k_opt_fused, nb, 256, 0, a); if (mg_active(c)) S.cur_xn ^= 1; break;
        default: return fail(ORBC_ERR_ARG, "orbc_integrate: unknown kernel id %d", kernel);
        }
    }
    if (mg_active(c) && reduces) {
        // the partial sums of all ranks, added in rank order on every rank; the barrier is also the one behind the halo push
        ORBC_TRY(mg_share_ke(c)); ORBC_TRY(mg_barrier(c));
        ORBC_LAUNCH(c, k_sum_ke, 1, 32, 0, c->d_acc, mg_ke_slots(c), c->mg.world);
    } else if (mg_active(c) && (kernel == ORBC_NH_INITIAL_FUSED || kernel == ORBC_OPT_FUSED)) ORBC_TRY(mg_barrier(c));
    if (res) {
        res->ke = 0.0; res->n = (long)(c->sp[0].n + c->sp[1].n);
        if (reduces) {
            ORBC_CUDA(cudaMemcpyAsync(c->h_acc, c->d_acc, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            ORBC_CUDA(cudaStreamSynchronize(c->stream));
            res->ke = c->h_acc[0];
        }
    }
    return ORBC_OK;
}

int orbc_compute_temperature(orbc_ctx *c, double *T) { if (c) cudaSetDevice(c->device);
    if (!c || !T) return fail(ORBC_ERR_ARG, "null argument");
    ORBC_CUDA(cudaMemsetAsync(c->d_acc, 0, sizeof(double), c->stream));
    size_t n = 0;
    for (int sp = 0; sp < 2; ++sp) {
        Species &S = c->sp[sp];
        n += S.n;
        if (S.n) ORBC_LAUNCH(c, k_kinetic, blocks_for(owned_bound(c, sp), 256), 256, 0, S.X(), S.V(), c->d_range + 2 * sp, 1.0f, c->d_acc);
    }
    ORBC_CUDA(cudaMemcpyAsync(c->h_acc, c->d_acc, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ORBC_CUDA(cudaStreamSynchronize(c->stream));
    *T = c->h_acc[0] / (3.0 * (double)n);     // decomposed run: this rank's share of the sum (the shares of all ranks add up to T)
    return ORBC_OK;
}

int orbc_run_langevin(orbc_ctx *c, const orbc_step_params *p, int n_steps, int freq_voronoi, int freq_sort_ctrd) { if (c) cudaSetDevice(c->device);
    if (!c || !p || freq_voronoi <= 0) return fail(ORBC_ERR_ARG, "orbc_run_langevin: bad argument");
    orbc_step_params q = *p;
    q.noise_lipid = q.noise_protein = nullptr;
    for (int s = 0; s < n_steps; ++s, ++q.nstep) {
        if (q.nstep % freq_voronoi == 0) ORBC_TRY(do_rebuild(c, q.nstep, freq_sort_ctrd));
        // f and t are accumulators in the reference (cleared by the integrator, integrate_langevin.h:144).  Inside this loop they are
        // dead between the integrator and the next force evaluation, so only the first evaluation accumulates (onto whatever the
        // caller left there) and only the last integration clears: 64 B per particle and step less traffic.
        const bool first = s == 0 || c->pair_impl != 2, last = s + 1 == n_steps || c->pair_impl != 2;   // (the cross-check kernels always accumulate)
        ORBC_TRY(launch_pairwise(c, first));
        ORBC_TRY(launch_bonded(c));
        if (c->cv_on) ORBC_TRY(do_constrain_volume(c, c->cv_target, c->cv_strength));   // openrbc.cpp:229
        ORBC_TRY(do_integrate_langevin(c, &q, !last && (q.nstep + 1) % freq_voronoi == 0, last));
    }
    if (mg_active(c) && n_steps > 1)     // slots this rank owned at some step but not at the last one still hold dead values: clear everything
        for (int sp = 0; sp < 2; ++sp) if (c->sp[sp].n) {
            ORBC_LAUNCH(c, k_zero4, blocks_for(c->sp[sp].n, kBlock), kBlock, 0, c->sp[sp].f, c->sp[sp].n, (const int *)nullptr);
            ORBC_LAUNCH(c, k_zero4, blocks_for(c->sp[sp].n, kBlock), kBlock, 0, c->sp[sp].t, c->sp[sp].n, (const int *)nullptr);
        }
    return ORBC_OK;
}

int orbc_run_minimize(orbc_ctx *c, const orbc_step_params *p, int n_steps, int freq_sort_ctrd) { if (c) cudaSetDevice(c->device);
    if (!c || !p) return fail(ORBC_ERR_ARG, "orbc_run_minimize: bad argument");
    // openrbc.cpp:88-133.  param.nstep stays 0 for the whole minimisation (it is only incremented in the main loop, :244), so every
    // iteration Morton-sorts the centroids (voronoi.h:82: nstep % freq_sort_ctrd == 0).  clear_force is folded into the first force
    // kernel (plain stores instead of accumulation); post_torque + mover + bounce_back are one pass (k_opt_fused).
    orbc_step_params q = *p; q.nstep = 0;
    for (int s = 0; s < n_steps; ++s) {
        ORBC_TRY(do_rebuild(c, 0, freq_sort_ctrd));
        if (c->pair_impl != 2)                                   // the cross-check kernels always accumulate
            for (int sp = 0; sp < 2; ++sp) if (c->sp[sp].n) ORBC_LAUNCH(c, k_clear_force, blocks_for(c->sp[sp].n, 256), 256, 0, c->sp[sp].f, c->sp[sp].t, c->sp[sp].n);
        ORBC_TRY(launch_pairwise(c, false));
        ORBC_TRY(launch_bonded(c));
        {
            ProfScope ps(c, ORBC_PROF_INTEGRATE);
            for (int sp = 0; sp < 2; ++sp) {
                Species &S = c->sp[sp];
                if (!S.n) continue;
                IntegArgs a; fill_integ(a, c, sp, &q); a.clear = 0;
                c->nl_valid = false;
                ORBC_LAUNCH(c, k_opt_fused, blocks_for(owned_bound(c, sp), 256), 256, 0, a);
                if (mg_active(c)) S.cur_xn ^= 1;
            }
        }
        if (s + 1 == n_steps) ORBC_TRY(mg_barrier(c));           // otherwise the first barrier of the next rebuild covers the push
    }
    return ORBC_OK;
}

int orbc_run_nh(orbc_ctx *c, const orbc_step_params *p, int n_steps, int freq_voronoi, int freq_sort_ctrd, float *zeta_io, float *Q_io) { if (c) cudaSetDevice(c->device);
    if (!c || !p || !zeta_io || !Q_io || freq_voronoi <= 0) return fail(ORBC_ERR_ARG, "orbc_run_nh: bad argument");
    c->h_nh[0] = *zeta_io; c->h_nh[1] = *Q_io;
    ORBC_CUDA(cudaMemcpyAsync(c->d_nh, c->h_nh, 2 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    ORBC_CUDA(cudaMemsetAsync(c->d_acc, 0, sizeof(double), c->stream));
    orbc_step_params q = *p;
    const long n = (long)(c->sp[0].n + c->sp[1].n);
    const bool mg = mg_active(c);
    const int world = mg ? c->mg.world : 1;
    for (int s = 0; s < n_steps; ++s, ++q.nstep) {
        ++c->nl_moves;
        {
            IntegArgs a[2]; unsigned blocks[2] = {0, 0}; int m = 0;                  // both containers in one launch
            for (int sp = 0; sp < 2; ++sp) if (c->sp[sp].n) {
                fill_integ(a[m], c, sp, &q); a[m].zeta_dev = c->d_nh;
                blocks[m++] = blocks_for(owned_bound(c, sp), 256);
                if (mg) c->sp[sp].cur_xn ^= 1;
            }
            if (m == 2) ORBC_LAUNCH(c, k_nh_initial_fused2, blocks[0] + blocks[1], 256, 0, a[0], a[1], blocks[0]);
            else if (m == 1) ORBC_LAUNCH(c, k_nh_initial_fused, blocks[0], 256, 0, a[0]);
        }
        ORBC_TRY(nl_share(c));
        // decomposed: one barrier stands behind both the halo push of the drift and the exchange of the partial kinetic energies
        ORBC_TRY(mg_share_ke(c)); ORBC_TRY(mg_barrier(c));
        ORBC_LAUNCH(c, k_nh_zeta_update, 1, 32, 0, c->d_nh, c->d_acc, q.dt, q.kBT, n, mg ? mg_ke_slots(c) : (const double *)nullptr, world);
        if (q.nstep % freq_voronoi == 0) ORBC_TRY(do_rebuild(c, q.nstep, freq_sort_ctrd));
        ORBC_TRY(launch_pairwise(c));
        ORBC_TRY(launch_bonded(c));
        if (c->cv_on) ORBC_TRY(do_constrain_volume(c, c->cv_target, c->cv_strength));   // openrbc.cpp:229
        {
            IntegArgs a[2]; unsigned blocks[2] = {0, 0}; int m = 0;
            for (int sp = 0; sp < 2; ++sp) if (c->sp[sp].n) {
                fill_integ(a[m], c, sp, &q); a[m].zeta_dev = c->d_nh;
                blocks[m++] = blocks_for(owned_bound(c, sp), 256);
            }
            if (m == 2) ORBC_LAUNCH(c, k_nh_final_fused2, blocks[0] + blocks[1], 256, 0, a[0], a[1], blocks[0]);
            else if (m == 1) ORBC_LAUNCH(c, k_nh_final_fused, blocks[0], 256, 0, a[0]);
        }
        ORBC_TRY(mg_share_ke(c)); ORBC_TRY(mg_barrier(c));
        ORBC_LAUNCH(c, k_nh_zeta_update, 1, 32, 0, c->d_nh, c->d_acc, q.dt, q.kBT, n, mg ? mg_ke_slots(c) : (const double *)nullptr, world);
    }
    ORBC_CUDA(cudaMemcpyAsync(c->h_nh, c->d_nh, 2 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    ORBC_CUDA(cudaStreamSynchronize(c->stream));
    *zeta_io = c->h_nh[0]; *Q_io = c->h_nh[1];
    return check_flags(c);
}

// ---- decomposition over the GPUs of one box (multi.cuh) ------------------------------------------------------------------------
int orbc_mg_init(orbc_ctx *c, int rank, int world) { if (c) cudaSetDevice(c->device);
    if (!c || world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return fail(ORBC_ERR_ARG, "orbc_mg_init: rank %d of %d (at most %d ranks)", rank, world, kMaxWorld);
    if (c->mg.connected) return fail(ORBC_ERR_ARG, "orbc_mg_init: already connected");
    c->mg.on = true; c->mg.rank = rank; c->mg.world = world;
    return preload_kernels();
}

size_t orbc_mg_blob_bytes(void) { return sizeof(MgBlob); }

int orbc_mg_cell_range(int n_cells, int rank, int world, int *cell_begin, int *cell_end) {
    // the reference's own static partition of a range over its workers (util_numa.h:41-42)
    if (n_cells < 0 || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || !cell_begin || !cell_end) return fail(ORBC_ERR_ARG, "orbc_mg_cell_range: bad argument");
    *cell_begin = (int)((long long)rank * n_cells / world);
    *cell_end = rank == world - 1 ? n_cells : (int)((long long)(rank + 1) * n_cells / world);
    return ORBC_OK;
}

int orbc_mg_export(orbc_ctx *c, void *blob_out, size_t bytes) { if (c) cudaSetDevice(c->device);
    if (!c || !blob_out || bytes < sizeof(MgBlob)) return fail(ORBC_ERR_ARG, "orbc_mg_export: bad argument");
    if (!c->mg.on) return fail(ORBC_ERR_ARG, "orbc_mg_export: call orbc_mg_init first");
    Species &L = c->sp[0], &P = c->sp[1];
    if (!c->n_cells || !L.has_partition || (P.n && !P.has_partition))
        return fail(ORBC_ERR_ARG, "orbc_mg_export: upload the whole state and its Voronoi partition on every rank first");
    const int nc = c->n_cells, w = c->mg.world;
    Decomp &m = c->mg;
    const CellOwners own = cell_owners(c);
    m.cb = own.beg[m.rank]; m.ce = own.beg[m.rank + 1];
    if (m.flags && c->mg.connected) {
        // a second call after a fresh upload of the same system: the peer-visible allocations stay where they are
        for (int sp = 0; sp < 2; ++sp) {
            ORBC_CUDA(cudaMemsetAsync(m.cnt_all[sp], 0, sizeof(int) * (size_t)w * (nc + 1), c->stream));
            ORBC_CUDA(cudaMemsetAsync(m.cnt_prev[sp], 0, sizeof(int) * ((size_t)nc + 1), c->stream));
        }
        ORBC_CUDA(cudaMemsetAsync(m.pmask, 0, P.n + 4, c->stream));
        if (m.cen_buf[m.cen_par] != c->centroid) return fail(ORBC_ERR_STATE, "orbc_mg_export: centroid buffers were reallocated after orbc_mg_connect");
    } else {
        for (int sp = 0; sp < 2; ++sp) {
            const size_t n = c->sp[sp].n;
            m.own_cap[sp] = w == 1 ? n : std::min(n, n / w + n / (4 * (size_t)w) + (size_t)c->mg_own_slack);
            ORBC_TRY(dev_alloc(&m.cnt_all[sp], (size_t)w * (nc + 1))); ORBC_TRY(dev_alloc(&m.off_me[sp], (size_t)nc + 1)); ORBC_TRY(dev_alloc(&m.cnt_prev[sp], (size_t)nc + 1));
            ORBC_CUDA(cudaMemsetAsync(m.cnt_prev[sp], 0, sizeof(int) * ((size_t)nc + 1), c->stream));
            ORBC_CUDA(cudaMemsetAsync(m.cnt_all[sp], 0, sizeof(int) * (size_t)w * (nc + 1), c->stream));
        }
        // barrier epochs [0, 8) and two sets of per-rank displacement bounds [8, 16), [16, 24) (hit lists, k_nl_share)
        ORBC_TRY(dev_alloc(&m.flags, 3 * kMaxWorld)); ORBC_CUDA(cudaMemsetAsync(m.flags, 0, sizeof(unsigned) * 3 * kMaxWorld, c->stream));
        ORBC_TRY(dev_alloc(&m.ke_all, 2 * kMaxWorld)); ORBC_CUDA(cudaMemsetAsync(m.ke_all, 0, sizeof(double) * 2 * kMaxWorld, c->stream));
        ORBC_TRY(dev_alloc(&m.vol_all, 2 * kMaxWorld)); ORBC_CUDA(cudaMemsetAsync(m.vol_all, 0, sizeof(double) * 2 * kMaxWorld, c->stream));
        ORBC_TRY(dev_alloc(&m.cv_ptype, nc)); ORBC_CUDA(cudaMemsetAsync(m.cv_ptype, 0, sizeof(int) * nc, c->stream));
        ORBC_TRY(dev_alloc(&m.dest_mask, nc)); ORBC_CUDA(cudaMemsetAsync(m.dest_mask, 0, nc, c->stream));
        ORBC_TRY(dev_alloc(&m.pmask, P.n + 4)); ORBC_CUDA(cudaMemsetAsync(m.pmask, 0, P.n + 4, c->stream));
        ORBC_TRY(dev_alloc(&m.need, nc)); ORBC_CUDA(cudaMemsetAsync(m.need, 0, sizeof(int) * nc, c->stream));
        ORBC_TRY(dev_alloc(&m.keep, L.n + 1));
        if (c->nl_on == 2 || (c->nl_on && w <= kNlAutoWorld)) ORBC_TRY(nl_ensure(c));   // the hit lists: allocated now, never while peers wait in a barrier
        // work list of the bonds with an owned atom (clipped and flagged at the capacity)
        m.my_bonds_cap = (int)std::min(c->n_bonds, c->n_bonds / w + c->n_bonds / (2 * (size_t)w) + (size_t)c->mg_own_slack);
        ORBC_TRY(dev_alloc(&m.my_bonds, (size_t)m.my_bonds_cap + 1));
        m.cen_buf[0] = c->centroid; m.cen_buf[1] = c->centroid_tmp; m.cen_par = 0; m.epoch = 0;
        // scratch that the single-GPU path allocates on first use: allocate it now, no cudaMalloc while peers spin in a barrier
        ORBC_CUDA(cudaMemsetAsync(m.off_me[0], 0, 2 * sizeof(int), c->stream));
        ORBC_TRY(scan_exclusive(c, m.off_me[0], 1));             // allocates the scan's descriptors
        const size_t hist_n = (size_t)256 * ((nc + kRadixTile - 1) / kRadixTile) + 1;
        if (c->radix_hist_cap < hist_n) { ORBC_TRY(dev_alloc(&c->radix_hist, hist_n)); c->radix_hist_cap = hist_n; }
        if (c->porder_cap < m.own_cap[1] + 1) { ORBC_TRY(dev_alloc(&c->porder, m.own_cap[1] + 1)); c->porder_cap = m.own_cap[1] + 1; }
    }
    // owned ranges, halo masks and the cells this rank reads, from the uploaded (complete, everywhere identical) state
    if (w > 1) {
        ORBC_LAUNCH(c, k_set_range, 1, 32, 0, RangeArgs{c->sp[0].cell_start, c->d_range, (int)m.own_cap[0]}, RangeArgs{c->sp[1].cell_start, c->d_range + 2, (int)m.own_cap[1]}, 2, m.cb, m.ce, c->d_flags);
        ORBC_TRY(build_index(c, true));
        ORBC_CUDA(cudaMemsetAsync(m.my_bonds, 0, sizeof(int), c->stream));
        if (P.n && c->n_bonds)
            ORBC_LAUNCH(c, k_bond_mask, blocks_for(c->n_bonds, kBlock), kBlock, 0, c->bonds, c->n_bonds, c->tag2idx, c->d_range, P.cell_start, own, (unsigned *)m.pmask,
                        m.my_bonds, m.my_bonds_cap, c->d_flags);
        c->porder_valid = false;
    }
    // a repeated export (fresh upload of the same system on connected ranks) is a collective call: no rank may start storing
    // counts, halo copies or migrating particles into a peer that is still uploading or clearing its tables
    if (m.connected) ORBC_TRY(mg_barrier(c));
    if (m.connected && w > 1 && (m.partial[0] || m.partial[1])) {
        // every rank uploaded only its own rows (orbc_upload_range): the halo copies come from their owners, over NVLink
        for (int sp = 0; sp < 2; ++sp) {
            Species &S = c->sp[sp];
            if (!S.n) continue;
            HaloArgs h;
            h.range2 = c->d_range + 2 * sp; h.cell_mask = m.dest_mask; h.pmask = sp == ORBC_PROTEIN ? m.pmask : (const unsigned char *)nullptr;
            h.cellid = S.C(); h.x = S.X(); h.nn = S.N();
            for (int r = 0; r < kMaxWorld; ++r) { h.d.x[r] = m.peers.x[sp][S.cur_xn][r]; h.d.nn[r] = m.peers.nn[sp][S.cur_xn][r]; }
            const unsigned nb = blocks_for(owned_bound(c, sp), kBlock);
            ORBC_LAUNCH(c, k_halo_push, nb, kBlock, 0, h, h, nb);
        }
        ORBC_TRY(mg_barrier(c));
    }
    ORBC_CUDA(cudaStreamSynchronize(c->stream));
    MgBlob *b = (MgBlob *)blob_out;
    memset(b, 0, sizeof(*b));
    b->magic = kMgMagic; b->pid = (int)getpid(); b->device = c->device; b->rank = m.rank; b->world = w; b->n_cells = nc; b->n_l = L.n; b->n_p = P.n;
    void *list[kMgShared]; mg_shared_list(c, list);
    for (int k = 0; k < kMgShared; ++k) {
        b->e[k].raw = (unsigned long long)(uintptr_t)list[k];
        if (list[k] && w > 1) ORBC_CUDA(cudaIpcGetMemHandle(&b->e[k].handle, list[k]));
    }
    return check_flags(c);
}

int orbc_mg_connect(orbc_ctx *c, const void *blobs, size_t bytes_each) { if (c) cudaSetDevice(c->device);
    if (!c || !blobs || bytes_each < sizeof(MgBlob)) return fail(ORBC_ERR_ARG, "orbc_mg_connect: bad argument");
    Decomp &m = c->mg;
    if (!m.on || !m.flags) return fail(ORBC_ERR_ARG, "orbc_mg_connect: call orbc_mg_init and orbc_mg_export first");
    ORBC_CUDA(cudaSetDevice(c->device));
    for (int r = 0; r < m.world; ++r) {
        const MgBlob *b = (const MgBlob *)((const char *)blobs + (size_t)r * bytes_each);
        if (b->magic != kMgMagic || b->rank != r || b->world != m.world) return fail(ORBC_ERR_ARG, "orbc_mg_connect: blob %d is not rank %d of %d", r, r, m.world);
        if (b->n_cells != c->n_cells || b->n_l != c->sp[0].n || b->n_p != c->sp[1].n) return fail(ORBC_ERR_ARG, "orbc_mg_connect: rank %d holds a different system", r);
        void *ptr[kMgShared];
        if (b->pid == (int)getpid() && b->device != c->device) { // same process, another device: peer access, once per pair of devices
            cudaError_t e = cudaDeviceEnablePeerAccess(b->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ORBC_CUDA(e);
            cudaGetLastError();                                  // (another context of this process may have enabled it already)
        }
        for (int k = 0; k < kMgShared; ++k) {
            ptr[k] = nullptr;
            if (!b->e[k].raw) continue;
            if (b->pid == (int)getpid()) {                       // same process (several ranks driven by one host program): plain pointers
                ptr[k] = (void *)(uintptr_t)b->e[k].raw;
            } else {                                             // one process per GPU: CUDA IPC mapping of the peer's allocation
                ORBC_CUDA(cudaIpcOpenMemHandle(&ptr[k], b->e[k].handle, cudaIpcMemLazyEnablePeerAccess));
                m.opened.push_back(ptr[k]);
            }
        }
        mg_fill_peers(c, r, ptr);
    }
    m.connected = true;
    return ORBC_OK;
}

int orbc_mg_range(orbc_ctx *c, int sp, size_t *begin, size_t *end) { if (c) cudaSetDevice(c->device);
    if (!c || sp < 0 || sp > 1 || !begin || !end) return fail(ORBC_ERR_ARG, "orbc_mg_range: bad argument");
    int r[2];
    ORBC_CUDA(cudaMemcpyAsync(r, c->d_range + 2 * sp, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ORBC_CUDA(cudaStreamSynchronize(c->stream));
    *begin = (size_t)r[0]; *end = (size_t)r[1];
    return check_flags(c);
}

int orbc_size(orbc_ctx *c, int sp, size_t *n) { if (c) cudaSetDevice(c->device); if (!c || sp < 0 || sp > 1 || !n) return fail(ORBC_ERR_ARG, "bad argument"); *n = c->sp[sp].n; return ORBC_OK; }
int orbc_n_cells(orbc_ctx *c, int *n) { if (c) cudaSetDevice(c->device); if (!c || !n) return fail(ORBC_ERR_ARG, "bad argument"); *n = c->n_cells; return ORBC_OK; }

int orbc_download(orbc_ctx *c, int sp, size_t stride, float *x, float *v, float *n_, float *o, float *f, float *t,
                  int *affiliation, int *type, int *tag, size_t *n_out) { if (c) cudaSetDevice(c->device);
    if (!c || sp < 0 || sp > 1 || stride < 3) return fail(ORBC_ERR_ARG, "orbc_download: bad argument");
    Species &S = c->sp[sp];
    if (n_out) *n_out = S.n;
    if (!S.n) return check_flags(c);
    // a rank of a decomposed run holds current values only in the slots it owns: it fills exactly those rows of the host arrays
    // (the ranks of one host program write disjoint rows of the same containers)
    size_t first = 0, count = S.n;
    if (mg_active(c)) {
        int r[2];
        ORBC_CUDA(cudaMemcpyAsync(r, c->d_range + 2 * sp, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        ORBC_CUDA(cudaStreamSynchronize(c->stream));
        first = (size_t)r[0]; count = (size_t)(r[1] - r[0]);
    }
    if (!count) return check_flags(c);
    ORBC_TRY(ensure_copy_stream(c));
    const float4 *src3[6] = {S.X(), S.V(), S.N(), S.O(), S.f, S.t};
    float *dst3[6] = {x, v, n_, o, f, t};
    const float4 *srcw[2] = {S.X(), S.N()};
    int *dstw[2] = {type, tag};
    size_t need = 0;
    for (int k = 0; k < 6; ++k) if (dst3[k]) need += count * stride;
    for (int k = 0; k < 2; ++k) if (dstw[k]) need += count;
    ORBC_TRY(ensure_stage(c, need + 16));
    const unsigned nb = blocks_for(count, kBlock);
    float *p = c->stage;
    int ev = 0;
    // unpack on the compute stream, copy out on the copy stream as soon as each array is ready; one wait at the end
    for (int k = 0; k < 6; ++k) if (dst3[k]) {
        ORBC_LAUNCH(c, k_unpack3, nb, kBlock, 0, src3[k] + first, count, stride, p);
        ORBC_CUDA(cudaEventRecord(c->xfer_ev[ev], c->stream));
        ORBC_CUDA(cudaStreamWaitEvent(c->copy_stream, c->xfer_ev[ev], 0));
        ORBC_CUDA(cudaMemcpyAsync(dst3[k] + first * stride, p, sizeof(float) * count * stride, cudaMemcpyDeviceToHost, c->copy_stream));
        p += count * stride; ev = (ev + 1) & 7;
    }
    for (int k = 0; k < 2; ++k) if (dstw[k]) {
        ORBC_LAUNCH(c, k_unpack_w, nb, kBlock, 0, srcw[k] + first, count, (int *)p);
        ORBC_CUDA(cudaEventRecord(c->xfer_ev[ev], c->stream));
        ORBC_CUDA(cudaStreamWaitEvent(c->copy_stream, c->xfer_ev[ev], 0));
        ORBC_CUDA(cudaMemcpyAsync(dstw[k] + first, p, sizeof(int) * count, cudaMemcpyDeviceToHost, c->copy_stream));
        p += count; ev = (ev + 1) & 7;
    }
    if (affiliation) {
        ORBC_CUDA(cudaEventRecord(c->xfer_ev[ev], c->stream));
        ORBC_CUDA(cudaStreamWaitEvent(c->copy_stream, c->xfer_ev[ev], 0));
        ORBC_CUDA(cudaMemcpyAsync(affiliation + first, S.C() + first, sizeof(int) * count, cudaMemcpyDeviceToHost, c->copy_stream));
    }
    ORBC_CUDA(cudaStreamSynchronize(c->copy_stream));
    // the staging area must not be reused by the compute stream before the copies have left it: they have (synchronised above)
    return check_flags(c);
}

// ---- save_frame -----------------------------------------------------------------------------------------------------------------------
static FrameLayout frame_layout(const orbc_ctx *c, int nstep, int dump_field, int tag_base, size_t *total) {
    FrameLayout L; memset(&L, 0, sizeof(L));
    L.n_l = c->sp[0].n; L.n_p = c->sp[1].n; L.nstep = nstep; L.tag_base = tag_base;
    const size_t n = L.n_l + L.n_p;
    size_t off = 8 + 4 + 8 + 8 + 8;                              // FRAMEBEG nstep NATOM n IDENTITY
    L.off_id = off; off += 8 * n;
    auto section = [&](bool on, size_t bytes_each) -> size_t { if (!on) return 0; off += 8; const size_t o = off; off += bytes_each * n; return o; };
    L.off_x = section(dump_field & 1, 12);                       // DumpField::position  (runtime_parameter.h:30-36; order of trajectory.h:77-101)
    L.off_v = section(dump_field & 8, 12);                       // velocity
    L.off_n = section(dump_field & 2, 12);                       // rotation
    L.off_aff = section(dump_field & 4, 4);                      // voronoi
    L.off_f = section(dump_field & 16, 12);                      // force
    L.off_end = off; off += 8;
    *total = off;
    return L;
}

int orbc_frame_bytes(orbc_ctx *c, int dump_field, size_t *bytes) {
    if (!c || !bytes) return fail(ORBC_ERR_ARG, "orbc_frame_bytes: bad argument");
    frame_layout(c, 0, dump_field, 1, bytes);
    return ORBC_OK;
}

// pack the frame into device buffer `slot` on the context's stream
static int frame_pack(orbc_ctx *c, int slot, int nstep, int dump_field, int tag_base, size_t *total) {
    FrameLayout L = frame_layout(c, nstep, dump_field, tag_base, total);
    Species &S0 = c->sp[0], &S1 = c->sp[1];
    if ((dump_field & 4) && ((S0.n && !S0.has_partition) || (S1.n && !S1.has_partition))) return fail(ORBC_ERR_ARG, "save_frame: the VORONOI section needs a partition");
    if (c->frame_dev_cap[slot] < *total) { ORBC_TRY(dev_alloc(&c->frame_dev[slot], *total + (*total >> 4))); c->frame_dev_cap[slot] = *total + (*total >> 4); }
    L.l0 = 0; L.l1 = (int)S0.n; L.p0 = 0; L.p1 = (int)S1.n;
    if (mg_active(c)) {
        // a rank's image holds the titles and the slots it owns; the other slots are zero (the images of all ranks OR together)
        int r[4];
        ORBC_CUDA(cudaMemcpyAsync(r, c->d_range, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        ORBC_CUDA(cudaStreamSynchronize(c->stream));
        L.l0 = r[0]; L.l1 = r[1]; L.p0 = r[2]; L.p1 = r[3];
        ORBC_CUDA(cudaMemsetAsync(c->frame_dev[slot], 0, *total, c->stream));
    }
    ORBC_LAUNCH(c, k_frame_pack, blocks_for(std::max<size_t>(1, L.n_l + L.n_p), 256), 256, 0, L, c->frame_dev[slot],
                S0.X(), S0.V(), S0.N(), S0.f, S0.C(), S1.X(), S1.V(), S1.N(), S1.f, S1.C());
    return ORBC_OK;
}

int orbc_save_frame(orbc_ctx *c, int nstep, int dump_field, int lipid_tag_base, void *dst, size_t cap, size_t *bytes) { if (c) cudaSetDevice(c->device);
    if (!c || !dst) return fail(ORBC_ERR_ARG, "orbc_save_frame: bad argument");
    size_t total = 0;
    frame_layout(c, nstep, dump_field, lipid_tag_base, &total);
    if (bytes) *bytes = total;
    if (cap < total) return fail(ORBC_ERR_ARG, "orbc_save_frame: the frame needs %zu bytes, the buffer holds %zu", total, cap);
    ORBC_TRY(frame_pack(c, 0, nstep, dump_field, lipid_tag_base, &total));
    ORBC_CUDA(cudaMemcpyAsync(dst, c->frame_dev[0], total, cudaMemcpyDeviceToHost, c->stream));
    ORBC_CUDA(cudaStreamSynchronize(c->stream));
    return check_flags(c);
}

int orbc_save_frame_begin(orbc_ctx *c, int nstep, int dump_field, int lipid_tag_base) { if (c) cudaSetDevice(c->device);
    if (!c) return fail(ORBC_ERR_ARG, "null ctx");
    if (c->frame_pending >= 2) return fail(ORBC_ERR_ARG, "orbc_save_frame_begin: two frames are already in flight (call orbc_save_frame_end)");
    ORBC_TRY(ensure_copy_stream(c));
    const int slot = (c->frame_head + c->frame_pending) & 1;
    size_t total = 0;
    ORBC_TRY(frame_pack(c, slot, nstep, dump_field, lipid_tag_base, &total));
    if (c->frame_host_cap[slot] < total) {
        if (c->frame_host[slot]) { ORBC_CUDA(cudaFreeHost(c->frame_host[slot])); c->frame_host[slot] = nullptr; }
        ORBC_CUDA(cudaMallocHost((void **)&c->frame_host[slot], total + (total >> 4)));
        c->frame_host_cap[slot] = total + (total >> 4);
    }
    // the copy runs on its own stream behind the pack; the next kernels of the run do not wait for it
    ORBC_CUDA(cudaEventRecord(c->frame_packed[slot], c->stream));
    ORBC_CUDA(cudaStreamWaitEvent(c->copy_stream, c->frame_packed[slot], 0));
    ORBC_CUDA(cudaMemcpyAsync(c->frame_host[slot], c->frame_dev[slot], total, cudaMemcpyDeviceToHost, c->copy_stream));
    ORBC_CUDA(cudaEventRecord(c->frame_copied[slot], c->copy_stream));
    c->frame_bytes[slot] = total;
    ++c->frame_pending;
    return ORBC_OK;
}

int orbc_save_frame_end(orbc_ctx *c, const void **data, size_t *bytes) { if (c) cudaSetDevice(c->device);
    if (!c || !data || !bytes) return fail(ORBC_ERR_ARG, "orbc_save_frame_end: bad argument");
    if (!c->frame_pending) return fail(ORBC_ERR_ARG, "orbc_save_frame_end: no frame in flight");
    const int slot = c->frame_head;
    ORBC_CUDA(cudaEventSynchronize(c->frame_copied[slot]));
    *data = c->frame_host[slot]; *bytes = c->frame_bytes[slot];
    c->frame_head ^= 1; --c->frame_pending;
    return ORBC_OK;
}

int orbc_debug_dump(orbc_ctx *c, int what, void *dst, size_t bytes) { if (c) cudaSetDevice(c->device);
    if (!c || !dst) return fail(ORBC_ERR_ARG, "bad argument");
    const int nc = c->n_cells;
    const void *src = nullptr; size_t need = 0;
    switch (what) {
    case ORBC_DUMP_CENTROIDS: {
        need = sizeof(float) * 3 * nc;
        if (bytes < need) return fail(ORBC_ERR_ARG, "dump buffer too small");
        ORBC_TRY(ensure_stage(c, (size_t)3 * nc));
        ORBC_LAUNCH(c, k_unpack3, blocks_for(nc, kBlock), kBlock, 0, c->centroid, (size_t)nc, (size_t)3, c->stage);
        src = c->stage; break; }
    case ORBC_DUMP_CELL_START_L: src = c->sp[0].cell_start; need = sizeof(int) * ((size_t)nc + 1); break;
    case ORBC_DUMP_CELL_START_P: src = c->sp[1].cell_start; need = sizeof(int) * ((size_t)nc + 1); break;
    case ORBC_DUMP_CELLS_L: src = c->sp[0].cells; need = sizeof(int) * c->sp[0].n; break;
    case ORBC_DUMP_CELLS_P: src = c->sp[1].cells; need = sizeof(int) * c->sp[1].n; break;
    case ORBC_DUMP_AFF_L: src = c->sp[0].aff; need = sizeof(int) * c->sp[0].n; break;
    case ORBC_DUMP_AFF_P: src = c->sp[1].aff; need = sizeof(int) * c->sp[1].n; break;
    case ORBC_DUMP_MORTON_KEYS:
        need = sizeof(uint32_t) * nc;
        ORBC_LAUNCH(c, k_morton_keys_only, blocks_for(nc, kBlock), kBlock, 0, c->centroid, nc, c->keys_tmp);
        src = c->keys_tmp; break;
    case ORBC_DUMP_MORTON_PERM: src = c->perm; need = sizeof(int) * nc; break;
    case ORBC_DUMP_STENCIL_COUNTS: {
        need = sizeof(int) * 3 * nc;
        if (bytes < need) return fail(ORBC_ERR_ARG, "dump buffer too small");
        std::vector<int> packed(nc);
        ORBC_CUDA(cudaMemcpyAsync(packed.data(), c->stencil_cnt, sizeof(int) * nc, cudaMemcpyDeviceToHost, c->stream));
        ORBC_CUDA(cudaStreamSynchronize(c->stream));
        int *o = (int *)dst;
        for (int i = 0; i < nc; ++i) { o[3 * i] = packed[i] & 255; o[3 * i + 1] = (packed[i] >> 8) & 255; o[3 * i + 2] = packed[i] >> 16; }
        return check_flags(c); }
    case ORBC_DUMP_STENCIL: src = c->stencil; need = sizeof(int) * (size_t)nc * kStencilStride; break;
    case ORBC_DUMP_TAG2IDX: src = c->tag2idx; need = sizeof(int) * c->tag2idx_size; break;
    case ORBC_DUMP_COUNTERS: src = c->d_counters; need = 8 * sizeof(unsigned long long); break;
    case ORBC_DUMP_NL_STATS: {
        need = 4 * sizeof(unsigned);
        if (bytes < need) return fail(ORBC_ERR_ARG, "dump buffer too small");
        unsigned out[4] = {0, 0, 0, 0};
        if (c->nl_state) {
            NlState h;
            ORBC_CUDA(cudaMemcpyAsync(&h, c->nl_state, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
            ORBC_CUDA(cudaStreamSynchronize(c->stream));
            out[0] = h.builds; out[1] = h.reuses; out[2] = (unsigned)h.overflow; out[3] = h.searches;
        }
        memcpy(dst, out, need);
        return check_flags(c); }
    default: return fail(ORBC_ERR_ARG, "unknown dump id %d", what);
    }
    if (bytes < need) return fail(ORBC_ERR_ARG, "dump buffer too small: %zu < %zu", bytes, need);
    if (need) ORBC_CUDA(cudaMemcpyAsync(dst, src, need, cudaMemcpyDeviceToHost, c->stream));
    ORBC_CUDA(cudaStreamSynchronize(c->stream));
    return check_flags(c);
}

int orbc_debug_noise(orbc_ctx *c, uint64_t seed, int nstep, int species, size_t n, float *dst) { if (c) cudaSetDevice(c->device);
    if (!c || !dst) return fail(ORBC_ERR_ARG, "bad argument");
    ORBC_TRY(ensure_stage(c, 3 * n));
    ORBC_LAUNCH(c, k_noise, blocks_for(n, kBlock), kBlock, 0, seed, (uint32_t)nstep, (uint32_t)species, n, c->stage);
    ORBC_CUDA(cudaMemcpyAsync(dst, c->stage, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    ORBC_CUDA(cudaStreamSynchronize(c->stream));
    return ORBC_OK;
}

int orbc_event_record(orbc_ctx *c, int slot) { if (c) cudaSetDevice(c->device);
    if (!c || slot < 0 || slot >= 16) return fail(ORBC_ERR_ARG, "bad event slot");
    ORBC_CUDA(cudaEventRecord(c->ev[slot], c->stream));
    return ORBC_OK;
}
int orbc_event_elapsed_ms(orbc_ctx *c, int a, int b, float *ms) { if (c) cudaSetDevice(c->device);
    if (!c || a < 0 || a >= 16 || b < 0 || b >= 16 || !ms) return fail(ORBC_ERR_ARG, "bad event slot");
    ORBC_CUDA(cudaEventSynchronize(c->ev[b]));
    ORBC_CUDA(cudaEventElapsedTime(ms, c->ev[a], c->ev[b]));
    return ORBC_OK;
}
int orbc_profile_enable(orbc_ctx *c, int on) { if (c) cudaSetDevice(c->device);
    if (!c) return fail(ORBC_ERR_ARG, "null ctx");
    ORBC_CUDA(cudaStreamSynchronize(c->stream));
    c->prof_on = on != 0;
    for (auto &u : c->prof_used) u = 0;
    return ORBC_OK;
}
int orbc_profile_read(orbc_ctx *c, int cls, double *total_ms, unsigned long long *count) { if (c) cudaSetDevice(c->device);
    if (!c || cls < 0 || cls >= ORBC_PROF_N || !total_ms || !count) return fail(ORBC_ERR_ARG, "orbc_profile_read: bad argument");
    ORBC_CUDA(cudaStreamSynchronize(c->stream));
    double sum = 0.0; const size_t pairs = c->prof_used[cls] / 2;
    for (size_t k = 0; k < pairs; ++k) { float ms = 0.f; ORBC_CUDA(cudaEventElapsedTime(&ms, c->prof_ev[cls][2 * k], c->prof_ev[cls][2 * k + 1])); sum += ms; }
    *total_ms = sum; *count = pairs;
    c->prof_used[cls] = 0;
    return ORBC_OK;
}
int orbc_profile_kernels(orbc_ctx *c, int on) {
    if (c) cudaSetDevice(c->device);
    if (!c) return fail(ORBC_ERR_ARG, "null ctx");
    ORBC_CUDA(cudaStreamSynchronize(c->stream));
    c->kprof_on = on != 0; c->kprof_used = 0;
    return ORBC_OK;
}
int orbc_profile_kernels_report(orbc_ctx *c, char *text, size_t bytes) {
    if (c) cudaSetDevice(c->device);
    if (!c || !text || !bytes) return fail(ORBC_ERR_ARG, "orbc_profile_kernels_report: bad argument");
    ORBC_CUDA(cudaStreamSynchronize(c->stream));
    struct Row { const char *name; double us; unsigned long long n; };
    std::vector<Row> rows;
    for (size_t k = 0; k + 1 < c->kprof_used; k += 2) {
        float ms = 0.f; ORBC_CUDA(cudaEventElapsedTime(&ms, c->kprof_ev[k], c->kprof_ev[k + 1]));
        const char *nm = c->kprof_name[k / 2];
        size_t r = 0; while (r < rows.size() && strcmp(rows[r].name, nm)) ++r;
        if (r == rows.size()) rows.push_back({nm, 0.0, 0});
        rows[r].us += 1e3 * ms; rows[r].n++;
    }
    std::sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) { return a.us > b.us; });
    size_t off = 0; text[0] = 0;
    for (const Row &r : rows) {
        const int w = snprintf(text + off, bytes - off, "%s %llu %.1f\n", r.name, r.n, r.us);
        if (w < 0 || (size_t)w >= bytes - off) break;
        off += (size_t)w;
    }
    c->kprof_used = 0;
    return ORBC_OK;
}
int orbc_launch_count(orbc_ctx *c, unsigned long long *n) { if (c) cudaSetDevice(c->device); if (!c || !n) return fail(ORBC_ERR_ARG, "bad argument"); *n = c->launches; return ORBC_OK; }

} // extern "C"
