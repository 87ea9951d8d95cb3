// pair_tile.cuh — k_pair_ll_t: the lipid-lipid forces, one WARP per Voronoi cell over a shared-memory tile.
//
// Same candidate set, same hit test and same per-pair arithmetic as k_pair_ll_r / k_pair_lipid (compute_pairwise_fused.h:238-320 with
// lipid_lipid::rmax = 6, :92; the guards r2 < cutsq && r2 > 1e-5 of :109,134; pairwise_kernel.h:30-68), one-sided: every lipid gathers
// its own force.  What changes is where the operands live and how the lanes are used:
//
//   tile      Particles are stored sorted by cell and cells are numbered in Morton order, so the members of the r<6 stencil cells
//             of a cell are a few contiguous slot RUNS (k_lipid_runs).  The warp pulls the runs' x and n into shared memory with
//             1-D bulk asynchronous copies (cp.async.bulk -> UBLKCP, completion on an mbarrier): ~113 candidates x 32 B per cell,
//             read once from L2/HBM instead of once per lane.
//   phase 1   TEST.  Lanes hold the candidates (up to 8 per lane, in registers, relative to the tile's origin); the cell's own
//             lipids are broadcast one after the other from the tile.  The test is a PREFILTER in the form
//             |xj|^2/2 - xi.xj < (cutsq + margin)/2 - |xi|^2/2   (3 FFMA + 1 compare per pair instead of 3 FADD + FMUL + 2 FFMA + window
//             test); the margin covers its rounding (tile-relative coordinates keep it ~1e-4), never the other way round.  Hits
//             are recorded as ballot bit masks, mask[i][word]: one VOTE + one store per 32 pairs, no queues, no compaction.
//   phase 2   EVALUATE.  The hits of the whole cell, in (i, candidate) order, are cut into 32 equal chunks, one per lane: every
//             lane evaluates the same number of pairs (+-1) whatever the cell looks like.  The exact fp32 test of the reference is
//             repeated on the true coordinates (d = xi - xj exactly as the reference rounds it), so hits, and with them the
//             forces, do not depend on the prefilter.  A lane's chunk covers one to three lipids i; the partial sums of every
//             (lane, i) segment go to a slot of a small staging array.
//   phase 3   lane = lipid: adds the partial sums of its segments in lane order — a fixed order that depends on nothing but the
//             cell's own hits, so results are bit-identical from run to run and across any number of GPUs — and stores f, t.
//
// Cells whose candidates do not fit the tile (kTileCap) raise a flag in k_lipid_runs; that step then runs on k_pair_ll_r (both
// kernels are launched, the one that is not needed returns at once).
#pragma once
#include "common.cuh"
#include "pair.cuh"
#include "pair_queue.cuh"

namespace orbc {

constexpr int kTileCap = 256;                 // candidates per tile (full RBC: mean 113, max 169)
constexpr int kTileWords = kTileCap / 32;     // ballot words per lipid
constexpr int kTileWarps = 4;                 // warps per block, each with its own tile
constexpr int kTileCells = 8;                 // consecutive cells per warp
constexpr int kTileBytes = kTileCap * 32 + 32 * kTileWords * 4 + 64 * 32 + 16;   // x + n tile, masks, segment staging, mbarrier

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done, polls = 0;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && ++polls > (1u << 24)) __trap();             // a copy that never lands is a bug: fail the launch instead of hanging the GPU
    } while (!done);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one pair on the gathering side, operands from the tile; the reference's own guards decide (compute_pairwise_fused.h:109,134)
__device__ __forceinline__ void ll_eval_tile(const LLConst &k, unsigned lo_bits, unsigned span, F3 xi, F3 mi, float4 xj, float4 nj,
                                             float &fx, float &fy, float &fz, float &tx, float &ty, float &tz, float &sB) {
    const float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
    const float r2 = dx * dx + dy * dy + dz * dz;
    if (!(__float_as_uint(r2) - lo_bits < span)) return;
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

// `kc`: the lipid-lipid constants as a kernel PARAMETER (constant bank operands of the FFMAs; read from c_ff they were
// rematerialised with loads and multiplies in every evaluation)
__global__ void __launch_bounds__(kTileWarps * 32, 5) k_pair_ll_t(PairArgs a, const LLConst kc, const int2 *__restrict__ lruns, const int *__restrict__ lrun_info,
                                                                   const int *__restrict__ tile_overflow) {
    if (*tile_overflow) return;                                  // a cell does not fit the tile: k_pair_ll_r takes this step
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned char *base = smem_raw + wib * kTileBytes;
    float4 *const tile_x = reinterpret_cast<float4 *>(base);
    float4 *const tile_n = tile_x + kTileCap;
    unsigned *const mask = reinterpret_cast<unsigned *>(tile_n + kTileCap);      // [32][kTileWords]
    float4 *const seg = reinterpret_cast<float4 *>(mask + 32 * kTileWords);      // [64][2]
    const unsigned bar = smem_u32(seg + 128);
    const unsigned tile_x_s = smem_u32(tile_x), tile_n_s = smem_u32(tile_n);
    constexpr unsigned full = 0xffffffffu;
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();

    const unsigned lo_bits = __float_as_uint(1e-5f) + 1u, span = __float_as_uint(kc.cutsq) - lo_bits;
    const int l0 = a.range[0], l1 = a.range[1];
    const int gw = blockIdx.x * kTileWarps + wib;
    const int c_beg = a.cb + gw * kTileCells, c_end = min(c_beg + kTileCells, a.ce);
    unsigned parity = 0;

    for (int c = c_beg; c < c_end; ++c) {
        const int s0 = __ldg(a.cs_l + c), n1 = __ldg(a.cs_l + c + 1) - s0;
        if (n1 <= 0) continue;
        const int info = __ldg(lrun_info + c);
        const int nr = info & 63, ntot = (info >> 6) & 8191, own_off = (int)((unsigned)info >> 19);
        // ---- tile: one bulk copy of x and one of n per run ------------------------------------------------------------------------------
        int2 run = make_int2(0, 0);
        if (lane < nr) run = __ldg(lruns + (size_t)c * kRunStride + lane);
        int off = run.y;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(full, off, d); if (lane >= d) off += t; }
        off -= run.y;                                            // first tile slot of this lane's run
        fence_async_smem();                                      // this lane's reads of the previous tile, before the copies overwrite it
        __syncwarp();
        if (lane == 0) mbar_expect_tx(bar, (unsigned)ntot * 32u);
        __syncwarp();
        if (lane < nr) {
            bulk_g2s(tile_x_s + (unsigned)off * 16u, a.xl + run.x, (unsigned)run.y * 16u, bar);
            bulk_g2s(tile_n_s + (unsigned)off * 16u, a.nl + run.x, (unsigned)run.y * 16u, bar);
        }
        mbar_wait(bar, parity); parity ^= 1u;

        const int U = (ntot + 31) >> 5;                          // ballot words in use (<= kTileWords)
        const float4 org = tile_x[own_off];                      // the tile's origin: the cell's first lipid

        for (int ib = 0; ib < n1; ib += 32) {                    // the cell's own lipids, 32 at a time (one pass unless the cell is huge)
            const int nb = min(32, n1 - ib);
            // ---- candidates into registers, relative to the origin ------------------------------------------------------------------------
            float cx[kTileWords], cy[kTileWords], cz[kTileWords], ch[kTileWords];
            float hmax = 0.f;
            #pragma unroll
            for (int u = 0; u < kTileWords; ++u) {
                cx[u] = cy[u] = cz[u] = 0.f; ch[u] = 3.0e38f;     // an empty slot never passes the test
                if (u < U) {
                    const int idx = u * 32 + lane;
                    const float4 p = tile_x[min(idx, ntot - 1)];
                    const float rx = p.x - org.x, ry = p.y - org.y, rz = p.z - org.z;
                    const float h = 0.5f * (rx * rx + ry * ry + rz * rz);
                    if (idx < ntot) { cx[u] = rx; cy[u] = ry; cz[u] = rz; ch[u] = h; hmax = fmaxf(hmax, h); }
                }
            }
            hmax = __uint_as_float(__reduce_max_sync(full, __float_as_uint(hmax)));   // non-negative floats order like their bits
            // the prefilter admits r2 < cutsq + margin; its rounding error is below 2e-6 hmax (three fmas on terms <= 3 hmax, fp32)
            const float half_lim = 0.5f * (kc.cutsq + 1e-3f + 8e-6f * hmax);
            // the cell's own lipids as prefilter operands (-x, -y, -z, limit), staged where the segment sums go later
            if (lane < nb) {
                const float4 p = tile_x[own_off + ib + lane];
                const float nx = org.x - p.x, ny = org.y - p.y, nz = org.z - p.z;
                seg[lane] = make_float4(nx, ny, nz, half_lim - 0.5f * (nx * nx + ny * ny + nz * nz));
            }
            __syncwarp();
            // ---- phase 1: prefilter, hits as ballot masks ---------------------------------------------------------------------------------
            // words 0-3 always (an empty slot never passes), words 4-5 and 6-7 only when the cell has that many candidates (11 % / <1 %
            // of the RBC's cells): straight-line code, 3 FFMA + compare + vote per 32 pairs, one store per lipid
            for (int ii = 0; ii < nb; ++ii) {
                const float4 q = seg[ii];                        // broadcast
                unsigned *const row = mask + ii * kTileWords;
                unsigned m[4];
                #pragma unroll
                for (int u = 0; u < 4; ++u) m[u] = __ballot_sync(full, fmaf(q.z, cz[u], fmaf(q.y, cy[u], fmaf(q.x, cx[u], ch[u]))) < q.w);
                if (lane == 0) *reinterpret_cast<uint4 *>(row) = make_uint4(m[0], m[1], m[2], m[3]);
                if (U > 4) {
                    #pragma unroll
                    for (int u = 4; u < 6; ++u) m[u - 4] = __ballot_sync(full, fmaf(q.z, cz[u], fmaf(q.y, cy[u], fmaf(q.x, cx[u], ch[u]))) < q.w);
                    if (lane == 0) *reinterpret_cast<uint2 *>(row + 4) = make_uint2(m[0], m[1]);
                    if (U > 6) {
                        #pragma unroll
                        for (int u = 6; u < 8; ++u) m[u - 6] = __ballot_sync(full, fmaf(q.z, cz[u], fmaf(q.y, cy[u], fmaf(q.x, cx[u], ch[u]))) < q.w);
                        if (lane == 0) *reinterpret_cast<uint2 *>(row + 6) = make_uint2(m[0], m[1]);
                    }
                }
            }
            __syncwarp();
            // ---- phase 1.5: the hit list of the cell in (i, candidate) order, cut into 32 equal chunks ------------------------------------
            int cnt = 0;
            if (lane < nb) for (int w = 0; w < U; ++w) cnt += __popc(mask[lane * kTileWords + w]);
            int incl = cnt;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(full, incl, d); if (lane >= d) incl += t; }
            const int excl = incl - cnt;
            const int H = __shfl_sync(full, incl, 31);
            const int chunk = (H + 31) >> 5;
            const int p0 = min(lane * chunk, H), p1 = min(p0 + chunk, H);
            int ii = 0;                                          // first lipid of this lane's chunk: number of lists that end at or before p0
            for (int q = 0; q < nb; ++q) ii += (__shfl_sync(full, incl, q) <= p0) ? 1 : 0;
            int skip = p0 - __shfl_sync(full, excl, min(ii, 31));
            int w = 0; unsigned bits = 0;
            if (p0 < p1) {
                for (w = 0; w < U; ++w) {                        // the word of the list of ii that holds position `skip`
                    bits = mask[ii * kTileWords + w];
                    const int k = __popc(bits);
                    if (skip < k) break;
                    skip -= k;
                }
                for (; skip > 0; --skip) bits &= bits - 1u;      // ... and the bit
            }
            // ---- phase 2: every lane evaluates its chunk ---------------------------------------------------------------------------------------
            float fx = 0, fy = 0, fz = 0, tx = 0, ty = 0, tz = 0, sB = 0;
            F3 xi = {0, 0, 0}, mi = {0, 0, 0};
            int cur = -1;
            for (int it = 0; it < chunk; ++it) {
                const bool act = p0 + it < p1;
                while (act && bits == 0u) {                      // next non-empty word of the list (of the next lipid when this one is exhausted)
                    if (++w >= U) { w = 0; ++ii; }
                    bits = mask[ii * kTileWords + w];
                }
                int j = 0;
                if (act) { j = w * 32 + __ffs(bits) - 1; bits &= bits - 1u; }
                if (act && ii != cur) {
                    if (cur >= 0) {                              // close the segment of the previous lipid: slot (lipid + lane) is this lane's alone
                        const int slot = cur + lane;
                        seg[2 * slot] = make_float4(fmaf(sB, mi.x, fx), fmaf(sB, mi.y, fy), fmaf(sB, mi.z, fz), tx);
                        seg[2 * slot + 1] = make_float4(ty, tz, 0.f, 0.f);
                    }
                    cur = ii;
                    const float4 p = tile_x[own_off + ib + ii], q = tile_n[own_off + ib + ii];
                    xi = {p.x, p.y, p.z}; mi = {q.x, q.y, q.z};
                    fx = fy = fz = tx = ty = tz = sB = 0.f;
                }
                if (act) ll_eval_tile(kc, lo_bits, span, xi, mi, tile_x[j], tile_n[j], fx, fy, fz, tx, ty, tz, sB);
            }
            if (cur >= 0) {
                const int slot = cur + lane;
                seg[2 * slot] = make_float4(fmaf(sB, mi.x, fx), fmaf(sB, mi.y, fy), fmaf(sB, mi.z, fz), tx);
                seg[2 * slot + 1] = make_float4(ty, tz, 0.f, 0.f);
            }
            __syncwarp();
            // ---- phase 3: lane = lipid; its segments in lane order -----------------------------------------------------------------------------
            {
                float gx = 0, gy = 0, gz = 0, hx = 0, hy = 0, hz = 0;
                const int i = s0 + ib + lane;
                const bool mine = lane < nb;
                if (mine && cnt > 0) {
                    const int L0 = excl / chunk, L1 = (incl - 1) / chunk;
                    for (int L = L0; L <= L1; ++L) {
                        const float4 u = seg[2 * (lane + L)], v = seg[2 * (lane + L) + 1];
                        gx += u.x; gy += u.y; gz += u.z; hx += u.w; hy += v.x; hz += v.y;
                    }
                }
                const bool live = mine && i >= l0 && i < l1;
                F3 xo = {0, 0, 0}, mo = {0, 0, 0};
                if (a.world > 1 && live) {                       // decomposed run: the foreign-protein epilogue needs the lipid itself
                    const float4 p = tile_x[own_off + ib + lane], q = tile_n[own_off + ib + lane];
                    xo = {p.x, p.y, p.z}; mo = {q.x, q.y, q.z};
                }
                ll_finish(a, i, live, a.stencil + (size_t)c * kStencilStride, xo, mo, gx, gy, gz, hx, hy, hz, 0.f);
            }
            __syncwarp();                                        // masks and segments are rewritten by the next pass
        }
    }
}

} // namespace orbc
