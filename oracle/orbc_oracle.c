/* TEST INFRASTRUCTURE — see orbc_oracle.h.  CPU restatement of OpenRBC's per-timestep hot path.
 * Compile with: gcc -std=c11 -O2 -fno-fast-math -ffp-contract=off (strict IEEE fp32, no FMA).
 * Every function names the reference file:line it follows; none of this is on the product path. */
#define _GNU_SOURCE
#include "orbc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NT 6 /* forcefield_canonical.h:37 n_type */

/* ============================================================================================
 * Force field — forcefield_canonical.h:23-28 (POW/cutsq/rep/att) and :30-156 (tables, init()).
 * rep() is evaluated in fp32, att() in fp64 (the literal -2.0 promotes), lj_* in fp64, exactly as the
 * constexpr / static-init expressions of the reference do.
 * ============================================================================================ */
static float powf_chain(float base, int expo) { return expo != 0 ? base * powf_chain(base, expo - 1) : 1.0f; }
static float ff_cutsq(float cut) { return cut * cut; }
static float ff_rep(float cut, float req, float eps) { return eps / powf_chain(cut - req, 8); }
static float ff_att(float cut, float req, float eps) { return (float)(-2.0 * eps / powf_chain(cut - req, 4)); }

void orc_forcefield_canonical(orc_forcefield *ff)
{
    memset(ff, 0, sizeof(*ff));
    const float mass[NT] = {1, 4, 4, 1, 10, 10};
    const float radius[NT] = {0.56125f, 1.12375f, 1.12375f, 0.56125f, 1.5f, 0.5f};
    memcpy(ff->mass, mass, sizeof mass);
    memcpy(ff->radius, radius, sizeof radius);
    const float cutll = 2.6f, reqll = 1.1225f, epsll = 1.2f, alphall = 1.55f;
    const float cutlp[NT] = {cutll, 2.6f, 2.6f, 2.6f, 0, 0};
    const float reqlp[NT] = {reqll, 1.685f, 1.685f, 1.1225f, 0, 0};
    const float epslp[NT] = {epsll, 1.4f, 2.8f, 2.8f, 0, 0};
    const float alphalp[NT] = {alphall, 5, 5, 5, 0, 0};
    float cutpp[36] = {0}, reqpp[36] = {0}, epspp[36] = {0};
    const float req_in[3][3] = {{2.245f, 2.245f, 1.685f}, {2.245f, 2.245f, 1.685f}, {1.685f, 1.685f, 1.1225f}};
    for (int c = 0; c < 4; ++c) { cutpp[c] = cutlp[c]; reqpp[c] = reqlp[c]; epspp[c] = epslp[c]; }
    for (int r = 1; r < 4; ++r) {
        cutpp[6 * r] = cutlp[r]; reqpp[6 * r] = reqlp[r]; epspp[6 * r] = epslp[r];
        for (int c = 1; c < 4; ++c) { cutpp[6 * r + c] = 2.6f; reqpp[6 * r + c] = req_in[r - 1][c - 1]; epspp[6 * r + c] = 1.0f; }
    }
    ff->cutll = cutll; ff->alphall = alphall;
    ff->cutsqll = ff_cutsq(cutll);
    ff->repll = ff_rep(cutll, reqll, epsll);
    ff->attll = ff_att(cutll, reqll, epsll);
    for (int i = 0; i < NT; ++i) { ff->cutlp[i] = cutlp[i]; ff->alphalp[i] = alphalp[i]; }
    ff->cutsqlp[0] = ff->cutsqll; ff->replp[0] = ff->repll; ff->attlp[0] = ff->attll;
    for (int i = 1; i < 4; ++i) {
        ff->cutsqlp[i] = ff_cutsq(cutlp[i]);
        ff->replp[i] = ff_rep(cutlp[i], reqlp[i], epslp[i]);
        ff->attlp[i] = ff_att(cutlp[i], reqlp[i], epslp[i]);
    }
    memcpy(ff->cutpp, cutpp, sizeof cutpp);
    for (int c = 0; c < 4; ++c) { ff->cutsqpp[c] = ff->cutsqlp[c]; ff->reppp[c] = ff->replp[c]; }
    for (int r = 1; r < 4; ++r) {
        ff->cutsqpp[6 * r] = ff->cutsqlp[r]; ff->reppp[6 * r] = ff->replp[r];
        for (int c = 1; c < 4; ++c) {
            int k = 6 * r + c;
            ff->cutsqpp[k] = ff_cutsq(cutpp[k]);
            ff->reppp[k] = ff_rep(cutpp[k], reqpp[k], epspp[k]);
        }
    }
    /* lj tables, forcefield_canonical.h:76-99,143-148 */
    const float lj_eps[36] = {0, 0, 0, 0, 1, 1,  0, 0, 0, 0, 1, 1,  0, 0, 0, 0, 1, 1,  0, 0, 0, 0, 0, 0,  1, 1, 1, 0, 0, 0,  1, 1, 1, 0, 0, 0};
    const float lj_sig[36] = {0, 0, 0, 0, 1, 1,  0, 0, 0, 0, 3.4f, 3.4f,  0, 0, 0, 0, 3.4f, 1,  0, 0, 0, 0, 0, 0,
                              1, 3.4f, 3.4f, 0, 3, 1.8f,  1, 3.4f, 1, 0, 1.8f, 1};
    const float c1 = 1.1225f, c34 = (float)(3.4 * 1.1225);
    const float lj_cut[36] = {0, 0, 0, 0, c1, c1,  0, 0, 0, 0, c34, c34,  0, 0, 0, 0, c34, c1,  0, 0, 0, 0, 0, 0,
                              c1, c34, c34, 0, 0, 0,  c1, c34, c1, 0, 0, 0};
    for (int i = 0; i < 36; ++i) {
        ff->lj_cutsq[i] = lj_cut[i] * lj_cut[i];
        ff->lj_lj1[i] = (float)(48.0 * lj_eps[i] * pow((double)lj_sig[i], 12.0));
        ff->lj_lj2[i] = (float)(24.0 * lj_eps[i] * pow((double)lj_sig[i], 6.0));
    }
    const float r0[4] = {2.25f, 1.1225f, 2.25f, 2.24f}, K[4] = {57, 57, 57, 57};
    memcpy(ff->r0, r0, sizeof r0);
    memcpy(ff->K, K, sizeof K);
}

/* ============================================================================================
 * small fp32 vector helpers in the reference's operation order (math_vector_base.h:203-246)
 * ============================================================================================ */
static inline float dot3(const float *u, const float *v) { float s = 0; s += u[0] * v[0]; s += u[1] * v[1]; s += u[2] * v[2]; return s; }
static inline float normsq3(const float *u) { return dot3(u, u); }
static inline void cross3(const float *u, const float *v, float *x)
{
    x[0] = u[1] * v[2] - u[2] * v[1];
    x[1] = u[2] * v[0] - u[0] * v[2];
    x[2] = u[0] * v[1] - u[1] * v[0];
}
static inline void normalize3(float *u) { float s = 1.0f / sqrtf(normsq3(u)); u[0] *= s; u[1] *= s; u[2] *= s; }

/* ============================================================================================
 * Spatial index
 * ============================================================================================ */
/* voronoi.h:123-140: fp32 sequential sum in slot order, then * (1/count). */
void orc_update_centroid(int n_cells, const int *cell_start, const float *x, float *centroids)
{
    for (int i = 0; i < n_cells; ++i) {
        float c[3] = {0, 0, 0};
        for (int j = cell_start[i]; j < cell_start[i + 1]; ++j) for (int d = 0; d < 3; ++d) c[d] += x[3 * j + d];
        float s = 1.0f / (float)(cell_start[i + 1] - cell_start[i]);
        for (int d = 0; d < 3; ++d) centroids[3 * i + d] = c[d] * s;
    }
}

/* reorder_morton.h:25-31 */
static uint32_t bit_space3(uint32_t x)
{
    x = (x | (x << 12)) & 0X00FC003FU;
    x = (x | (x << 6)) & 0X381C0E07U;
    x = (x | (x << 4)) & 0X190C8643U;
    x = (x | (x << 2)) & 0X49249249U;
    return x;
}

/* reorder_morton.h:33-42 with bsize = 2000 (runtime_parameter.h:111-115): 2*x in fp32, + bsize in fp64, cast to uint32 */
uint32_t orc_morton_encode(float x, float y, float z)
{
    uint32_t i = (uint32_t)((double)(2 * x) + 2000.0);
    uint32_t j = (uint32_t)((double)(2 * y) + 2000.0);
    uint32_t k = (uint32_t)((double)(2 * z) + 2000.0);
    return bit_space3(i) | (bit_space3(j) << 1) | (bit_space3(k) << 2);
}

typedef struct { uint32_t key; int idx; } keyidx;
static int cmp_keyidx(const void *a, const void *b)
{
    const keyidx *p = a, *q = b;
    if (p->key != q->key) return p->key < q->key ? -1 : 1;
    return p->idx < q->idx ? -1 : (p->idx > q->idx);
}

/* reorder_morton.h:44-122: ascending by key; equal keys keep ascending original index (the reference's order among
 * equal keys is std::sort's — SURVEY §8 a9: no duplicates occur on the fixtures). */
void orc_morton_perm(int n, const float *pts, int *perm, uint32_t *keys_out)
{
    keyidx *a = malloc(sizeof(keyidx) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) { a[i].key = orc_morton_encode(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]); a[i].idx = i; }
    if (keys_out) for (int i = 0; i < n; ++i) keys_out[i] = a[i].key;
    qsort(a, n, sizeof(keyidx), cmp_keyidx);
    for (int i = 0; i < n; ++i) perm[i] = a[i].idx;
    free(a);
}

/* uniform grid over the centroids; replaces the k-d tree's role (kdtree.h) with identical results:
 * exact nearest / exact within-radius sets, candidates visited in ascending id inside a bin. */
typedef struct {
    float lo[3], h;
    int dim[3];
    int *start, *items;
} cgrid;

static void cgrid_bin(const cgrid *g, const float *p, int *b)
{
    for (int d = 0; d < 3; ++d) {
        float t = floorf((p[d] - g->lo[d]) / g->h);
        int v = t < 0 ? 0 : (t >= (float)g->dim[d] ? g->dim[d] - 1 : (int)t);
        if (!(t == t)) v = 0; /* NaN */
        b[d] = v;
    }
}

static void cgrid_build(cgrid *g, int n, const float *c, float h)
{
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = 0; i < n; ++i) for (int d = 0; d < 3; ++d) {
        float v = c[3 * i + d];
        if (v == v) { if (v < lo[d]) lo[d] = v; if (v > hi[d]) hi[d] = v; }
    }
    long nb = 1;
    for (;;) {
        nb = 1;
        for (int d = 0; d < 3; ++d) { g->lo[d] = lo[d]; g->dim[d] = (int)floorf((hi[d] - lo[d]) / h) + 1; if (g->dim[d] < 1) g->dim[d] = 1; nb *= g->dim[d]; }
        if (nb <= (8L << 20)) break;
        h *= 1.5f;
    }
    g->h = h;
    g->start = calloc((size_t)nb + 1, sizeof(int));
    g->items = malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int *bin = malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) {
        if (c[3 * i] != c[3 * i]) { bin[i] = -1; continue; } /* NaN centroid of an empty cell: never a candidate */
        int b[3]; cgrid_bin(g, c + 3 * i, b);
        bin[i] = (b[2] * g->dim[1] + b[1]) * g->dim[0] + b[0];
        g->start[bin[i] + 1]++;
    }
    for (long k = 0; k < nb; ++k) g->start[k + 1] += g->start[k];
    int *fill = calloc((size_t)nb, sizeof(int));
    for (int i = 0; i < n; ++i) if (bin[i] >= 0) g->items[g->start[bin[i]] + fill[bin[i]]++] = i;
    free(fill); free(bin);
}
static void cgrid_free(cgrid *g) { free(g->start); free(g->items); }

static inline float dist2(const float *a, const float *b)
{
    float d[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    return normsq3(d);
}

/* voronoi.h:179-216 + kdtree.h:206-236,336-347.  The reference keeps the candidate with the smallest sqrt-ed fp32
 * distance (strict <, so search order decides exact ties); the restatement takes the smallest squared distance and
 * the lowest id among equals, and flags the points where the two readings can differ. */
int orc_assign_nearest(long n, const float *x, int n_cells, const float *c, int *aff, int *tie_flag)
{
    cgrid g; cgrid_build(&g, n_cells, c, 6.0f);
    int ties = 0;
    for (long i = 0; i < n; ++i) {
        const float *p = x + 3 * i;
        int b[3]; cgrid_bin(&g, p, b);
        float best = INFINITY, second = INFINITY; int bi = -1;
        int maxring = g.dim[0] > g.dim[1] ? g.dim[0] : g.dim[1]; if (g.dim[2] > maxring) maxring = g.dim[2];
        for (int k = 0; k <= maxring; ++k) {
            /* shell k of bins (Chebyshev distance exactly k) */
            for (int dz = -k; dz <= k; ++dz) for (int dy = -k; dy <= k; ++dy) for (int dx = -k; dx <= k; ++dx) {
                int m = abs(dx) > abs(dy) ? abs(dx) : abs(dy); if (abs(dz) > m) m = abs(dz);
                if (m != k) continue;
                int bx = b[0] + dx, by = b[1] + dy, bz = b[2] + dz;
                if (bx < 0 || by < 0 || bz < 0 || bx >= g.dim[0] || by >= g.dim[1] || bz >= g.dim[2]) continue;
                int bin = (bz * g.dim[1] + by) * g.dim[0] + bx;
                for (int q = g.start[bin]; q < g.start[bin + 1]; ++q) {
                    int j = g.items[q];
                    float d2 = dist2(p, c + 3 * j);
                    if (d2 < best || (d2 == best && j < bi)) { second = best; best = d2; bi = j; }
                    else if (d2 < second) second = d2;
                }
            }
            /* everything outside shells 0..k is farther than k*h from p */
            float reach = (float)k * g.h;
            if (bi >= 0 && second <= reach * reach) break;
        }
        aff[i] = bi;
        int tie = 0;
        if (bi >= 0 && second < INFINITY) {
            float r1 = sqrtf(best), r2 = sqrtf(second);
            tie = (r2 <= nextafterf(r1, INFINITY));
        }
        if (tie_flag) tie_flag[i] = tie;
        ties += tie;
    }
    cgrid_free(&g);
    return ties;
}

/* voronoi.h:214-231 at one thread: local_index = arrival order = ascending particle index; exclusive scan; scatter. */
void orc_partition(long n, int n_cells, const int *aff, int *cell_start, int *cells, int *local_index)
{
    for (int i = 0; i <= n_cells; ++i) cell_start[i] = 0;
    for (long i = 0; i < n; ++i) local_index[i] = cell_start[aff[i] + 1]++;
    for (int i = 0; i < n_cells; ++i) cell_start[i + 1] += cell_start[i];
    for (long i = 0; i < n; ++i) cells[local_index[i] + cell_start[aff[i]]] = (int)i;
}

void orc_gather3(long n, const int *cells, const float *src, float *dst)
{
    for (long j = 0; j < n; ++j) for (int d = 0; d < 3; ++d) dst[3 * j + d] = src[3 * (long)cells[j] + d];
}
void orc_gather1(long n, const int *cells, const int *src, int *dst) { for (long j = 0; j < n; ++j) dst[j] = src[cells[j]]; }

static int cmp_int(const void *a, const void *b) { int p = *(const int *)a, q = *(const int *)b; return p < q ? -1 : p > q; }

static int stencil_from_grid(const cgrid *g, const float *c, int cell, float rmax, int *out, int cap)
{
    const float *p = c + 3 * cell;
    if (p[0] != p[0]) return 0;
    int b[3]; cgrid_bin(g, p, b);
    int k = (int)ceilf(rmax / g->h);
    int n = 0;
    const float r2 = rmax * rmax;
    for (int bz = b[2] - k; bz <= b[2] + k; ++bz) for (int by = b[1] - k; by <= b[1] + k; ++by) for (int bx = b[0] - k; bx <= b[0] + k; ++bx) {
        if (bx < 0 || by < 0 || bz < 0 || bx >= g->dim[0] || by >= g->dim[1] || bz >= g->dim[2]) continue;
        int bin = (bz * g->dim[1] + by) * g->dim[0] + bx;
        for (int q = g->start[bin]; q < g->start[bin + 1]; ++q) {
            int j = g->items[q];
            if (dist2(c + 3 * j, p) < r2) { if (n < cap) out[n] = j; ++n; }
        }
    }
    qsort(out, n < cap ? n : cap, sizeof(int), cmp_int);
    return n;
}

int orc_stencil(int n_cells, const float *c, int cell, float rmax, int *out, int cap)
{
    cgrid g; cgrid_build(&g, n_cells, c, 9.0f);
    int n = stencil_from_grid(&g, c, cell, rmax, out, cap);
    cgrid_free(&g);
    return n;
}

/* ============================================================================================
 * Pair forces
 * ============================================================================================ */
/* pairwise_kernel.h:30-68 (type 0 constants) and pairwise_kernel_fused.h:22-61 (per protein type): the same poly 4-8
 * form.  dx = x1 - x2.  Outputs the force on particle 1 (particle 2 gets -f) and both torque increments. */
static inline void poly48(float cut, float att, float rep, float alpha, const float *dx, float r_sq, const float *mu1, const float *mu2,
                          float *f, float *t1 /* -= on particle 1 */, float *t2 /* -= on particle 2 */)
{
    const float r = sqrtf(r_sq);
    const float rinv = 1.0f / r;
    const float u[3] = {dx[0] * rinv, dx[1] * rinv, dx[2] * rinv};
    const float ninj = dot3(mu1, mu2);
    const float niu = dot3(mu1, u);
    const float nju = dot3(mu2, u);
    const float a = ninj - niu * nju;
    const float A = 1.0f + alpha * (a - 1.0f);
    const float rc = cut - r;
    const float rc3 = rc * rc * rc;
    const float rc4 = rc * rc3;
    const float rc7 = rc3 * rc4;
    const float pni[3] = {mu1[0] - niu * u[0], mu1[1] - niu * u[1], mu1[2] - niu * u[2]};
    const float pnj[3] = {mu2[0] - nju * u[0], mu2[1] - nju * u[1], mu2[2] - nju * u[2]};
    const float ua = att * rc4;
    const float alphaua = alpha * ua;
    const float alphauar = alphaua / r;
    const float fra = 8.0f * rep * rc7 + A * 4.0f * att * rc3;
    for (int d = 0; d < 3; ++d) {
        f[d] = fra * u[d] + alphauar * (nju * pni[d] + niu * pnj[d]);
        t1[d] = alphaua * pnj[d];
        t2[d] = alphaua * pni[d];
    }
}

/* pairwise_kernel_fused.h:63-77 */
static inline void lj126(float lj1, float lj2, const float *dx, float rsq, float *f)
{
    const float r2inv = 1.0f / rsq;
    const float r6inv = r2inv * r2inv * r2inv;
    const float forcelj = r6inv * (lj1 * r6inv - lj2);
    const float fpair = forcelj * r2inv;
    for (int d = 0; d < 3; ++d) f[d] = dx[d] * fpair;
}

/* pairwise_kernel_fused.h:79-97 */
static inline void rep8(float cut, float rep, const float *dx, float r_sq, float *f)
{
    const float r = sqrtf(r_sq);
    const float rc = cut - r;
    const float rc3 = rc * rc * rc;
    const float rc4 = rc * rc3;
    const float rc7 = rc3 * rc4;
    const float fr = 8.0f * rep * rc7;
    for (int d = 0; d < 3; ++d) f[d] = fr * (dx[d] / r);
}

#define ADD3(a, i, v) do { a[3 * (i)] += v[0]; a[3 * (i) + 1] += v[1]; a[3 * (i) + 2] += v[2]; } while (0)
#define SUB3(a, i, v) do { a[3 * (i)] -= v[0]; a[3 * (i) + 1] -= v[1]; a[3 * (i) + 2] -= v[2]; } while (0)

/* compute_pairwise_fused.h:91-139 */
static void ll_pair(const orc_forcefield *ff, const float *x, const float *nn, float *f, float *t, int i, int j, long *cnt)
{
    float dx[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
    float r_sq = normsq3(dx);
    if (cnt) cnt[0]++;
    if (r_sq < ff->cutsqll && r_sq > 1e-5) {
        float fv[3], t1[3], t2[3];
        poly48(ff->cutll, ff->attll, ff->repll, ff->alphall, dx, r_sq, nn + 3 * i, nn + 3 * j, fv, t1, t2);
        ADD3(f, i, fv); SUB3(t, i, t1);
        SUB3(f, j, fv); SUB3(t, j, t2);
        if (cnt) cnt[1]++;
    }
}

/* compute_pairwise_fused.h:143-179: protein i (cell1) against lipid j (cell2) */
static void pl_pair(const orc_forcefield *ff, const float *xp, const float *np_, const int *type_p, float *fp, float *tp,
                    const float *xl, const float *nl, float *fl, float *tl, int i, int j, long *cnt)
{
    float dx[3] = {xp[3 * i] - xl[3 * j], xp[3 * i + 1] - xl[3 * j + 1], xp[3 * i + 2] - xl[3 * j + 2]};
    float r_sq = normsq3(dx);
    const int type = type_p[i];
    if (cnt) cnt[2]++;
    if (r_sq < ff->cutsqlp[type] && r_sq > 1e-5) {
        float fv[3], t1[3], t2[3];
        poly48(ff->cutlp[type], ff->attlp[type], ff->replp[type], ff->alphalp[type], dx, r_sq, np_ + 3 * i, nl + 3 * j, fv, t1, t2);
        ADD3(fp, i, fv); SUB3(tp, i, t1);
        SUB3(fl, j, fv); SUB3(tl, j, t2);
        if (cnt) cnt[3]++;
    } else if (r_sq < ff->lj_cutsq[type] && r_sq > 1e-5) {
        float fv[3];
        lj126(ff->lj_lj1[type], ff->lj_lj2[type], dx, r_sq, fv);
        ADD3(fp, i, fv); SUB3(fl, j, fv);
        if (cnt) cnt[4]++;
    }
}

/* compute_pairwise_fused.h:182-236 */
static void pp_pair(const orc_forcefield *ff, const float *x, const int *type, float *f, int i, int j, long *cnt)
{
    float dx[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
    float r_sq = normsq3(dx);
    const int type12 = type[i] + type[j] * NT;
    if (cnt) cnt[5]++;
    if (r_sq < ff->cutsqpp[type12] && r_sq > 1e-5) {
        float fv[3]; rep8(ff->cutpp[type12], ff->reppp[type12], dx, r_sq, fv);
        ADD3(f, i, fv); SUB3(f, j, fv);
        if (cnt) cnt[6]++;
    } else if (r_sq < ff->lj_cutsq[type12] && r_sq > 1e-5) {
        float fv[3]; lj126(ff->lj_lj1[type12], ff->lj_lj2[type12], dx, r_sq, fv);
        ADD3(f, i, fv); SUB3(f, j, fv);
        if (cnt) cnt[7]++;
    }
}

/* compute_pairwise_fused.h:238-320 with one thread owning [0, n_cells): every cell2 is "in range", so the Newton-on
 * branches are the ones taken.  Stencil cells are visited in ascending id (the reference visits them in k-d tree
 * traversal order; only the fp32 summation order differs). */
void orc_pairwise_fused(const orc_forcefield *ff, int n_cells, const float *centroids,
                        long n_l, const float *xl, const float *nl, const int *cs_l, float *fl, float *tl,
                        long n_p, const float *xp, const float *np_, const int *type_p, const int *cs_p, float *fp, float *tp,
                        long *cnt)
{
    (void)n_l; (void)n_p;
    if (cnt) memset(cnt, 0, 8 * sizeof(long));
    cgrid g; cgrid_build(&g, n_cells, centroids, 9.0f);
    enum { CAP = 4096 };
    int *st = malloc(sizeof(int) * CAP);
    for (int c1 = 0; c1 < n_cells; ++c1) {
        const int l1b = cs_l[c1], l1e = cs_l[c1 + 1];
        const int p1b = cs_p ? cs_p[c1] : 0, p1e = cs_p ? cs_p[c1 + 1] : 0;
        for (int i = l1b; i < l1e; ++i) for (int j = i + 1; j < l1e; ++j) ll_pair(ff, xl, nl, fl, tl, i, j, cnt);   /* :256 */
        for (int i = p1b; i < p1e; ++i) for (int j = i + 1; j < p1e; ++j) pp_pair(ff, xp, type_p, fp, i, j, cnt);   /* :257 */
        int n9 = stencil_from_grid(&g, centroids, c1, 9.0f, st, CAP);                                                 /* :260 */
        if (n9 > CAP) n9 = CAP;
        const float *q = centroids + 3 * c1;
        for (int s = 0; s < n9; ++s) {
            const int c2 = st[s];
            const float d2 = dist2(centroids + 3 * c2, q);
            if (cs_p && c2 > c1)                                                                                      /* :262-276 */
                for (int i = p1b; i < p1e; ++i) for (int j = cs_p[c2]; j < cs_p[c2 + 1]; ++j) pp_pair(ff, xp, type_p, fp, i, j, cnt);
            if (cs_p && d2 < 8.0f * 8.0f)                                                                             /* :278-297 */
                for (int i = p1b; i < p1e; ++i) for (int j = cs_l[c2]; j < cs_l[c2 + 1]; ++j) pl_pair(ff, xp, np_, type_p, fp, tp, xl, nl, fl, tl, i, j, cnt);
            if (d2 < 6.0f * 6.0f && c2 > c1)                                                                          /* :299-315 */
                for (int i = l1b; i < l1e; ++i) for (int j = cs_l[c2]; j < cs_l[c2 + 1]; ++j) ll_pair(ff, xl, nl, fl, tl, i, j, cnt);
        }
    }
    free(st);
    cgrid_free(&g);
}

/* container.h:39-58 */
void orc_build_tag2idx(long n, const int *tag, int *map, long map_size)
{
    for (long i = 0; i < map_size; ++i) map[i] = -1;
    for (long i = 0; i < n; ++i) map[tag[i]] = (int)i;
}

/* compute_bonded.h:89-146 (one thread: every particle is in range, no shadow buffer needed) */
void orc_bonded(const orc_forcefield *ff, long n_bonds, const int *bonds, const int *tag2idx, const float *x, float *f)
{
    for (long l = 0; l < n_bonds; ++l) {
        const int type = bonds[3 * l], p1 = tag2idx[bonds[3 * l + 1]], p2 = tag2idx[bonds[3 * l + 2]];
        const float dx[3] = {x[3 * p2] - x[3 * p1], x[3 * p2 + 1] - x[3 * p1 + 1], x[3 * p2 + 2] - x[3 * p1 + 2]};
        const float rinv = 1.0f / sqrtf(normsq3(dx));
        const float cur_r = ff->K[type] * (1 - ff->r0[type] * rinv);
        const float force[3] = {cur_r * dx[0], cur_r * dx[1], cur_r * dx[2]};
        ADD3(f, p1, force);
        SUB3(f, p2, force);
    }
}

/* ============================================================================================
 * Integrators
 * ============================================================================================ */
void orc_clear_force(long n, float *f, float *t) { memset(f, 0, sizeof(float) * 3 * (size_t)n); memset(t, 0, sizeof(float) * 3 * (size_t)n); }

void orc_post_torque(long n, const float *nn, float *t)
{
    for (long i = 0; i < n; ++i) { float r[3]; cross3(nn + 3 * i, t + 3 * i, r); memcpy(t + 3 * i, r, sizeof r); }
}

static inline void bounce(float *x, float *v, double lo, double hi)
{
    for (int d = 0; d < 3; ++d) {
        if (x[d] < lo) { x[d] = (float)(lo + (lo - x[d])); v[d] = -v[d]; }
        else if (x[d] > hi) { x[d] = (float)(hi - (x[d] - hi)); v[d] = -v[d]; }
    }
}
void orc_bounce_back(long n, float *x, float *v, double lo, double hi) { for (long i = 0; i < n; ++i) bounce(x + 3 * i, v + 3 * i, lo, hi); }

/* n = normalize(n + cross(o, n) * dt) — integrate_langevin.h:126 / integrate_nh.h:218 */
static inline void rotate_director(float *nn, const float *o, float dt)
{
    float c[3]; cross3(o, nn, c);
    float g[3] = {nn[0] + c[0] * dt, nn[1] + c[1] * dt, nn[2] + c[2] * dt};
    normalize3(g);
    memcpy(nn, g, sizeof g);
}

/* integrate_langevin.h:99-149 */
void orc_verlet_langevin(const orc_forcefield *ff, long n, float *x, float *v, float *f, float *nn, float *o, float *t,
                         const int *type, double dt_, float eta, float kBT, const float *noise)
{
    const float dt = (float)dt_;
    const float inertia = 1.0f;
    const float dt_over_m = dt / inertia;
    float gamma[NT], sigma[NT];
    for (int i = 0; i < NT; ++i) {
        gamma[i] = (float)(6.0 * M_PI * eta * ff->radius[i]);
        sigma[i] = (float)(sqrtf(2 * kBT * gamma[i]) * sqrt(3.0 / dt_));
    }
    for (long i = 0; i < n; ++i) {
        float *X = x + 3 * i, *V = v + 3 * i, *F = f + 3 * i, *N = nn + 3 * i, *O = o + 3 * i, *T = t + 3 * i;
        float tq[3]; cross3(N, T, tq);
        for (int d = 0; d < 3; ++d) O[d] += dt_over_m * tq[d];
        rotate_director(N, O, dt);
        T[0] = T[1] = T[2] = 0;
        const int ty = type ? type[i] : 0;
        for (int d = 0; d < 3; ++d) {
            const float r = noise ? noise[3 * i + d] : 0.0f;
            F[d] -= gamma[ty] * V[d] + sigma[ty] * r;
        }
        const float s = dt / ff->mass[ty];
        for (int d = 0; d < 3; ++d) V[d] += F[d] * s;
        for (int d = 0; d < 3; ++d) X[d] += V[d] * dt;
        F[0] = F[1] = F[2] = 0;
    }
}

/* integrate_nh.h:178-235 (operator()); the zeta update of the destructor is orc_nh_zeta_update */
void orc_nh_initial_fused(const orc_forcefield *ff, long n, float *x, float *v, float *f, float *nn, float *o, float *t,
                          const int *type, double dt_, float zeta, double lo, double hi, double *ke)
{
    const float inertia = 1.0f;
    const float dt = (float)dt_;
    const float gamma = 1.0f / (1.0f + 0.5f * dt * zeta);
    double ke_local = 0;
    for (long i = 0; i < n; ++i) {
        float *X = x + 3 * i, *V = v + 3 * i, *F = f + 3 * i, *N = nn + 3 * i, *O = o + 3 * i, *T = t + 3 * i;
        const int ty = type ? type[i] : 0;
        const float s = 0.5f / ff->mass[ty] * dt;
        for (int d = 0; d < 3; ++d) V[d] = (V[d] + s * F[d]) * gamma;
        for (int d = 0; d < 3; ++d) X[d] += V[d] * dt;
        bounce(X, V, lo, hi);
        ke_local += 0.5f * ff->mass[ty] * normsq3(V);
        const float so = 0.5f / inertia * dt;
        for (int d = 0; d < 3; ++d) O[d] += so * T[d];
        rotate_director(N, O, dt);
        F[0] = F[1] = F[2] = 0; T[0] = T[1] = T[2] = 0;
    }
    *ke += ke_local;
}

/* integrate_nh.h:237-273 */
void orc_nh_final_fused(const orc_forcefield *ff, long n, float *v, const float *f, const float *nn, float *o, float *t,
                        const int *type, double dt_, float zeta, double *ke)
{
    const float dt = (float)dt_;
    const float inertia = 1.0f;
    const float dt_over_m = dt / inertia;
    double ke_local = 0;
    for (long i = 0; i < n; ++i) {
        float *V = v + 3 * i, *O = o + 3 * i, *T = t + 3 * i;
        const float *F = f + 3 * i, *N = nn + 3 * i;
        const int ty = type ? type[i] : 0;
        float tq[3]; cross3(N, T, tq); memcpy(T, tq, sizeof tq);
        const float s = 0.5f * dt;
        for (int d = 0; d < 3; ++d) V[d] += s * (F[d] / ff->mass[ty] - zeta * V[d]);
        const float so = 0.5f * dt_over_m;
        for (int d = 0; d < 3; ++d) O[d] += so * T[d];
        ke_local += 0.5f * ff->mass[ty] * normsq3(V);
    }
    *ke += ke_local;
}

/* verlet_nh_final, integrate_nh.h:155-176 (the unfused second half-kick; post_torque has run before it) */
void orc_nh_final(const orc_forcefield *ff, long n, float *v, const float *f, float *o, const float *t, const int *type, double dt_, float zeta)
{
    const float dt = (float)dt_;
    const float inertia = 1.0f;
    for (long i = 0; i < n; ++i) {
        float *V = v + 3 * i, *O = o + 3 * i;
        const float *F = f + 3 * i, *T = t + 3 * i;
        const int ty = type ? type[i] : 0;
        for (int d = 0; d < 3; ++d) V[d] += 0.5f * (F[d] / ff->mass[ty] - zeta * V[d]) * dt;
        for (int d = 0; d < 3; ++d) O[d] += 0.5f * T[d] / inertia * dt;
    }
}

/* verlet_nh_update::operator(), integrate_nh.h:78-89: the kinetic energy of one container (its destructor is orc_nh_zeta_update) */
void orc_nh_update(const orc_forcefield *ff, long n, const float *v, const int *type, double *ke)
{
    double ke_local = 0;
    for (long i = 0; i < n; ++i) ke_local += 0.5f * ff->mass[type ? type[i] : 0] * normsq3(v + 3 * i);
    *ke += ke_local;
}

/* destructor of the fused NH kernels, integrate_nh.h:181-185 / 240-244; Q defaults to 0.01 n (runtime_parameter.h:76) */
float orc_nh_zeta_update(float zeta, float *Q, double dt, float kBT, double ke, int n)
{
    if (!*Q) *Q = (float)(n * 0.01);
    zeta += 0.5 * dt / *Q * (ke - 0.5 * 3.0 * n * kBT);
    return zeta;
}

/* destructor of the UNFUSED verlet_nh_update, integrate_nh.h:72-76: `constant::onehalf * n * parameter.kBT` is a product of
 * floats there (real kBT, runtime_parameter.h:49), where the fused kernels spell 0.5 * 3.0 * n * kBT in double */
float orc_nh_zeta_update_unfused(float zeta, float *Q, double dt, float kBT, double ke, int n)
{
    if (!*Q) *Q = (float)(n * 0.01);
    const float target = 1.5f * n * kBT;
    zeta += 0.5f * dt / *Q * (ke - target);
    return zeta;
}

/* openrbc.cpp:114-131 */
void orc_opt_move(const orc_forcefield *ff, long n, float *x, float *nn, const float *f, const float *t, const int *type,
                  double dt_, double dr_opt, double dn_opt)
{
    for (long i = 0; i < n; ++i) {
        const int ty = type ? type[i] : 0;
        const float m = ff->mass[ty];
        const float dx[3] = {f[3 * i] / m, f[3 * i + 1] / m, f[3 * i + 2] / m};
        float dn[3]; cross3(t + 3 * i, nn + 3 * i, dn);
        const float ndx = sqrtf(normsq3(dx)), ndn = sqrtf(normsq3(dn));
        double dt;
        if (ndx > dr_opt || ndn > dn_opt) { double a = dr_opt / ndx, b = dn_opt / ndn; dt = b < a ? b : a; } /* std::min(a, b) */
        else dt = dt_;
        const float sx = (float)(dt / m), sn = (float)dt;
        for (int d = 0; d < 3; ++d) x[3 * i + d] += f[3 * i + d] * sx;
        for (int d = 0; d < 3; ++d) nn[3 * i + d] += dn[d] * sn;
        normalize3(nn + 3 * i);
    }
}

/* compute_temperature.h:23-29 */
double orc_temperature(const orc_forcefield *ff, long n_l, const float *vl, long n_p, const float *vp, const int *type_p)
{
    double ek = 0.0;
    for (long i = 0; i < n_l; ++i) ek += ff->mass[0] * normsq3(vl + 3 * i);
    for (long i = 0; i < n_p; ++i) ek += ff->mass[type_p[i]] * normsq3(vp + 3 * i);
    return ek / (3.0 * (double)(n_l + n_p));
}

/* constrain_volume.h:26-83 at one thread, oddities kept: cell_normal persists across calls (function-static in the
 * reference) and the mass factor is indexed by the CELL index i (lipid.type[i] == 0 always; prote.type[i]). */
float orc_constrain_volume(const orc_forcefield *ff, int n_cells, const float *centroids, float *cell_normal,
                           long n_l, const float *nl, const int *cs_l, float *fl,
                           long n_p, const int *type_p, const int *cs_p, float *fp, float target, float strength)
{
    (void)n_l;
    float gc[3] = {0, 0, 0};
    for (int i = 0; i < n_cells; ++i) for (int d = 0; d < 3; ++d) gc[d] += centroids[3 * i + d];
    const float center[3] = {gc[0] / n_cells, gc[1] / n_cells, gc[2] / n_cells};
    float volume = 0;
    for (int i = 0; i < n_cells; ++i) {
        float *cn = cell_normal + 3 * i;
        for (int j = cs_l[i]; j < cs_l[i + 1]; ++j) for (int d = 0; d < 3; ++d) cn[d] += nl[3 * j + d];
        normalize3(cn);
        const float dist[3] = {centroids[3 * i] - center[0], centroids[3 * i + 1] - center[1], centroids[3 * i + 2] - center[2]};
        if (dot3(cn, dist) < 0) { cn[0] = -cn[0]; cn[1] = -cn[1]; cn[2] = -cn[2]; }
        const float height = dot3(dist, cn);
        volume += height * (cs_l[i + 1] - cs_l[i]) * 3.1415926 * 1.26 / 4.0 / 3.0 * 1e-6;
    }
    const float f = strength * (target - volume) / target;
    for (int i = 0; i < n_cells; ++i) {
        const float *cn = cell_normal + 3 * i;
        for (int j = cs_l[i]; j < cs_l[i + 1]; ++j) for (int d = 0; d < 3; ++d) fl[3 * j + d] += f * cn[d] * ff->mass[0];
        if (cs_p) for (int j = cs_p[i]; j < cs_p[i + 1]; ++j) {
            const float m = ff->mass[(long)i < n_p ? type_p[i] : 0];
            for (int d = 0; d < 3; ++d) fp[3 * j + d] += f * cn[d] * m;
        }
    }
    return volume;
}

static int cmp_float(const void *a, const void *b) { float p = *(const float *)a, q = *(const float *)b; return p < q ? -1 : p > q; }

/* cleanup.h:29-60 */
long orc_delete_lipid_mask(int n_cells, const float *centroids, const int *cs_l, const float *xl, float tol, int *keep)
{
    long kept = 0;
    float *dr2 = NULL; int cap = 0;
    for (int i = 0; i < n_cells; ++i) {
        const int b = cs_l[i], e = cs_l[i + 1];
        if (e <= b) continue;
        if (e - b > cap) { cap = 2 * (e - b); dr2 = realloc(dr2, sizeof(float) * cap); }
        for (int j = b; j < e; ++j) dr2[j - b] = dist2(xl + 3 * j, centroids + 3 * i);
        qsort(dr2, e - b, sizeof(float), cmp_float);
        const float threshold = dr2[(e - b) / 2] * tol * tol;
        for (int j = b; j < e; ++j) { keep[j] = dist2(xl + 3 * j, centroids + 3 * i) < threshold ? 1 : 0; kept += keep[j]; }
    }
    free(dr2);
    return kept;
}

/* ============================================================================================
 * RNG
 * ============================================================================================ */
/* rng.h:107-137: MT19937 block generation with the reference's SIGNED intermediate y (int), i.e. an arithmetic
 * right shift and y % 2 on a possibly negative value — not the textbook generator. */
static void mt_twist(orc_mt19937 *g, int want_real)
{
    for (int i = 0; i < 624; i++) {
        int32_t y = (int32_t)((g->state[i] & 0x80000000u) + (g->state[(i + 1) % 624] & 0x7fffffffu));
        g->state[i] = g->state[(i + 397) % 624] ^ (uint32_t)(y >> 1);
        if (y % 2) g->state[i] ^= 0x9908b0dfu;
        uint32_t z = g->state[i];
        z ^= (z >> 11);
        z ^= ((z << 7) & 0x9d2c5680u);
        z ^= ((z << 15) & 0xefc60000u);
        z ^= (z >> 18);
        if (want_real) g->rdata[i] = z / (float)0xFFFFFFFF;
        else g->idata[i] = z;
    }
    if (want_real) g->rpos = 0; else g->ipos = 0;
}

/* rng.h:42-51 */
void orc_mt_init(orc_mt19937 *g, uint32_t seed)
{
    g->state[0] = seed;
    for (int i = 1; i < 624; ++i) g->state[i] = (1812433253u * (g->state[i - 1] ^ (g->state[i - 1] >> 30)) + (uint32_t)i);
    mt_twist(g, 0);
    mt_twist(g, 1);
}
uint32_t orc_mt_uint(orc_mt19937 *g) { if (g->ipos == 624) mt_twist(g, 0); return g->idata[g->ipos++]; }
float orc_mt_u01(orc_mt19937 *g) { if (g->rpos == 624) mt_twist(g, 1); return g->rdata[g->rpos++]; }

/* math_vector_integer.h:62-66 */
float orc_uint2u11(uint32_t u) { return u * 4.6566129e-10f - 1.f; }

/* integrate_langevin.h:116-137: 12 seeds then one 3-lane xorshift128 step per particle */
void orc_langevin_noise(orc_mt19937 *g, long n, float *noise)
{
    uint32_t x[3], y[3], z[3], w[3];
    for (int d = 0; d < 3; ++d) x[d] = orc_mt_uint(g);
    for (int d = 0; d < 3; ++d) y[d] = orc_mt_uint(g);
    for (int d = 0; d < 3; ++d) z[d] = orc_mt_uint(g);
    for (int d = 0; d < 3; ++d) w[d] = orc_mt_uint(g);
    for (long i = 0; i < n; ++i) for (int d = 0; d < 3; ++d) {
        uint32_t t = x[d];
        t ^= t << 11;
        t ^= t >> 8;
        x[d] = y[d]; y[d] = z[d]; z[d] = w[d];
        w[d] ^= w[d] >> 19;
        w[d] ^= t;
        noise[3 * i + d] = orc_uint2u11(w[d]);
    }
}

/* Philox4x32-10, Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11 */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox_noise(uint64_t seed, uint32_t step, uint32_t species, long n, float *noise)
{
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (long i = 0; i < n; ++i) {
        const uint32_t ctr[4] = {(uint32_t)i, step, species, 0u};
        uint32_t o[4]; orc_philox4x32_10(ctr, key, o);
        for (int d = 0; d < 3; ++d) noise[3 * i + d] = orc_uint2u11(o[d]);
    }
}
