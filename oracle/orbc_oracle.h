/* TEST INFRASTRUCTURE — CPU restatement ("port") of OpenRBC's per-timestep hot path.
 *
 * Plain C, scalar, one thread, strict IEEE fp32 (compiled with -fno-fast-math -ffp-contract=off).
 * Every function cites the reference file:line it restates.  The port is pinned against the
 * reference itself (oracle/_ref/libref_strict.so, the unmodified headers compiled here) by
 * tests/test_oracle_vs_ref.py and against the committed fixtures under tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this library; the
 * product (openrbc_b200/) never links, loads or calls it.
 *
 * Conventions: vectors are packed N x 3 float; indices are int32; "species" 0 = lipid, 1 = protein.
 */
#ifndef ORBC_ORACLE_H_
#define ORBC_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* forcefield_canonical.h:30-156.  Same field order as orbc_forcefield in include/orbc_b200.h. */
typedef struct {
    float mass[6], radius[6];
    float cutlp[6], cutsqlp[6], replp[6], attlp[6], alphalp[6];
    float cutpp[36], cutsqpp[36], reppp[36];
    float lj_cutsq[36], lj_lj1[36], lj_lj2[36];
    float r0[4], K[4];
    float cutll, cutsqll, repll, attll, alphall;
} orc_forcefield;

void orc_forcefield_canonical(orc_forcefield *ff);

/* ---- spatial index -------------------------------------------------------------------------- */
void     orc_update_centroid(int n_cells, const int *cell_start, const float *x, float *centroids); /* voronoi.h:123-140 */
uint32_t orc_morton_encode(float x, float y, float z);                                                /* reorder_morton.h:25-42 */
void     orc_morton_perm(int n, const float *pts, int *perm_new2old, uint32_t *keys_out);              /* reorder_morton.h:44-122 */
/* nearest centroid per point (voronoi.h:179-216 + kdtree.h:206-236).  Returns the number of points whose two best
 * candidates tie within 1 ulp of the sqrt-ed distance (where the reference's search order decides). */
int      orc_assign_nearest(long n, const float *x, int n_cells, const float *centroids, int *affiliation, int *tie_flag);
void     orc_partition(long n, int n_cells, const int *affiliation, int *cell_start, int *cells, int *local_index); /* voronoi.h:214-231 */
void     orc_gather3(long n, const int *cells, const float *src, float *dst);                        /* reorder.h:73-149 */
void     orc_gather1(long n, const int *cells, const int *src, int *dst);
/* {c2 : |c2-c1|^2 < r^2} ascending by id (voronoi.h:105-117, kdtree.h:263-285).  Returns count. */
int      orc_stencil(int n_cells, const float *centroids, int cell, float rmax, int *out, int cap);

/* ---- forces ----------------------------------------------------------------------------------- */
/* compute_pairwise_fused.h:238-320 at one thread (every cell in range => Newton on everywhere).  Accumulates. */
void orc_pairwise_fused(const orc_forcefield *ff, int n_cells, const float *centroids,
                        long n_l, const float *xl, const float *nl, const int *cs_l, float *fl, float *tl,
                        long n_p, const float *xp, const float *np_, const int *type_p, const int *cs_p, float *fp, float *tp,
                        long *counters /* [8] or NULL: LL cand, LL hit, PL cand, PL poly hit, PL lj hit, PP cand, PP poly hit, PP lj hit */);
/* compute_bonded.h:89-146.  bonds = (type, tag_i, tag_j); tag2idx maps tag -> index.  Accumulates into f. */
void orc_bonded(const orc_forcefield *ff, long n_bonds, const int *bonds, const int *tag2idx, const float *x, float *f);
void orc_build_tag2idx(long n, const int *tag, int *map, long map_size);                             /* container.h:39-58 */

/* ---- integrators (integrate_nh.h, integrate_langevin.h, openrbc.cpp:114-131) ------------------------ */
/* type == NULL means a lipid container (type 0 everywhere, container.h:128-130). */
void orc_clear_force(long n, float *f, float *t);                                                    /* integrate_nh.h:58-67 */
void orc_post_torque(long n, const float *nn, float *t);                                             /* :146-154 */
void orc_bounce_back(long n, float *x, float *v, double lo, double hi);                              /* :124-144 */
void orc_verlet_langevin(const orc_forcefield *ff, long n, float *x, float *v, float *f, float *nn, float *o, float *t,
                         const int *type, double dt, float eta, float kBT, const float *noise /* N x 3 in [-1,1) or NULL = 0 */); /* integrate_langevin.h:99-149 */
void orc_nh_initial_fused(const orc_forcefield *ff, long n, float *x, float *v, float *f, float *nn, float *o, float *t,
                          const int *type, double dt, float zeta, double lo, double hi, double *ke); /* integrate_nh.h:178-235 */
void orc_nh_final_fused(const orc_forcefield *ff, long n, float *v, const float *f, const float *nn, float *o, float *t,
                        const int *type, double dt, float zeta, double *ke);                         /* :237-273 */
void orc_nh_final(const orc_forcefield *ff, long n, float *v, const float *f, float *o, const float *t, const int *type, double dt, float zeta); /* :155-176 */
void orc_nh_update(const orc_forcefield *ff, long n, const float *v, const int *type, double *ke);  /* :78-89 */
float orc_nh_zeta_update_unfused(float zeta, float *Q, double dt, float kBT, double ke, int n);      /* :72-76 */
float orc_nh_zeta_update(float zeta, float *Q, double dt, float kBT, double ke, int n);              /* :72-76,181-185 */
void orc_opt_move(const orc_forcefield *ff, long n, float *x, float *nn, const float *f, const float *t, const int *type,
                  double dt, double dr_opt, double dn_opt);                                          /* openrbc.cpp:114-131 */
double orc_temperature(const orc_forcefield *ff, long n_l, const float *vl, long n_p, const float *vp, const int *type_p); /* compute_temperature.h:23-29 */
/* constrain_volume.h:26-83.  cell_normal (n_cells x 3) is the function-static scratch of the reference: it is NOT
 * cleared between calls; pass zeros for "first call on fresh memory".  Returns the volume estimate. */
float orc_constrain_volume(const orc_forcefield *ff, int n_cells, const float *centroids, float *cell_normal,
                           long n_l, const float *nl, const int *cs_l, float *fl,
                           long n_p, const int *type_p, const int *cs_p, float *fp, float target, float strength);
/* cleanup.h:29-91: keep[] mask; returns the number kept. */
long orc_delete_lipid_mask(int n_cells, const float *centroids, const int *cs_l, const float *xl, float stray_tolerance, int *keep);

/* ---- RNG (rng.h:29-139, integrate_langevin.h:116-137, math_vector_integer.h:62-66) ------------------- */
typedef struct {
    uint32_t idata[624];
    float    rdata[624];
    uint32_t state[624];
    int      ipos, rpos;
} orc_mt19937;
void     orc_mt_init(orc_mt19937 *g, uint32_t seed);
uint32_t orc_mt_uint(orc_mt19937 *g);
float    orc_mt_u01(orc_mt19937 *g);
float    orc_uint2u11(uint32_t u);
/* the noise vectors verlet_langevin draws for one container slice of n particles from generator g */
void     orc_langevin_noise(orc_mt19937 *g, long n, float *noise);
/* Philox4x32-10 (Salmon et al., SC'11) — the counter-based generator of the CUDA path */
void     orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* noise of the CUDA path: counter = (index, step, species, 0), key = (seed_lo, seed_hi) -> uint2u11 of out[0..2] */
void     orc_philox_noise(uint64_t seed, uint32_t step, uint32_t species, long n, float *noise);

#ifdef __cplusplus
}
#endif
#endif
