/* orbc_b200.h — C ABI of the B200-native replacement for OpenRBC's per-timestep hot path.
 *
 * The reference (yhtang/OpenRBC) has no plugin / FFI layer: its "operator interface" for this path is the set of
 * C++ free functions and functor kernels that src/openrbc.cpp calls from its optimisation and main loops.  Each entry
 * point below replaces one of those call sites; the reference file:line is given next to it.  A maintainer keeps
 * openrbc.cpp, runtime_parameter.h, init_*.h, topology.h and trajectory.h as they are and forwards the hot-path
 * functions through openrbc_b200/host/orbc_shim.h.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types; one opaque context per GPU; single host thread per context.
 *   - every function returns 0 on success and a negative orbc_status otherwise; orbc_last_error() gives the text
 *     (the reference's own error behaviour is exit(0)/SIGSEGV, util_misc.h:54-59, reorder.h:103).
 *   - host arrays are borrowed for the duration of a call; the library owns all device memory.
 *   - vectors cross the boundary as `stride_floats` floats per particle (3 for the default build's vector<float,3>,
 *     4 under _ESIMD/_VEC4; config_static.h:36-44); only the first three are read / written.
 *   - species: 0 = lipid container, 1 = protein container (container.h:117-157).
 *   - there is no CPU fallback: without a CUDA device orbc_create() fails.
 */
#ifndef ORBC_B200_H_
#define ORBC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBC_API __attribute__((visibility("default")))

typedef struct orbc_ctx orbc_ctx;

typedef enum {
    ORBC_OK = 0,
    ORBC_ERR_CUDA = -1,        /* a CUDA runtime call failed */
    ORBC_ERR_ARG = -2,         /* bad argument / call order */
    ORBC_ERR_STATE = -3,       /* device-side consistency check failed (stencil overflow, lost particle, ...) */
    ORBC_ERR_NO_DEVICE = -4
} orbc_status;

enum { ORBC_LIPID = 0, ORBC_PROTEIN = 1 };

/* Byte image of the reference's ForceField tables (forcefield_canonical.h:30-156), n_type = 6, n_bondtype = 4.
 * orbc_forcefield_canonical() fills the same values the reference's constexpr / static-init code produces. */
typedef struct {
    float mass[6], radius[6];
    float cutlp[6], cutsqlp[6], replp[6], attlp[6], alphalp[6];
    float cutpp[36], cutsqpp[36], reppp[36];
    float lj_cutsq[36], lj_lj1[36], lj_lj2[36];
    float r0[4], K[4];
    float cutll, cutsqll, repll, attll, alphall;
} orbc_forcefield;

/* Functor kernels of integrate() — integrate_nh.h:29-55 driver, kernels at the cited lines. */
typedef enum {
    ORBC_CLEAR_FORCE = 0,        /* clear_force                               integrate_nh.h:58-67   */
    ORBC_POST_TORQUE = 1,        /* post_torque                               integrate_nh.h:146-154 */
    ORBC_BOUNCE_BACK = 2,        /* bounce_back                               integrate_nh.h:124-144 */
    ORBC_VERLET_LANGEVIN = 3,    /* verlet_langevin                           integrate_langevin.h:99-149 */
    ORBC_NH_INITIAL_FUSED = 4,   /* verlet_initial_bounce_clearforce_update   integrate_nh.h:178-235 */
    ORBC_NH_FINAL_FUSED = 5,     /* post_toque_final_update                   integrate_nh.h:237-273 */
    ORBC_NH_FINAL = 6,           /* verlet_nh_final                           integrate_nh.h:156-176 */
    ORBC_NH_UPDATE = 7,          /* verlet_nh_update (KE only)                integrate_nh.h:69-94   */
    ORBC_OPT_MOVE = 9,           /* steepest-descent mover                    openrbc.cpp:114-131    */
    ORBC_OPT_FUSED = 10          /* post_torque + mover + bounce_back in one pass: openrbc.cpp:110-133 (f is kept, t becomes n x t) */
} orbc_integrator;

/* The RTParameter fields the kernels read (runtime_parameter.h:38-80). */
typedef struct {
    double dt;                  /* RTParameter::dt */
    float  kBT, eta;            /* RTParameter::kBT, eta */
    float  zeta;                /* RTParameter::zeta (Nose-Hoover friction), read only */
    double box_lo, box_hi;      /* RTParameter::box (same on the three axes: -1000 / 1000) */
    double dr_opt, dn_opt;      /* RTParameter::dr_opt, dn_opt (ORBC_OPT_MOVE) */
    int    nstep;               /* RTParameter::nstep — also the RNG counter */
    uint64_t seed;              /* RTParameter::rseed; keys the counter-based generator */
    /* test-only noise injection for ORBC_VERLET_LANGEVIN: host arrays of n x 3 floats in [-1,1) that replace the
     * generator (lets a Langevin step be compared with the reference's xorshift stream exactly); NULL in production */
    const float *noise_lipid, *noise_protein;
} orbc_step_params;

typedef struct {
    double ke;                  /* sum 1/2 m v^2 over both containers (fp64), for the kernels that reduce it */
    long   n;                   /* particles visited */
} orbc_step_result;

/* ---- lifetime ----------------------------------------------------------------------------------------- */
ORBC_API int  orbc_create(orbc_ctx **ctx, int device);
ORBC_API void orbc_destroy(orbc_ctx *ctx);
ORBC_API const char *orbc_last_error(void);
ORBC_API int  orbc_synchronize(orbc_ctx *ctx);
/* adopt an existing CUDA stream (cudaStream_t passed as void*); NULL = the context's own stream */
ORBC_API int  orbc_set_stream(orbc_ctx *ctx, void *cuda_stream);

/* tuning / test switches, by name.
 *   "pair_impl"   2 = production kernels (default), 1 = the simple thread-per-particle kernels kept as an independent cross-check
 *   "ll_variant"  lipid-lipid kernel: 1 = thread per lipid over candidate runs, with hit lists (default); 0 = warp-per-cell tile kernel
 *   "nl_reuse"    1 (default) = hit lists: a force evaluation after a rebuild may record, per particle, every candidate closer than
 *                 cutoff + skin, and the evaluations up to the next rebuild walk those lists (exact re-test of every entry; a bound on
 *                 the displacements guards the skin: same hits, same forces).  The device decides at every evaluation whether it
 *                 walks, records or just searches (a recording that would not be walked is not made).  0 = every evaluation
 *                 searches the stencils.  On a decomposed context the ranks exchange their displacement bounds and decide alike;
 *                 there the default means "up to four ranks" (measured: the gain shrinks with the rank's share of the work), 2 = on for every world size.
 *   "nl_skin", "nl_skin_max"  the skin of those lists: at least nl_skin (default 0.1); the gate thickens it up to nl_skin_max (default 0.3)
 *                 when the fastest particle of the last step would outrun the thinner one
 *   "stencil_refresh"  1 (default) = rebuilds that keep the cell numbering re-classify the recorded r < 9 + 1 neighbours of every cell
 *                 instead of searching the centroid grid (cells whose centroid jumped are searched in full; same stencils); 0 = always search
 *   "prot_lanes"  lanes per protein of the protein kernel (0 = automatic); debug_*: test aids */
ORBC_API int  orbc_set_option(orbc_ctx *ctx, const char *name, double value);

/* ---- force field (forcefield_canonical.h) ------------------------------------------------------------------ */
ORBC_API int  orbc_forcefield_canonical(orbc_forcefield *ff);
ORBC_API int  orbc_set_forcefield(orbc_ctx *ctx, const orbc_forcefield *ff);

/* ---- state upload: the containers filled by init_rbc()/init_random_sphere() (openrbc.cpp:55-65) ----------------- */
/* type/tag are NULL for lipids (implicit 0 and base+i, container.h:122-130). f and t start at zero. */
ORBC_API int  orbc_upload(orbc_ctx *ctx, int species, size_t n, size_t stride_floats,
                          const float *x, const float *v, const float *n_, const float *o, const int *type, const int *tag);
/* the same for a rank of a CONNECTED decomposed run that re-uploads the system it already holds (a restart, the next job of a
 * sweep): only rows [first, first + count) of the vector arrays travel over PCIe — the slots this rank owns, cell_start[cell_begin] ..
 * cell_start[cell_end] of orbc_mg_cell_range — while type / tag are taken for every slot.  The halo copies are then fetched
 * from their owners over NVLink by the orbc_mg_export that must follow on every rank.  The array pointers are those of the
 * whole containers (row 0). */
ORBC_API int  orbc_upload_range(orbc_ctx *ctx, int species, size_t n, size_t first, size_t count, size_t stride_floats,
                                const float *x, const float *v, const float *n_, const float *o, const int *type, const int *tag);
/* Bond[] as (type, tag_i, tag_j) triples, container.h:30-34 */
ORBC_API int  orbc_upload_bonds(orbc_ctx *ctx, size_t n_bonds, const int *type_i_j);
/* VoronoiDiagram::centroids + VCellList::cell_start of both containers after voronoi.init() (openrbc.cpp:69-74).
 * Particles must already be stored sorted by cell.  cell_start_* may be NULL (no previous partition known). */
ORBC_API int  orbc_voronoi_upload(orbc_ctx *ctx, int n_cells, const float *centroids3, const int *cell_start_l, const int *cell_start_p);
/* VoronoiDiagram::init(lipid, cell_lipid, param, n_iterate) (voronoi.h:54-75, openrbc.cpp:71) on the device, for hosts that do
 * not want to spend the reference's seconds of k-means on the CPU: needs the lipids uploaded (any order), leaves them sorted by
 * cell with centroids, cell_start and the index in place; follow with orbc_cell_update(ORBC_PROTEIN).  Single GPU. */
ORBC_API int  orbc_voronoi_init(orbc_ctx *ctx, int n_cells, int n_iterate);
/* overwrite f / t / v of one container (tests, and hosts that compute extra forces on the CPU); field: 'f','t','v','x','n','o' */
ORBC_API int  orbc_set_field(orbc_ctx *ctx, int species, char field, size_t stride_floats, const float *src);

/* ---- spatial index ---------------------------------------------------------------------------------------- */
/* VoronoiDiagram::update (voronoi.h:77-86): centroids from the previous partition, Morton sort of the centroids when
 * nstep % freq_sort_ctrd == 0 (reorder_morton.h:44-122), neighbour structure rebuild (replaces kdtree.h:116). */
ORBC_API int  orbc_voronoi_update(orbc_ctx *ctx, int nstep, int freq_sort_ctrd);
/* VCellList::update (voronoi.h:153-163): nearest-centroid partition (voronoi.h:179-237), reorder (reorder.h:73-149),
 * tag->index map (container.h:39-58).  reorder_bond (reorder.h:33-68) only changes bond storage order and is a no-op here. */
ORBC_API int  orbc_cell_update(orbc_ctx *ctx, int species, int nstep, int freq_sort_bond);
/* the three calls of openrbc.cpp:202-204 in one */
ORBC_API int  orbc_rebuild(orbc_ctx *ctx, int nstep, int freq_sort_ctrd, int freq_sort_bond);
/* delete_lipid (cleanup.h:29-91) */
ORBC_API int  orbc_delete_lipid(orbc_ctx *ctx, float stray_tolerance, size_t *n_lipid_out);

/* ---- forces ----------------------------------------------------------------------------------------------- */
ORBC_API int  orbc_compute_pairwise_fused(orbc_ctx *ctx);   /* compute_pairwise_fused.h:238-320; accumulates into f, t */
ORBC_API int  orbc_compute_bonded(orbc_ctx *ctx);           /* compute_bonded.h:89-146; accumulates into protein f */
ORBC_API int  orbc_constrain_volume(orbc_ctx *ctx, float target_volume, float strength, float *volume_out); /* constrain_volume.h:26-83 */
/* switch for the (commented-out) call of openrbc.cpp:229: when on, orbc_run_langevin / orbc_run_nh apply constrain_volume(target,
 * strength) between compute_bonded and the integrator of every step (BASELINE configs[2]: 3.15, 0.05) */
ORBC_API int  orbc_set_volume_constraint(orbc_ctx *ctx, int on, float target_volume, float strength);

/* ---- integrators ------------------------------------------------------------------------------------------- */
/* integrate(KERNEL&&, lipid, protein) — integrate_nh.h:29-37.  The Nose-Hoover kernels return (ke, n); the zeta update
 * of their destructors (integrate_nh.h:181-185) stays on the host, see orbc_nh_zeta_update(). */
ORBC_API int  orbc_integrate(orbc_ctx *ctx, int kernel, const orbc_step_params *p, orbc_step_result *res);
ORBC_API float orbc_nh_zeta_update(float zeta, float *Q, double dt, float kBT, double ke, long n);
/* same for the unfused verlet_nh_update (integrate_nh.h:72-76), whose target kinetic energy 1.5 n kBT is a product of floats */
ORBC_API float orbc_nh_zeta_update_unfused(float zeta, float *Q, double dt, float kBT, double ke, long n);
ORBC_API int  orbc_compute_temperature(orbc_ctx *ctx, double *temperature);   /* compute_temperature.h:23-29 */

/* ---- whole-loop entry points (what the GPU build of openrbc.cpp's while-loop calls; openrbc.cpp:189-256) ------- */
/* n_steps of: [rebuild if nstep % freq_voronoi == 0] -> pair forces -> bonded -> verlet_langevin, with no host
 * synchronisation inside.  p->nstep is the first step's index. */
ORBC_API int  orbc_run_langevin(orbc_ctx *ctx, const orbc_step_params *p, int n_steps, int freq_voronoi, int freq_sort_ctrd);
/* the energy-minimisation loop, openrbc.cpp:88-133: n_steps of rebuild (Morton sort included: param.nstep stays 0 there) ->
 * clear_force -> pair forces -> bonded -> post_torque -> capped steepest-descent move -> bounce_back, the last three as one
 * kernel.  Uses p->dt, dr_opt, dn_opt, box.  On return f holds the last forces and t the last n x t, as in the reference. */
ORBC_API int  orbc_run_minimize(orbc_ctx *ctx, const orbc_step_params *p, int n_steps, int freq_sort_ctrd);
/* same for the Nose-Hoover build (openrbc.cpp:192-241); zeta/Q are updated on the device and returned */
ORBC_API int  orbc_run_nh(orbc_ctx *ctx, const orbc_step_params *p, int n_steps, int freq_voronoi, int freq_sort_ctrd, float *zeta_inout, float *Q_inout);

/* ---- one cell over the GPUs of a box (no counterpart in the reference, which is one shared-memory process; the model is
 * its own thread decomposition: contiguous ranges of Morton-ordered cells per worker, util_numa.h:41-42, cross-range cell
 * pairs evaluated one-sidedly by both owners, compute_pairwise_fused.h:264-275,287-295,303-314) -------------------------------
 * Call order on every rank (one process per GPU, or several contexts in one process): orbc_create, orbc_mg_init,
 * orbc_upload / orbc_upload_bonds / orbc_voronoi_upload with the WHOLE system, orbc_mg_export, exchange the blobs
 * (any transport: torch.distributed, MPI, a file), orbc_mg_connect with all of them in rank order.  After that
 * orbc_rebuild / orbc_compute_pairwise_fused / orbc_compute_bonded / orbc_integrate(VERLET_LANGEVIN) / orbc_run_langevin work on
 * the rank's own cells; halo copies, particle migration and the synchronisation between ranks happen on the device over
 * NVLink (peer stores and epoch flags), with no host synchronisation and no collective library on the data path.
 * A fresh orbc_upload of the same system on connected ranks is followed by orbc_mg_export again (no second connect); that call
 * is collective: it returns when every rank has exported.
 * orbc_constrain_volume, orbc_integrate(NH_*_FUSED / OPT_FUSED), orbc_run_nh, orbc_run_minimize and orbc_delete_lipid are decomposed
 * the same way.  Every rank must issue the same sequence of calls. */
ORBC_API int  orbc_mg_init(orbc_ctx *ctx, int rank, int world /* <= 8 */);
ORBC_API size_t orbc_mg_blob_bytes(void);
/* cells [begin, end) that rank `rank` of `world` owns: the reference's static range partition, util_numa.h:41-42 (host-only helper) */
ORBC_API int  orbc_mg_cell_range(int n_cells, int rank, int world, int *cell_begin, int *cell_end);
ORBC_API int  orbc_mg_export(orbc_ctx *ctx, void *blob_out, size_t bytes);
ORBC_API int  orbc_mg_connect(orbc_ctx *ctx, const void *blobs /* world x bytes_each, rank order */, size_t bytes_each);
/* particle slots [begin, end) of one container this rank owns right now (synchronises); orbc_download returns whole
 * containers of which only these slots (and the halo) are current */
ORBC_API int  orbc_mg_range(orbc_ctx *ctx, int species, size_t *begin, size_t *end);

/* ---- download: what save_frame()/display() need (trajectory.h:61-105, openrbc.cpp:248-254) ------------------------ */
/* any destination may be NULL.  affiliation = VCellList::update_particle_affiliation (voronoi.h:166-175). */
ORBC_API int  orbc_download(orbc_ctx *ctx, int species, size_t stride_floats,
                            float *x, float *v, float *n_, float *o, float *f, float *t,
                            int *affiliation, int *type, int *tag, size_t *n);
/* save_frame (trajectory.h:61-105) with update_particle_affiliation (voronoi.h:166-175) folded in: the frame is assembled on
 * the device in the .orbc byte layout (FRAMEBEG nstep NATOM n IDENTITY ... FRAMEEND, lipids before proteins, the sections that
 * dump_field selects: DumpField bits of runtime_parameter.h:30-36) and leaves the GPU as ONE copy; the host appends the bytes
 * to its trajectory stream.  lipid_tag_base = LipidContainer::tag.base (container.h:122-126, default 1).
 *   orbc_save_frame        synchronous, into the caller's buffer (orbc_frame_bytes tells its size)
 *   orbc_save_frame_begin  packs on the context's stream and starts the copy into a pinned buffer of the library on a second
 *                          stream; the run continues meanwhile.  At most two frames in flight.
 *   orbc_save_frame_end    waits for the oldest frame in flight; *data stays valid until the next-but-one _begin.
 * Decomposed run: a rank's image holds the titles and the slots it owns, zeros elsewhere (OR the images of all ranks). */
ORBC_API int  orbc_frame_bytes(orbc_ctx *ctx, int dump_field, size_t *bytes);
ORBC_API int  orbc_save_frame(orbc_ctx *ctx, int nstep, int dump_field, int lipid_tag_base, void *dst, size_t cap, size_t *bytes);
ORBC_API int  orbc_save_frame_begin(orbc_ctx *ctx, int nstep, int dump_field, int lipid_tag_base);
ORBC_API int  orbc_save_frame_end(orbc_ctx *ctx, const void **data, size_t *bytes);
ORBC_API int  orbc_size(orbc_ctx *ctx, int species, size_t *n);
ORBC_API int  orbc_n_cells(orbc_ctx *ctx, int *n_cells);

/* ---- introspection for tests ---------------------------------------------------------------------------------- */
typedef enum {
    ORBC_DUMP_CENTROIDS = 0,     /* n_cells x 3 float */
    ORBC_DUMP_CELL_START_L = 1,  /* n_cells + 1 int */
    ORBC_DUMP_CELL_START_P = 2,
    ORBC_DUMP_CELLS_L = 3,       /* permutation applied by the last reorder, n int */
    ORBC_DUMP_CELLS_P = 4,
    ORBC_DUMP_AFF_L = 5,         /* affiliation in pre-reorder order, n int (voronoi.h:212) */
    ORBC_DUMP_AFF_P = 6,
    ORBC_DUMP_MORTON_KEYS = 7,   /* keys of the current centroids, n_cells uint32 */
    ORBC_DUMP_MORTON_PERM = 8,   /* new -> old permutation of the last centroid sort, n_cells int */
    ORBC_DUMP_STENCIL_COUNTS = 9,/* n_cells x 3 int: entries with centroid distance < 6, < 8, < 9 */
    ORBC_DUMP_STENCIL = 10,      /* n_cells x ORBC_STENCIL_STRIDE int, ordered (class, id) */
    ORBC_DUMP_TAG2IDX = 11,      /* protein tag -> index, (max_tag + 1) int */
    ORBC_DUMP_COUNTERS = 12,     /* 8 x uint64 device counters: [0] nearest-centroid fallback searches, [1..6] pair tests / hits of the pair_impl 1 kernels,
                                  * [7] stencil refresh: cells searched in full (low 40 bits), refreshes redone by a full search (high bits) */
    ORBC_DUMP_NL_STATS = 13      /* 4 x uint32: force evaluations that recorded the hit lists, that walked them, overflow flag, that searched without recording */
} orbc_dump;
#define ORBC_STENCIL_STRIDE 64
ORBC_API int  orbc_debug_dump(orbc_ctx *ctx, int what, void *dst, size_t bytes);
/* the generator of the CUDA path, evaluated on the device: noise for (seed, nstep, species) as n x 3 floats */
ORBC_API int  orbc_debug_noise(orbc_ctx *ctx, uint64_t seed, int nstep, int species, size_t n, float *dst);

/* ---- timing on the context's stream (CUDA events) ----------------------------------------------------------------- */
ORBC_API int  orbc_event_record(orbc_ctx *ctx, int slot /* 0..15 */);
ORBC_API int  orbc_event_elapsed_ms(orbc_ctx *ctx, int slot_a, int slot_b, float *ms);
/* per-kernel-class device time, measured with CUDA event pairs recorded on the context's stream around every launch
 * of the class while profiling is enabled (bench.py's roofline figures come from here).  orbc_profile_read synchronises,
 * returns the summed duration and launch count since the last read of that class, and resets them. */
typedef enum {
    ORBC_PROF_PAIR_LIPID = 0,    /* lipid side of compute_pairwise_fused (LL + lipid side of protein-lipid) */
    ORBC_PROF_PAIR_PROTEIN = 1,  /* protein side (PP + protein side of protein-lipid) */
    ORBC_PROF_BONDED = 2,
    ORBC_PROF_INTEGRATE = 3,
    ORBC_PROF_REBUILD = 4,       /* voronoi.update + both cell updates, all kernels together */
    ORBC_PROF_N = 5
} orbc_prof_class;
ORBC_API int  orbc_profile_enable(orbc_ctx *ctx, int on);
ORBC_API int  orbc_profile_read(orbc_ctx *ctx, int cls, double *total_ms, unsigned long long *count);
/* per-kernel launch list: while enabled every launch is bracketed by a CUDA event pair on the context's stream; the report is
 * one line per kernel, "name launches total_us", sorted by total time, and resets the list (also usable on decomposed runs,
 * where a replaying profiler cannot be: the ranks wait for each other inside kernels) */
ORBC_API int  orbc_profile_kernels(orbc_ctx *ctx, int on);
ORBC_API int  orbc_profile_kernels_report(orbc_ctx *ctx, char *text, size_t bytes);
/* number of kernels this context has launched since creation */
ORBC_API int  orbc_launch_count(orbc_ctx *ctx, unsigned long long *n);

#ifdef __cplusplus
}
#endif
#endif
