// common.cuh — context, error handling and small device helpers shared by every kernel file.
#pragma once

#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/orbc_b200.h"

namespace orbc {

// ------------------------------------------------------------------------------------------------
// errors: thread-local message, negative status codes (include/orbc_b200.h)
// ------------------------------------------------------------------------------------------------
inline char *err_buf() { static thread_local char buf[512] = "no error"; return buf; }
inline int fail(int code, const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(err_buf(), 512, fmt, ap); va_end(ap);
    return code;
}
#define ORBC_CUDA(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return ::orbc::fail(ORBC_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e__)); } while (0)
#define ORBC_TRY(x) do { int r__ = (x); if (r__ != ORBC_OK) return r__; } while (0)

constexpr int kNType = 6;              // forcefield_canonical.h:37
constexpr int kStencilStride = ORBC_STENCIL_STRIDE;
constexpr int kMoversCap = 4096;
constexpr int kNlAutoWorld = 4;      // hit lists on a decomposed run: automatic up to this many ranks (orbc_b200.cu: nl_active)
constexpr float kBin = 10.0f;          // centroid grid bin = largest centroid stencil radius (9, compute_pairwise_fused.h:183) + the margin of the wide stencils (rebuild.cuh)
constexpr int kMaxWorld = 8;           // ranks of one spatially decomposed run (one B200 box)

// device image of the force field + derived Langevin coefficients, in __constant__ memory
struct DevForceField {
    orbc_forcefield ff;
};

// one container (container.h:65-157) on the device.  SoA of float4:
//   x  = (x, y, z, type as int bits)      n = (nx, ny, nz, tag as int bits)
//   v, o, f, t = (.., .., .., unused)
// x, n, v, o are double-buffered for the out-of-place reorder (reorder.h:73-149), like the reference's shadow arrays.
struct Species {
    size_t n = 0, cap = 0;
    int cur = 0;                          // current buffer of v, o, cellid
    int cur_xn = 0;                       // current buffer of x, n (a decomposed run also flips it at every integration step)
    float4 *x[2] = {nullptr, nullptr}, *nn[2] = {nullptr, nullptr}, *v[2] = {nullptr, nullptr}, *o[2] = {nullptr, nullptr};
    float4 *f = nullptr, *t = nullptr;
    int *cellid[2] = {nullptr, nullptr};  // cell of every particle in storage order (-1: unknown), double-buffered with x
    int *aff = nullptr;                   // nearest centroid per particle, pre-reorder order (VCellList::affiliation)
    int *li = nullptr;                    // arrival slot inside the new cell (VCellList::local_index, order fixed later)
    int *cells = nullptr, *cells_tmp = nullptr;  // gather permutation (VCellList::cells)
    int *cell_start = nullptr;            // n_cells + 1 (VCellList::cell_start)
    bool has_partition = false;
    float4 *X() const { return x[cur_xn]; }
    float4 *N() const { return nn[cur_xn]; }
    float4 *V() const { return v[cur]; }
    float4 *O() const { return o[cur]; }
    int *C() const { return cellid[cur]; }
};

struct Grid {                // uniform grid over the centroids (replaces the k-d tree, kdtree.h)
    float lo[3] = {0, 0, 0};
    float h = kBin;
    int dim[3] = {1, 1, 1};
    int nbins = 1;
    int *bin_start = nullptr;   // nbins + 1
    int *bin_items = nullptr;   // n_cells
    int *bin_of = nullptr;      // n_cells
    int *bin_slot = nullptr;    // n_cells
    float4 *sorted = nullptr;   // n_cells: centroids in bin order, id in .w
    size_t cap_bins = 0;
};

// ---- spatial decomposition over the GPUs of one box (multi.cuh) -----------------------------------------------------------
// Every rank keeps the containers in the SAME global index space (particles sorted by Voronoi cell, cells in Morton order)
// and owns a contiguous range of cells [cb, ce) together with the particle slots of those cells; only the owned slots and
// the halo (members of cells inside the r<9 centroid stencil of an owned cell, and bonded partners) are kept up to date.
// Peers write into each other's arrays directly over NVLink (peer-mapped pointers); PeerTable holds those pointers.
struct PeerTable {
    float4 *x[2][2][kMaxWorld], *nn[2][2][kMaxWorld], *v[2][2][kMaxWorld], *o[2][2][kMaxWorld];   // [species][buffer][rank]
    int *cellid[2][2][kMaxWorld];
    float4 *centroid[2][kMaxWorld];       // [parity][rank]
    int *cnt_all[2][kMaxWorld];           // [species][rank]: world x (n_cells + 1) arrival counts
    int *tag2idx[kMaxWorld];
    unsigned *flags[kMaxWorld];           // barrier epochs, one slot per writer
    double *ke_all[kMaxWorld];            // partial kinetic energies, one slot per writer
    double *vol_all[kMaxWorld];           // partial volumes of constrain_volume, one slot per writer
    int *cv_ptype[kMaxWorld];             // type of the protein in slot c, c < n_cells (constrain_volume.h:73 indexes it by CELL)
};
struct Decomp {
    bool on = false, connected = false;
    bool partial[2] = {false, false};     // the last upload of the container held only this rank's own rows (orbc_upload_range)
    int rank = 0, world = 1;
    int cb = 0, ce = 0;                   // owned cells
    size_t own_cap[2] = {0, 0};           // launch bound of the owned-particle kernels
    unsigned epoch = 0;                   // barriers issued so far
    int need_epoch = 0;                   // rebuild counter marking owned + halo cells in `need`
    int cen_par = 0;                      // which of the two centroid buffers is current (they swap on rebuilds)
    float4 *cen_buf[2] = {nullptr, nullptr};
    unsigned *flags = nullptr;            // kMaxWorld epochs written by the peers
    double *ke_all = nullptr;             // 2 x kMaxWorld partial kinetic energies written by the peers, halves used alternately
    int ke_par = 0;
    int disp_par = 0;                     // which of the two sets of displacement bounds (behind the barrier epochs in `flags`) is current
    double *vol_all = nullptr;            // 2 x kMaxWorld partial volumes written by the peers (constrain_volume), halves used alternately
    int cv_par = 0;
    int *cv_ptype = nullptr;              // n_cells protein types by slot, written by the slots' owners (constrain_volume)
    int *cnt_all[2] = {nullptr, nullptr}; // world rows of arrival counts per species (row r written by rank r)
    int *off_me[2] = {nullptr, nullptr};  // members of each cell that come from lower ranks
    int *cnt_prev[2] = {nullptr, nullptr};// this rank's counts at the previous exchange (which entries the peers hold non-zero)
    unsigned char *dest_mask = nullptr;   // per cell: ranks (other than the owner) that own a cell of its r<9 stencil
    unsigned char *pmask = nullptr;       // per protein slot: ranks that own a bonded partner
    int *need = nullptr;                  // per cell: == need_epoch for owned and halo cells
    int *keep = nullptr;                  // per lipid slot: survives delete_lipid
    int *my_bonds = nullptr; int my_bonds_cap = 0;   // [0] = count, then the bonds with an owned atom
    PeerTable peers;
    std::vector<void *> opened;           // IPC mappings to close
};

} // namespace orbc

struct orbc_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    orbc::Species sp[2];
    // Voronoi diagram
    int n_cells = 0;
    float4 *centroid = nullptr, *centroid_tmp = nullptr;
    uint32_t *keys = nullptr, *keys_tmp = nullptr;
    int *perm = nullptr, *perm_tmp = nullptr;     // new -> old of the last Morton sort
    int *inv = nullptr;                           // old -> new
    bool inv_identity = true;
    orbc::Grid grid;
    int *stencil = nullptr;                       // n_cells x kStencilStride
    int *stencil_cnt = nullptr;                   // n_cells, packed n6 | n8 << 8 | n9 << 16
    bool stencil_valid = false;
    int *wide = nullptr, *wide_cnt = nullptr; float4 *cen_ref = nullptr;   // wide stencils (r < 9 + margin) and the centroids they were recorded at (rebuild.cuh)
    bool wide_on = true;                          // option stencil_refresh
    int *movers = nullptr;                        // cells whose centroid has outrun the margin, and cells a mover has newly reached (2 x kMoversCap; counts in wide_ok[1], [2])
    int *wide_ok = nullptr;                       // device flag: the refreshed stencils can be trusted (no mover was missed, the list of movers did not overflow)
    bool wide_valid = false;                      // the wide stencils match the current numbering of the cells (host's view)
    float4 *cell_normal = nullptr;                // constrain_volume's persistent scratch
    float4 *lbound = nullptr, *pbound = nullptr;  // per-cell bounding spheres of the current lipids / proteins (pair_queue.cuh)
    int2 *lruns = nullptr; int *lrun_cnt = nullptr; size_t lruns_cells = 0;   // merged candidate runs of every cell's r<6 stencil + packed per-cell info (k_lipid_runs)
    int tile_cap = 256;                           // candidates a tile holds (kTileCap; smaller only under the test option debug_tile_cap)
    int *tile_overflow = nullptr;                 // raised by k_lipid_runs when a cell's candidates do not fit the tile of k_pair_ll_t
    bool lruns_valid = false;
    int *porder = nullptr; size_t porder_cap = 0;  // thread -> protein map of the protein pair kernel (heavy types first)
    bool porder_valid = false;
    unsigned type_mask = 0;                       // protein types present (bit t), from the last protein upload
    // bonds
    size_t n_bonds = 0;
    int *bonds = nullptr; size_t bonds_cap = 0;   // (type, tag_i, tag_j)
    int *tag2idx = nullptr; size_t tag2idx_size = 0;
    // scratch
    int *scan_tmp = nullptr; size_t scan_tmp_cap = 0; unsigned scan_epoch = 0;   // tile descriptors of the single-pass scan
    int *radix_hist = nullptr; size_t radix_hist_cap = 0;
    float *stage = nullptr; size_t stage_cap = 0;         // device staging for strided host<->device packing
    double *d_acc = nullptr;                              // 8 doubles: reductions
    unsigned long long *d_counters = nullptr;             // 8 counters
    int *d_flags = nullptr;                               // 4 ints: device-side error flags
    int *d_check = nullptr, *h_check = nullptr;           // 8 ints: what the upload checks found (k_check_ids, k_check_bonds), and the pinned mirror
    float *d_nh = nullptr;                                // zeta, Q on the device for orbc_run_nh
    double *h_acc = nullptr; int *h_flags = nullptr; unsigned long long *h_counters = nullptr; float *h_nh = nullptr;  // pinned mirrors
    float *noise[2] = {nullptr, nullptr}; size_t noise_cap[2] = {0, 0};
    cudaEvent_t ev[16] = {};
    unsigned long long launches = 0;
    bool ff_set = false;
    orbc_forcefield host_ff;                       // host copy of this context's force field (Langevin coefficients, cull radii)
    int pair_impl = 2;
    int ll_variant = 1;                            // 1: thread-per-lipid run-list kernel k_pair_ll_r + hit lists (default); 0: warp-per-cell tile kernel k_pair_ll_t
    // hit lists with a skin (pair_queue.cuh): recorded by the force evaluation after a rebuild, walked until the next one
    int nl_on = 1;                                 // option "nl_reuse": 0 off, 1 automatic, 2 on
    bool nl_valid = false;                         // lists match the current partition and were built (host's view)
    float nl_skin = 0.1f, nl_skin_max = 0.3f;      // options "nl_skin", "nl_skin_max": the gate picks the skin of a recording between them
    int nl_moves = 0;                              // tracked integration steps since the last gate
    void *nl_state = nullptr;                      // orbc::NlState on the device
    int *ll_list = nullptr, *ll_cnt = nullptr; size_t ll_list_lipids = 0;
    int *pl_list = nullptr, *pl_cnt = nullptr, *pp_list = nullptr, *pp_cnt = nullptr; size_t pl_list_proteins = 0;
    int nl_cap_ll = 96, nl_cap_pl = 64, nl_cap_pp = 32;   // entries per lipid / per protein (lipid partners, protein partners)
    int prot_lanes = 0;                            // lanes per protein in k_pair_prot (0 = by the number of owned proteins)
    int *d_range = nullptr;                               // {l0, l1, p0, p1}: particle slots this context computes (all of them on one GPU)
    // volume constraint inside the whole-loop entry points (openrbc.cpp:229)
    bool cv_on = false; float cv_target = 0.f, cv_strength = 0.f;
    // save_frame: frame image assembled on the device, copied out on a second stream into two pinned buffers
    unsigned char *frame_dev[2] = {nullptr, nullptr}; size_t frame_dev_cap[2] = {0, 0};
    unsigned char *frame_host[2] = {nullptr, nullptr}; size_t frame_host_cap[2] = {0, 0};
    size_t frame_bytes[2] = {0, 0};
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t xfer_ev[8] = {};                          // upload / download pipeline (copies on copy_stream, packing on stream)
    cudaEvent_t frame_packed[2] = {}, frame_copied[2] = {};
    int frame_head = 0, frame_pending = 0;                // ring of two frames in flight
    orbc::Decomp mg;
    int nl_debug_mode = -1;                               // measurement aid (option debug_nl_mode): 1 / 2 = every evaluation records / searches
    int mg_own_slack = 8192;                              // slack of the owned-particle launch bounds (test option debug_own_slack)
    // per-class event-pair profiling (orbc_profile_*)
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev[ORBC_PROF_N];   // even = start, odd = stop
    size_t prof_used[ORBC_PROF_N] = {};
    bool kprof_on = false;
    std::vector<cudaEvent_t> kprof_ev;               // even = start, odd = stop
    std::vector<const char *> kprof_name;            // one per pair
    size_t kprof_used = 0;
};

namespace orbc {

// launch helper: counts launches (bench.py reports gpu_launches from this)
#define ORBC_LAUNCH(ctx, kernel, grid, block, smem, ...) do { \
        if ((ctx)->kprof_on) ::orbc::kprof_mark((ctx), #kernel, false); \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__); (ctx)->launches++; \
        if ((ctx)->kprof_on) ::orbc::kprof_mark((ctx), #kernel, true); \
        ORBC_CUDA(cudaGetLastError()); } while (0)

// per-launch event pairs keyed by kernel name (orbc_profile_kernels): an in-process launch list that also works on decomposed
// runs, where a replaying profiler cannot be used (the ranks wait for each other inside kernels)
inline void kprof_mark(orbc_ctx *c, const char *name, bool stop);

// event pair around the launches of one kernel class; no-ops unless profiling is on
inline int prof_mark(orbc_ctx *c, int cls, bool stop) {
    if (!c->prof_on) return ORBC_OK;
    auto &v = c->prof_ev[cls];
    size_t &u = c->prof_used[cls];
    if (u >= ((size_t)1 << 17)) return ORBC_OK;          // bounded: the oldest 65536 launches are kept
    if (u >= v.size()) { cudaEvent_t e; ORBC_CUDA(cudaEventCreate(&e)); v.push_back(e); }
    if (stop != (bool)(u & 1)) return ORBC_OK;           // unmatched stop / nested start: ignore
    ORBC_CUDA(cudaEventRecord(v[u], c->stream));
    ++u;
    return ORBC_OK;
}
struct ProfScope {
    orbc_ctx *c; int cls;
    ProfScope(orbc_ctx *c_, int cls_) : c(c_), cls(cls_) { prof_mark(c, cls, false); }
    ~ProfScope() { prof_mark(c, cls, true); }
};

inline void kprof_mark(orbc_ctx *c, const char *name, bool stop) {
    if (c->kprof_used >= ((size_t)1 << 18)) return;
    if (c->kprof_used >= c->kprof_ev.size()) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return; c->kprof_ev.push_back(e); }
    if (!stop) { if (c->kprof_name.size() <= c->kprof_used / 2) c->kprof_name.push_back(name); else c->kprof_name[c->kprof_used / 2] = name; }
    cudaEventRecord(c->kprof_ev[c->kprof_used++], c->stream);
}

inline unsigned blocks_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

template <class T> inline int dev_alloc(T **p, size_t n) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    ORBC_CUDA(cudaMalloc((void **)p, sizeof(T) * (n ? n : 1)));
    return ORBC_OK;
}
template <class T> inline void dev_free(T *&p) { if (p) cudaFree(p); p = nullptr; }

// strict fp32 (no FMA contraction) helpers — used wherever integer structures are derived from floating point so that
// they match the reference built without contraction (SURVEY.md §7 "bit-exact integer structures")
__device__ __forceinline__ float sq3_rn(float dx, float dy, float dz) {
    // ((0 + dx*dx) + dy*dy) + dz*dz — math_vector_base.h:209-213
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
__device__ __forceinline__ float dist2_rn(float4 a, float4 b) {
    return sq3_rn(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z));
}

} // namespace orbc
