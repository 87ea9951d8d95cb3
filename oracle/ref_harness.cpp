// TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT PATH.
//
// Thin C-callable harness around the UNMODIFIED OpenRBC reference headers.  It is compiled
// from the sources where they lie (-I/root/reference/src) by oracle/Makefile into
// oracle/_ref/libref_{strict,fast}.so; no reference source is copied into this repository.
// The harness owns one reference "world" (RTParameter, LipidContainer, ProteContainer,
// VoronoiDiagram, two VCellLists) and exposes each hot-path function of SURVEY.md §8(a) as a
// single call so that tests can teacher-force identical inputs into the reference, the C port
// (oracle/orbc_oracle.c) and the CUDA library, and compare the outputs of that one call.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.
//
// Reference call sites mirrored here: src/openrbc.cpp:43-264.

#include <limits>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <memory>

#include "config_static.h"
#include "compute_bonded.h"
#include "compute_pairwise.h"
#include "compute_pairwise_simd.h"
#include "compute_pairwise_fused.h"
#include "compute_pairwise_lp.h"
#include "compute_temperature.h"
#include "cleanup.h"
#include "runtime_parameter.h"
#include "container.h"
#include "display.h"
#include "init_random.h"
#include "init_rbc.h"
#include "integrate_nh.h"
#include "integrate_langevin.h"
#include "forcefield.h"
#include "remove_bonds.h"
#include "rng.h"
#include "timer.h"
#include "topology.h"
#include "trajectory.h"
#include "voronoi.h"
#include "constrain_volume.h"
#include "zero_bulk_velocity.h"
#include "assign_temperature.h"

using namespace openrbc;
using namespace openrbc::config;

namespace {

struct World {
    RTParameter param;
    LipidContainer lipid;
    ProteContainer protein;
    std::unique_ptr<VoronoiDiagram> voronoi;
    VCellList cell_lipid, cell_protein;
    World( int argc, char ** argv ) : param( argc, argv ), lipid( "lipid" ), protein( "protein" ) {}
};

World * W = nullptr;

Container & species( int s ) { return s == 0 ? static_cast<Container &>( W->lipid ) : static_cast<Container &>( W->protein ); }
VCellList & celllist( int s ) { return s == 0 ? W->cell_lipid : W->cell_protein; }

AlignedArray<vect, true> * field( int s, const char * name ) {
    Container & c = species( s );
    switch ( name[0] ) {
    case 'x': return &c.x;
    case 'v': return &c.v;
    case 'f': return &c.f;
    case 'n': return &c.n;
    case 'o': return &c.o;
    case 't': return &c.t;
    }
    return nullptr;
}

// steepest-descent mover of the energy-minimisation loop, openrbc.cpp:114-131 (restated: the
// reference has it inline in main()).
template<class C> static void opt_move( C & c, RTParameter & param ) {
    #pragma omp parallel for
    for ( std::size_t i = 0; i < c.size(); ++i ) {
        auto dx = c.f[i] / ForceField::mass[c.type[i]];
        auto dn = cross( c.t[i], c.n[i] );
        auto dt = ( norm( dx ) > param.dr_opt || norm( dn ) > param.dn_opt ) ? std::min( param.dr_opt / norm( dx ), param.dn_opt / norm( dn ) ) : param.dt;
        c.x[i] += c.f[i] * ( dt / ForceField::mass[c.type[i]] );
        c.n[i] += cross( c.t[i], c.n[i] ) * dt;
        c.n[i]  = normalize( c.n[i] );
    }
}

}

extern "C" {

// ---- lifetime ------------------------------------------------------------------------------
// argv-style construction so that every CLI default of runtime_parameter.h:41-71 is the
// reference's own.  n_threads <= 0 keeps the OpenMP default.
int ref_create( int argc, char ** argv, int n_threads ) {
    if ( n_threads > 0 ) omp_set_num_threads( n_threads );
    omp_set_nested( 1 ); // openrbc.cpp:47
    delete W;
    W = new World( argc, argv );
    return 0;
}

void ref_destroy() { delete W; W = nullptr; }

int ref_num_threads() { return omp_get_max_threads(); }
int ref_vect_dim() { return vect::d; }

// ---- initialisation (host driver code that stays in place; openrbc.cpp:55-73) -----------------
int ref_init_lipid_sphere( float r ) { init_random_sphere( W->lipid, W->param, r ); return 0; }
int ref_init_trimesh() { init_rbc( W->lipid, W->protein, W->param, 0 ); return 0; }

// openrbc.cpp:69-74
int ref_voronoi_init( int n_iterate ) {
    W->voronoi.reset( new VoronoiDiagram( std::max<std::size_t>( 1, W->lipid.size() / W->param.voronoi_cell_size ) ) );
    W->voronoi->init( W->lipid, W->cell_lipid, W->param, n_iterate );
    W->cell_lipid.update_particle_affiliation( W->lipid );
    W->cell_protein.update( W->protein, *W->voronoi, W->param );
    W->cell_protein.update_particle_affiliation( W->protein );
    return W->voronoi->n_cells;
}

// ---- parameters -----------------------------------------------------------------------------
int ref_set_param( const char * name, double v ) {
    RTParameter & p = W->param;
    std::string n( name );
    if ( n == "nstep" ) p.nstep = (int)v;
    else if ( n == "dt" ) p.dt = v;
    else if ( n == "kBT" ) p.kBT = v;
    else if ( n == "eta" ) p.eta = v;
    else if ( n == "zeta" ) p.zeta = v;
    else if ( n == "Q" ) p.Q = v;
    else if ( n == "stray_tolerance" ) p.stray_tolerance = v;
    else if ( n == "freq_voronoi" ) p.freq_voronoi = (int)v;
    else if ( n == "freq_sort_ctrd" ) p.freq_sort_ctrd = (int)v;
    else if ( n == "freq_sort_bond" ) p.freq_sort_bond = (int)v;
    else if ( n == "freq_cleanup" ) p.freq_cleanup = (int)v;
    else if ( n == "dr_opt" ) p.dr_opt = v;
    else if ( n == "dn_opt" ) p.dn_opt = v;
    else if ( n == "rho" ) p.rho = v;
    else if ( n == "voronoi_cell_size" ) p.voronoi_cell_size = (int)v;
    // the reflecting walls (runtime_parameter.h:45,111-115); bsize (the Morton quantisation) is left alone
    else if ( n == "box_lo" ) { for ( int d = 0; d < 3; ++d ) p.box[d][0] = v; }
    else if ( n == "box_hi" ) { for ( int d = 0; d < 3; ++d ) p.box[d][1] = v; }
    else return -1;
    return 0;
}

double ref_get_param( const char * name ) {
    RTParameter & p = W->param;
    std::string n( name );
    if ( n == "nstep" ) return p.nstep;
    if ( n == "dt" ) return p.dt;
    if ( n == "kBT" ) return p.kBT;
    if ( n == "eta" ) return p.eta;
    if ( n == "zeta" ) return p.zeta;
    if ( n == "Q" ) return p.Q;
    if ( n == "stray_tolerance" ) return p.stray_tolerance;
    if ( n == "lost_lipid" ) return Service<Variable<int, 0> >::call().value;
    return std::numeric_limits<double>::quiet_NaN();
}

// ---- sizes / getters / setters -------------------------------------------------------------------
long ref_size( int s ) { return (long)species( s ).size(); }
int ref_n_cells() { return W->voronoi ? W->voronoi->n_cells : 0; }
long ref_n_bonds() { return (long)W->protein.bonds.size(); }
int ref_lipid_tag_base() { return W->lipid.tag[0]; }

void ref_get( int s, const char * name, float * dst ) {
    auto & a = *field( s, name );
    const long n = species( s ).size();
    for ( long i = 0; i < n; ++i ) for ( int d = 0; d < 3; ++d ) dst[3 * i + d] = a[i][d];
}

void ref_set( int s, const char * name, const float * src ) {
    auto & a = *field( s, name );
    const long n = species( s ).size();
    for ( long i = 0; i < n; ++i ) {
        for ( int d = 0; d < 3; ++d ) a[i][d] = src[3 * i + d];
        for ( uint d = 3; d < vect::d; ++d ) a[i][d] = 0;
    }
}

void ref_get_protein_ids( int * type, int * tag ) {
    for ( std::size_t i = 0; i < W->protein.size(); ++i ) { type[i] = W->protein.type[i]; tag[i] = W->protein.tag[i]; }
}

void ref_get_bonds( int * type_i_j ) {
    for ( std::size_t b = 0; b < W->protein.bonds.size(); ++b ) {
        type_i_j[3 * b + 0] = W->protein.bonds[b].type;
        type_i_j[3 * b + 1] = W->protein.bonds[b].i;
        type_i_j[3 * b + 2] = W->protein.bonds[b].j;
    }
}

void ref_get_centroids( float * dst ) {
    for ( int i = 0; i < W->voronoi->n_cells; ++i ) for ( int d = 0; d < 3; ++d ) dst[3 * i + d] = W->voronoi->centroids[i][d];
}

// what: 0 cell_start (n_cells+1), 1 cells (n), 2 affiliation (n), 3 local_index (n)
void ref_get_cell_array( int s, int what, int * dst ) {
    VCellList & c = celllist( s );
    AlignedArray<int> * a = what == 0 ? &c.cell_start : what == 1 ? &c.cells : what == 2 ? &c.affiliation : &c.local_index;
    for ( std::size_t i = 0; i < a->size(); ++i ) dst[i] = ( *a )[i];
}

// Replace the whole world state (teacher forcing).  Arrays are N x 3 floats, packed.
// Lipids: type/tag are implicit (container.h:117-131).  cell_start arrays may be NULL.
int ref_set_lipids( long n, const float * x, const float * v, const float * nn, const float * o ) {
    W->lipid.resize( n );
    ref_set( 0, "x", x ); ref_set( 0, "v", v ); ref_set( 0, "n", nn ); ref_set( 0, "o", o );
    W->lipid.f.assign( n, real( 0 ) ); W->lipid.t.assign( n, real( 0 ) );
    W->lipid.tag2idx.build_map( W->lipid );
    return 0;
}

int ref_set_proteins( long n, const float * x, const float * v, const float * nn, const float * o,
                      const int * type, const int * tag, long n_bonds, const int * type_i_j ) {
    W->protein.resize( n );
    ref_set( 1, "x", x ); ref_set( 1, "v", v ); ref_set( 1, "n", nn ); ref_set( 1, "o", o );
    W->protein.f.assign( n, real( 0 ) ); W->protein.t.assign( n, real( 0 ) );
    for ( long i = 0; i < n; ++i ) { W->protein.type[i] = type[i]; W->protein.tag[i] = tag[i]; }
    W->protein.bonds.resize( 0 );
    for ( long b = 0; b < n_bonds; ++b ) W->protein.bonds.emplace_back( type_i_j[3 * b], type_i_j[3 * b + 1], type_i_j[3 * b + 2] );
    Service<BalancerMap>::call()[ W->protein.id() + "-bonds" ].set_range( W->protein.bonds.size() );
    W->protein.tag2idx.build_map( W->protein );
    W->lipid.tag.set_base( n + 1 ); // init_rbc.h:296
    return 0;
}

// Install a Voronoi diagram: centroids + the (sorted-by-cell) partition of both containers.
int ref_set_voronoi( int n_cells, const float * centroids, const int * cell_start_l, const int * cell_start_p ) {
    if ( !W->voronoi || W->voronoi->n_cells != n_cells ) W->voronoi.reset( new VoronoiDiagram( n_cells ) );
    W->voronoi->centroids.resize( n_cells );
    for ( int i = 0; i < n_cells; ++i ) for ( int d = 0; d < 3; ++d ) W->voronoi->centroids[i][d] = centroids[3 * i + d];
    W->voronoi->tree.build( W->voronoi->centroids );
    for ( int s = 0; s < 2; ++s ) {
        const int * cs = s == 0 ? cell_start_l : cell_start_p;
        VCellList & c = celllist( s );
        const std::size_t n = species( s ).size();
        c.n_cells = n_cells;
        c.cell_start.resize( n_cells + 1 );
        c.cells.resize( n ); c.affiliation.resize( n ); c.local_index.resize( n );
        for ( int i = 0; i <= n_cells; ++i ) c.cell_start[i] = cs ? cs[i] : 0;
        for ( std::size_t i = 0; i < n; ++i ) { c.cells[i] = i; c.affiliation[i] = -1; c.local_index[i] = 0; }
        if ( cs ) c.update_particle_affiliation( species( s ) );
    }
    return 0;
}

// ---- hot-path calls, one reference function each -------------------------------------------------
void ref_voronoi_update() { W->voronoi->update( W->lipid, W->cell_lipid, W->param ); }                 // openrbc.cpp:202
void ref_cell_update( int s ) {                                                                        // openrbc.cpp:203-204
    if ( s == 0 ) W->cell_lipid.update( W->lipid, *W->voronoi, W->param );
    else W->cell_protein.update( W->protein, *W->voronoi, W->param );
}
void ref_update_particle_affiliation( int s ) { celllist( s ).update_particle_affiliation( species( s ) ); }
void ref_compute_pairwise_fused() { compute_pairwise_fused( *W->voronoi, W->lipid, W->protein, W->cell_lipid, W->cell_protein ); } // :219
void ref_compute_bonded() { compute_bonded( W->protein ); }                                            // :225
double ref_compute_temperature() { return compute_temperature( W->lipid, W->protein, W->param ); }     // :253
void ref_constrain_volume( float target, float strength ) {                                            // :229
    constrain_volume( W->lipid, W->protein, *W->voronoi, W->cell_lipid, W->cell_protein, W->param, target, strength );
}
long ref_delete_lipid() { delete_lipid( W->lipid, *W->voronoi, W->cell_lipid, W->param ); return W->lipid.size(); } // :201

// kernel ids follow include/orbc_b200.h (orbc_integrator)
int ref_integrate( int kernel ) {
    switch ( kernel ) {
    case 0: integrate( clear_force(), W->lipid, W->protein ); break;
    case 1: integrate( post_torque(), W->lipid, W->protein ); break;
    case 2: integrate( bounce_back( W->param ), W->lipid, W->protein ); break;
    case 3: integrate( verlet_langevin( W->param ), W->lipid, W->protein ); break;
    case 4: integrate( verlet_initial_bounce_clearforce_update( W->param ), W->lipid, W->protein ); break;
    case 5: integrate( post_toque_final_update( W->param ), W->lipid, W->protein ); break;
    case 6: integrate( verlet_nh_final( W->param ), W->lipid, W->protein ); break;
    case 7: integrate( verlet_nh_update( W->param ), W->lipid, W->protein ); break;
    case 8: integrate( assign_temperature( W->param ), W->lipid, W->protein ); break;
    default: return -1;
    }
    return 0;
}

void ref_opt_move() { opt_move( W->lipid, W->param ); opt_move( W->protein, W->param ); }

int ref_get_stencil( int cell, float rmax, int * out, int cap ) {
    static AlignedArray<int, true> stencil;
    int n = W->voronoi->get_stencil_whole( cell, stencil, rmax );
    for ( int i = 0; i < n && i < cap; ++i ) out[i] = stencil[i];
    return n;
}

// stencil at r<9 then refined to 8 and 6 exactly as compute_pairwise_fused.h:260,278,299
int ref_get_stencil_refined( int cell, int * out9, int * out8, int * out6, int cap, int * n986 ) {
    static AlignedArray<int, true> stencil;
    int n = W->voronoi->get_stencil_whole( cell, stencil, 9.0f );
    n986[0] = n; for ( int i = 0; i < n && i < cap; ++i ) out9[i] = stencil[i];
    n = W->voronoi->refine_stencil( cell, n, stencil, 8.0f );
    n986[1] = n; for ( int i = 0; i < n && i < cap; ++i ) out8[i] = stencil[i];
    n = W->voronoi->refine_stencil( cell, n, stencil, 6.0f );
    n986[2] = n; for ( int i = 0; i < n && i < cap; ++i ) out6[i] = stencil[i];
    return 0;
}

unsigned ref_morton_encode( float x, float y, float z ) { return morton_encode( x, y, z, W->param ); }

// Morton-sort an arbitrary point set with the reference's own reorder_morton (reorder_morton.h:44)
void ref_reorder_morton( long n, float * pts ) {
    AlignedArray<vector<real, 3>, true> a;
    a.resize( n );
    for ( long i = 0; i < n; ++i ) for ( int d = 0; d < 3; ++d ) a[i][d] = pts[3 * i + d];
    reorder_morton( a, W->param );
    for ( long i = 0; i < n; ++i ) for ( int d = 0; d < 3; ++d ) pts[3 * i + d] = a[i][d];
}

float ref_uint2u11( unsigned u ) { vector<uint, 3> w( u, u, u ); return uint2u11( w )[0]; }

// ---- the reference's main MD loop, default build (LANGEVIN + FUSED_PAIRWISE), openrbc.cpp:189-256,
// restated call-for-call (I/O and display omitted).  Used for CPU-baseline timing only.
// Returns wall seconds for n_steps; param.nstep advances.
double ref_run_langevin( int n_steps, int with_cleanup ) {
    RTParameter & param = W->param;
    double t0 = omp_get_wtime();
    for ( int s = 0; s < n_steps; ++s ) {
        if ( param.nstep % param.freq_voronoi == 0 ) {
            if ( with_cleanup && param.nstep % param.freq_cleanup == 0 ) delete_lipid( W->lipid, *W->voronoi, W->cell_lipid, param );
            W->voronoi->update( W->lipid, W->cell_lipid, param );
            W->cell_lipid.update( W->lipid, *W->voronoi, param );
            W->cell_protein.update( W->protein, *W->voronoi, param );
        }
        compute_pairwise_fused( *W->voronoi, W->lipid, W->protein, W->cell_lipid, W->cell_protein );
        compute_bonded( W->protein );
        integrate( verlet_langevin( param ), W->lipid, W->protein );
        ++param.nstep;
    }
    return omp_get_wtime() - t0;
}

// Nose-Hoover build of the same loop (LANGEVIN undefined, FUSED_INTEGRATOR), openrbc.cpp:192-241
double ref_run_nh( int n_steps, int with_cleanup ) {
    RTParameter & param = W->param;
    double t0 = omp_get_wtime();
    for ( int s = 0; s < n_steps; ++s ) {
        integrate( verlet_initial_bounce_clearforce_update( param ), W->lipid, W->protein );
        if ( param.nstep % param.freq_voronoi == 0 ) {
            if ( with_cleanup && param.nstep % param.freq_cleanup == 0 ) delete_lipid( W->lipid, *W->voronoi, W->cell_lipid, param );
            W->voronoi->update( W->lipid, W->cell_lipid, param );
            W->cell_lipid.update( W->lipid, *W->voronoi, param );
            W->cell_protein.update( W->protein, *W->voronoi, param );
        }
        compute_pairwise_fused( *W->voronoi, W->lipid, W->protein, W->cell_lipid, W->cell_protein );
        compute_bonded( W->protein );
        integrate( post_toque_final_update( param ), W->lipid, W->protein );
        ++param.nstep;
    }
    return omp_get_wtime() - t0;
}

// energy-minimisation loop, openrbc.cpp:88-135
double ref_run_opt( int n_steps ) {
    RTParameter & param = W->param;
    double t0 = omp_get_wtime();
    for ( int s = 0; s < n_steps; ++s ) {
        W->voronoi->update( W->lipid, W->cell_lipid, param );
        W->cell_lipid.update( W->lipid, *W->voronoi, param );
        W->cell_protein.update( W->protein, *W->voronoi, param );
        integrate( clear_force(), W->lipid, W->protein );
        compute_pairwise_fused( *W->voronoi, W->lipid, W->protein, W->cell_lipid, W->cell_protein );
        compute_bonded( W->protein );
        integrate( post_torque(), W->lipid, W->protein );
        ref_opt_move();
        integrate( bounce_back( param ), W->lipid, W->protein );
    }
    return omp_get_wtime() - t0;
}

// openrbc.cpp:247-250: update_particle_affiliation of both cell lists, then save_frame (trajectory.h:61-105) into `path`
int ref_save_frame( const char * path, int dump_field ) {
    W->cell_lipid.update_particle_affiliation( W->lipid );
    W->cell_protein.update_particle_affiliation( W->protein );
    const int keep = W->param.dump_field;
    W->param.dump_field = dump_field;
    std::ofstream f( path, std::ios::binary );
    save_frame( f, W->lipid, W->protein, W->cell_lipid, W->cell_protein, W->param );
    W->param.dump_field = keep;
    return f.good() ? 0 : -1;
}

double ref_timer( const char * name ) { return Service<Timers>::call()[ name ].read(); }
void ref_timers_report() { Service<Timers>::call().report( false ); }

// force-field tables, for byte comparison with the port / device constants (forcefield_canonical.h:30-156)
// layout documented in include/orbc_b200.h (orbc_forcefield)
void ref_get_forcefield( float * out ) {
    int k = 0;
    for ( int i = 0; i < 6; ++i ) out[k++] = ForceField::mass[i];
    for ( int i = 0; i < 6; ++i ) out[k++] = ForceField::radius[i];
    for ( int i = 0; i < 6; ++i ) out[k++] = ForceField::cutlp[i];
    for ( int i = 0; i < 6; ++i ) out[k++] = ForceField::cutsqlp[i];
    for ( int i = 0; i < 6; ++i ) out[k++] = ForceField::replp[i];
    for ( int i = 0; i < 6; ++i ) out[k++] = ForceField::attlp[i];
    for ( int i = 0; i < 6; ++i ) out[k++] = ForceField::alphalp[i];
    for ( int i = 0; i < 36; ++i ) out[k++] = ForceField::cutpp[i];
    for ( int i = 0; i < 36; ++i ) out[k++] = ForceField::cutsqpp[i];
    for ( int i = 0; i < 36; ++i ) out[k++] = ForceField::reppp[i];
    for ( int i = 0; i < 36; ++i ) out[k++] = ForceField::lj_cutsq[i];
    for ( int i = 0; i < 36; ++i ) out[k++] = ForceField::lj_lj1[i];
    for ( int i = 0; i < 36; ++i ) out[k++] = ForceField::lj_lj2[i];
    for ( int i = 0; i < 4; ++i ) out[k++] = ForceField::r0[i];
    for ( int i = 0; i < 4; ++i ) out[k++] = ForceField::K[i];
    out[k++] = ForceField::cutll; out[k++] = ForceField::cutsqll; out[k++] = ForceField::repll;
    out[k++] = ForceField::attll; out[k++] = ForceField::alphall;
}
int ref_forcefield_floats() { return 6 * 7 + 36 * 6 + 8 + 5; }

}
