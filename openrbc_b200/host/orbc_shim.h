// orbc_shim.h — the binding between OpenRBC's C++ host driver and liborbc_b200.so (include/orbc_b200.h).
//
// Include AFTER the reference's own headers (container.h, voronoi.h, runtime_parameter.h, integrate_nh.h ...).  It declares,
// in namespace openrbc::b200, functions with the NAMES and ARGUMENT MEANING of the hot-path calls of src/openrbc.cpp so that
// the driver's loops read the same; the reference's host code (RTParameter, init_rbc / init_random_sphere, VoronoiDiagram::init,
// save_topology, save_frame, display, Timers) is used unchanged.  Error behaviour follows the reference: a failed call prints
// and exit(0)s like rt_assert (util_misc.h:54-59).
//
//   reference call (src/openrbc.cpp)                                   here
//   voronoi.update(lipid, cell_lipid, param)                 :89,155,202   voronoi_update(dev, param)
//   cell_lipid.update(lipid, voronoi, param)                 :90,156,203   cell_update(dev, ORBC_LIPID, param)
//   cell_protein.update(protein, voronoi, param)             :91,157,204   cell_update(dev, ORBC_PROTEIN, param)
//   delete_lipid(lipid, voronoi, cell_lipid, param)          :201          delete_lipid(dev, lipid, param)
//   compute_pairwise_fused(voronoi, lipid, protein, cl, cp)  :100,167,219  compute_pairwise_fused(dev)
//   compute_bonded(protein)                                  :106,173,225  compute_bonded(dev)
//   integrate(KERNEL(param), lipid, protein)                 :94,110,...   integrate(dev, b200::KERNEL(param))  (same functor names)
//   compute_temperature(lipid, protein, param)               :143,253      compute_temperature(dev)
//   update_particle_affiliation + save_frame                 :138-140,248  save_frame(dev, traj, lipid, param): the frame is assembled on the
//                                                                          device in the .orbc layout and copied out while the run goes on
//                                                                          (or: download(dev, ...) then the reference's own save_frame)
#ifndef ORBC_SHIM_H_
#define ORBC_SHIM_H_

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <omp.h>

#include "orbc_b200.h"

namespace openrbc {
namespace b200 {

inline void check( int rc, const char * what ) {
    if ( rc != ORBC_OK ) {
        std::fprintf( stderr, "<Error> %s: %s\n", what, orbc_last_error() );
        std::exit( 0 );
    }
}

constexpr std::size_t vect_floats = sizeof( vect ) / sizeof( float );   // 3, or 4 under _ESIMD / _VEC4 (config_static.h:36-44)

// Owns the device contexts — one per GPU the cell is split over (one context = the single-GPU path) — uploads the containers
// the host initialised and downloads what save_frame / display read.  With several GPUs the decomposition follows the reference's
// own partition of the cells over its workers (util_numa.h:30-45): rank r owns a contiguous range of the Morton-ordered Voronoi
// cells; the contexts live in ONE process, are connected by peer pointers (no IPC), and every call below is issued for all
// ranks at once, one OpenMP host thread per rank, because the ranks wait for each other on the device.
struct Device {
    std::vector<orbc_ctx *> ranks;
    orbc_ctx * ctx = nullptr;  // rank 0 (the only one on a single GPU)
    bool time_calls = false;   // bracket calls with the reference's Timers (needs a stream sync per call)

    explicit Device( int device = 0 ) : Device( std::vector<int>( 1, device ) ) {}
    explicit Device( std::vector<int> const & devices ) {
        ranks.resize( devices.size(), nullptr );
        for ( std::size_t r = 0; r < devices.size(); ++r ) {
            check( orbc_create( &ranks[r], devices[r] ), "orbc_create" );
            if ( devices.size() > 1 ) check( orbc_mg_init( ranks[r], (int) r, (int) devices.size() ), "orbc_mg_init" );
        }
        ctx = ranks[0];
    }
    ~Device() { for ( auto c : ranks ) orbc_destroy( c ); }
    Device( Device const & ) = delete;
    int world() const { return (int) ranks.size(); }

    // f(rank, ctx) for every rank, concurrently
    template<class F> void each( F f ) {
        if ( ranks.size() == 1 ) { f( 0, ranks[0] ); return; }
        #pragma omp parallel num_threads( (int) ranks.size() )
        {
            const int r = omp_get_thread_num();
            f( r, ranks[r] );
        }
    }

    // after init_*() and voronoi.init() (openrbc.cpp:55-74): particles are stored sorted by cell
    void upload( LipidContainer const & lipid, ProteContainer const & protein, VoronoiDiagram const & voronoi,
                 VCellList const & cell_lipid, VCellList const & cell_protein ) {
        static_assert( sizeof( Bond ) == 3 * sizeof( int ), "Bond is (type, i, j)" );
        each( [&]( int, orbc_ctx * c ) {
            check( orbc_upload( c, ORBC_LIPID, lipid.size(), vect_floats, (const float *) lipid.x.data(), (const float *) lipid.v.data(),
                                (const float *) lipid.n.data(), (const float *) lipid.o.data(), nullptr, nullptr ), "orbc_upload(lipid)" );
            check( orbc_upload( c, ORBC_PROTEIN, protein.size(), vect_floats, (const float *) protein.x.data(), (const float *) protein.v.data(),
                                (const float *) protein.n.data(), (const float *) protein.o.data(), protein.type.data(), protein.tag.data() ), "orbc_upload(protein)" );
            check( orbc_upload_bonds( c, protein.bonds.size(), (const int *) protein.bonds.data() ), "orbc_upload_bonds" );
            check( orbc_voronoi_upload( c, voronoi.n_cells, (const float *) voronoi.centroids.data(), cell_lipid.cell_start.data(),
                                        protein.size() ? cell_protein.cell_start.data() : nullptr ), "orbc_voronoi_upload" );
        } );
        if ( ranks.size() > 1 ) {
            // connect the ranks: every context exports the pointers its peers write through, every context maps all of them
            const std::size_t each_bytes = orbc_mg_blob_bytes();
            std::vector<char> blobs( each_bytes * ranks.size() );
            each( [&]( int r, orbc_ctx * c ) { check( orbc_mg_export( c, blobs.data() + each_bytes * r, each_bytes ), "orbc_mg_export" ); } );
            each( [&]( int, orbc_ctx * c ) { check( orbc_mg_connect( c, blobs.data(), each_bytes ), "orbc_mg_connect" ); } );
        }
    }
    // type and tag of every protein slot in the device's CURRENT storage order (every rebuild reorders the containers): needed
    // before any host-side per-particle work that depends on the type, e.g. assign_temperature's sigma ~ 1 / sqrt(mass[type])
    void download_ids( ProteContainer & protein ) {
        if ( !protein.size() ) return;
        each( [&]( int, orbc_ctx * c ) {         // (a rank fills the rows it owns)
            check( orbc_download( c, ORBC_PROTEIN, vect_floats, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                  protein.type.data(), protein.tag.data(), nullptr ), "orbc_download(protein ids)" );
        } );
    }
    void upload_velocities( LipidContainer const & lipid, ProteContainer const & protein ) {
        each( [&]( int, orbc_ctx * c ) {
            if ( lipid.size() ) check( orbc_set_field( c, ORBC_LIPID, 'v', vect_floats, (const float *) lipid.v.data() ), "orbc_set_field" );
            if ( protein.size() ) check( orbc_set_field( c, ORBC_PROTEIN, 'v', vect_floats, (const float *) protein.v.data() ), "orbc_set_field" );
        } );
    }
    // what save_frame reads (trajectory.h:61-105): x, n, (v, f), affiliation, type, tag — in the device's current storage order.
    // Every rank writes the rows it owns into the same host containers.
    void download( LipidContainer & lipid, ProteContainer & protein, VCellList & cell_lipid, VCellList & cell_protein, int dump_field ) {
        std::size_t nl = 0, np = 0;
        check( orbc_size( ctx, ORBC_LIPID, &nl ), "orbc_size" );
        check( orbc_size( ctx, ORBC_PROTEIN, &np ), "orbc_size" );
        if ( nl != lipid.size() ) lipid.resize( nl );
        cell_lipid.affiliation.resize( nl );
        cell_protein.affiliation.resize( np );
        const bool vel = dump_field & 8, frc = dump_field & 16;
        each( [&]( int, orbc_ctx * c ) {
            check( orbc_download( c, ORBC_LIPID, vect_floats, (float *) lipid.x.data(), vel ? (float *) lipid.v.data() : nullptr, (float *) lipid.n.data(), nullptr,
                                  frc ? (float *) lipid.f.data() : nullptr, nullptr, cell_lipid.affiliation.data(), nullptr, nullptr, nullptr ), "orbc_download(lipid)" );
            check( orbc_download( c, ORBC_PROTEIN, vect_floats, (float *) protein.x.data(), vel ? (float *) protein.v.data() : nullptr, (float *) protein.n.data(), nullptr,
                                  frc ? (float *) protein.f.data() : nullptr, nullptr, cell_protein.affiliation.data(), protein.type.data(), protein.tag.data(), nullptr ), "orbc_download(protein)" );
        } );
    }
    void synchronize() { each( [&]( int, orbc_ctx * c ) { check( orbc_synchronize( c ), "orbc_synchronize" ); } ); }
};

// save_frame( traj, lipid, protein, cell_lipid, cell_protein, param ) (trajectory.h:61-105) with update_particle_affiliation folded in.
// One GPU: the frame is assembled on the device in the .orbc layout; the bytes of frame k are appended to the stream when frame
// k + 1 is requested (or by flush_frames), so the device-to-host copy overlaps the steps in between.  Several GPUs: every rank
// downloads its rows into the host's containers and the reference's own save_frame writes them.
inline void flush_frames( Device & dev, std::ostream & traj ) {
    const void * data = nullptr; std::size_t bytes = 0;
    while ( dev.world() == 1 && orbc_save_frame_end( dev.ctx, &data, &bytes ) == ORBC_OK ) traj.write( (const char *) data, (std::streamsize) bytes );
    traj << std::flush;
}
inline void save_frame( Device & dev, std::ostream & traj, LipidContainer & lipid, ProteContainer & protein, VCellList & cell_lipid, VCellList & cell_protein,
                        RTParameter const & param ) {
    if ( dev.world() == 1 ) {
        flush_frames( dev, traj );
        check( orbc_save_frame_begin( dev.ctx, param.nstep, param.dump_field, lipid.tag[0] ), "save_frame" );
    } else {
        dev.download( lipid, protein, cell_lipid, cell_protein, param.dump_field );
        openrbc::save_frame( traj, lipid, protein, cell_lipid, cell_protein, param );
    }
}

// brackets one device call with the reference's timer of the same name (timer.h:26-69)
struct TimedCall {
    Device & dev; std::string name;
    TimedCall( Device & d, const char * n ) : dev( d ), name( n ) { if ( dev.time_calls ) Service<Timers>::call()[name].start(); }
    ~TimedCall() { if ( dev.time_calls ) { dev.synchronize(); Service<Timers>::call()[name].stop(); } }
};

inline void voronoi_update( Device & dev, RTParameter const & param ) {
    TimedCall t( dev, "VoronoiDiagram::update" );
    dev.each( [&]( int, orbc_ctx * c ) { check( orbc_voronoi_update( c, param.nstep, param.freq_sort_ctrd ), "voronoi.update" ); } );
}
inline void cell_update( Device & dev, int species, RTParameter const & param ) {
    TimedCall t( dev, "VCellList::update" );
    dev.each( [&]( int, orbc_ctx * c ) { check( orbc_cell_update( c, species, param.nstep, param.freq_sort_bond ), "cell.update" ); } );
}
inline void delete_lipid( Device & dev, LipidContainer & lipid, RTParameter const & param ) {
    TimedCall t( dev, "cleanup_stray" );
    std::vector<std::size_t> n_new( dev.world(), 0 );
    dev.each( [&]( int r, orbc_ctx * c ) { check( orbc_delete_lipid( c, param.stray_tolerance, &n_new[r] ), "delete_lipid" ); } );
    Service<Variable<int, 0> >::call().value += int( lipid.size() - n_new[0] );   // the "Lost lipid" column (cleanup.h:87, display.h:46)
    if ( n_new[0] != lipid.size() ) lipid.resize( n_new[0] );
}
inline void compute_pairwise_fused( Device & dev ) {
    TimedCall t( dev, "compute_pairwise_fused" );
    dev.each( [&]( int, orbc_ctx * c ) { check( orbc_compute_pairwise_fused( c ), "compute_pairwise_fused" ); } );
}
inline void compute_bonded( Device & dev ) {
    TimedCall t( dev, "compute_bonded" );
    dev.each( [&]( int, orbc_ctx * c ) { check( orbc_compute_bonded( c ), "compute_bonded" ); } );
}
inline double compute_temperature( Device & dev ) {
    std::vector<double> T( dev.world(), 0.0 );
    dev.each( [&]( int r, orbc_ctx * c ) { check( orbc_compute_temperature( c, &T[r] ), "compute_temperature" ); } );
    double sum = 0;
    for ( double t : T ) sum += t;             // every rank returns its additive share of sum(m v^2) / 3N
    return sum;
}

// integrate(KERNEL, containers...) — integrate_nh.h:29-37.  The functor types are the reference's own; their names pick the
// device kernel, their RTParameter reference supplies dt, kBT, eta, zeta, the box, and receives the Nose-Hoover friction
// update that the reference performs in the functor's destructor (integrate_nh.h:181-185, 240-244).
inline orbc_step_params step_params( RTParameter const & param ) {
    orbc_step_params p;
    p.dt = param.dt; p.kBT = param.kBT; p.eta = param.eta; p.zeta = param.zeta;
    p.box_lo = param.box[0][0]; p.box_hi = param.box[0][1];
    p.dr_opt = param.dr_opt; p.dn_opt = param.dn_opt;
    p.nstep = param.nstep; p.seed = (uint64_t) param.rseed;
    p.noise_lipid = p.noise_protein = nullptr;
    return p;
}
inline void integrate_id( Device & dev, int kernel, const char * name, RTParameter * param ) {
    TimedCall t( dev, name );
    orbc_step_params p;
    if ( param ) p = step_params( *param );
    const bool reduces = kernel == ORBC_NH_INITIAL_FUSED || kernel == ORBC_NH_FINAL_FUSED || kernel == ORBC_NH_UPDATE;
    std::vector<orbc_step_result> res( dev.world() );
    dev.each( [&]( int r, orbc_ctx * c ) { check( orbc_integrate( c, kernel, param ? &p : nullptr, reduces ? &res[r] : nullptr ), name ); } );
    if ( reduces ) {                           // (on several GPUs every rank returns the kinetic energy of the whole system)
        float Q = param->Q;
        param->zeta = kernel == ORBC_NH_UPDATE ? orbc_nh_zeta_update_unfused( param->zeta, &Q, param->dt, param->kBT, res[0].ke, res[0].n )
                                               : orbc_nh_zeta_update( param->zeta, &Q, param->dt, param->kBT, res[0].ke, res[0].n );
        param->Q = Q;
    }
}
// Kernel tags with the reference's functor names (integrate_nh.h:58-273, integrate_langevin.h:99-149).  They are separate
// types, not the reference's structs, because those do their work in operator() on host arrays and — the Nose-Hoover ones —
// update zeta in their destructor from members that only operator() fills.
struct clear_force {};
struct post_torque {};
struct bounce_back { RTParameter & parameter; explicit bounce_back( RTParameter & p ) : parameter( p ) {} };
struct verlet_langevin { RTParameter & parameter; explicit verlet_langevin( RTParameter & p ) : parameter( p ) {} };
struct verlet_initial_bounce_clearforce_update { RTParameter & parameter; explicit verlet_initial_bounce_clearforce_update( RTParameter & p ) : parameter( p ) {} };
struct post_toque_final_update { RTParameter & parameter; explicit post_toque_final_update( RTParameter & p ) : parameter( p ) {} };
struct verlet_nh_final { RTParameter & parameter; explicit verlet_nh_final( RTParameter & p ) : parameter( p ) {} };
struct verlet_nh_update { RTParameter & parameter; explicit verlet_nh_update( RTParameter & p ) : parameter( p ) {} };

inline void integrate( Device & dev, clear_force const & ) { integrate_id( dev, ORBC_CLEAR_FORCE, "clear_force", nullptr ); }
inline void integrate( Device & dev, post_torque const & ) { integrate_id( dev, ORBC_POST_TORQUE, "post_torque", nullptr ); }
inline void integrate( Device & dev, bounce_back const & k ) { integrate_id( dev, ORBC_BOUNCE_BACK, "bounce_back", &k.parameter ); }
inline void integrate( Device & dev, verlet_langevin const & k ) { integrate_id( dev, ORBC_VERLET_LANGEVIN, "verlet_langevin", &k.parameter ); }
inline void integrate( Device & dev, verlet_initial_bounce_clearforce_update const & k ) { integrate_id( dev, ORBC_NH_INITIAL_FUSED, "verlet_initial_bounce_clearforce", &k.parameter ); }
inline void integrate( Device & dev, post_toque_final_update const & k ) { integrate_id( dev, ORBC_NH_FINAL_FUSED, "post_toque_final_update", &k.parameter ); }
inline void integrate( Device & dev, verlet_nh_final const & k ) { integrate_id( dev, ORBC_NH_FINAL, "verlet_nh_final", &k.parameter ); }
inline void integrate( Device & dev, verlet_nh_update const & k ) { integrate_id( dev, ORBC_NH_UPDATE, "verlet_nh_update", &k.parameter ); }
// the minimiser's capped steepest-descent move (openrbc.cpp:114-131), a plain loop in the reference
inline void opt_move( Device & dev, RTParameter & param ) { integrate_id( dev, ORBC_OPT_MOVE, "OptIntegration", &param ); }
// post_torque + mover + bounce_back of one minimisation step (openrbc.cpp:110-133) as one pass; the form a decomposed run uses
// (the new positions are pushed to the neighbouring ranks from the same kernel)
inline void opt_fused( Device & dev, RTParameter & param ) { integrate_id( dev, ORBC_OPT_FUSED, "OptIntegration", &param ); }

}  // namespace b200
}  // namespace openrbc

#endif
