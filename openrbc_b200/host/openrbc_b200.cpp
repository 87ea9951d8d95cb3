// openrbc_b200.cpp — OpenRBC's host driver with the per-timestep loop on a B200.
//
// Everything the north star leaves on the host is the REFERENCE's own code, included unchanged from its source tree
// (-I$(REF)/src): RTParameter and its command line (runtime_parameter.h), init_random_sphere / init_rbc, save_topology,
// VoronoiDiagram::init (the 64 Lloyd iterations that seed the Voronoi cells), save_frame, display, Timers.  The calls of the
// minimisation loop (src/openrbc.cpp:88-146) and of the main loop (:189-256) go through orbc_shim.h to liborbc_b200.so.
// The loop bodies below are written against the call sequence of the reference's driver, in the default build configuration
// (LANGEVIN + FUSED_PAIRWISE, config_static.h:30-32); ORBC_INTEGRATOR=nh selects the Nose-Hoover pair (FUSED_INTEGRATOR) at run time.
//
// Same command line as the reference (`./openrbc_b200 -i trimesh -m <mesh> -E 100 -t 10 ...`), same outputs (cell.data, cell.orbc,
// the 4-column progress table, the final "T s on K steps * N particles" line).  Extra environment: ORBC_DEVICE (default 0) or
// ORBC_DEVICES=0,1,... (one cell split over several GPUs of the box),
// ORBC_TIMERS=1 (per-call timers with a stream sync, like the reference's Timers report).
#include <limits>
#include <fstream>
#include <iostream>
#include <sstream>

#include "config_static.h"
#include "container.h"
#include "runtime_parameter.h"
#include "forcefield.h"
#include "voronoi.h"
#include "integrate_nh.h"
#include "integrate_langevin.h"
#include "assign_temperature.h"
#include "init_random.h"
#include "init_rbc.h"
#include "remove_bonds.h"
#include "topology.h"
#include "trajectory.h"
#include "display.h"
#include "timer.h"
#include "citation.h"

#include "orbc_shim.h"

int main( int argc, char ** argv ) {
    using namespace openrbc;
    using namespace openrbc::config;

    RTParameter param( argc, argv );
    const char * env_dev = std::getenv( "ORBC_DEVICE" );
    const char * env_int = std::getenv( "ORBC_INTEGRATOR" );
    const bool nose_hoover = env_int && std::string( env_int ) == "nh";
    const bool fused_opt = std::getenv( "ORBC_OPT_FUSED" ) != nullptr;   // one pass per minimisation step also on a single GPU

    // ---- host initialisation: the reference's code, unchanged (openrbc.cpp:51-79) --------------------------------------------
    std::cout << "Initializing system ..." << std::flush;
    LipidContainer lipid( "lipid" );
    ProteContainer protein( "protein" );
    if ( param.init == "lipid" ) {
        init_random_sphere( lipid, param, 100 );
    } else if ( param.init == "vesicle" ) {
        init_rbc( lipid, protein, param, 500 );
        remove_bonds( protein, UnaryPredicate() );
    } else if ( param.init == "trimesh" ) {
        init_rbc( lipid, protein, param, 0 );
    }
    save_topology( std::ofstream( param.file_topo ), param, protein, lipid );
    std::cout << "Done." << std::endl;

    std::cout << "Initializing Voronoi cells " << std::flush;
    VoronoiDiagram voronoi( std::max<std::size_t>( 1, lipid.size() / param.voronoi_cell_size ) );
    VCellList cell_lipid, cell_protein;
    voronoi.init( lipid, cell_lipid, param, 64 );
    cell_lipid.update_particle_affiliation( lipid );
    cell_protein.update( protein, voronoi, param );
    cell_protein.update_particle_affiliation( protein );
    std::cout << "Done." << std::endl;
    std::ofstream traj( param.file_traj );
    save_frame( traj, lipid, protein, cell_lipid, cell_protein, param );
    std::cout << "Initialization complete." << std::endl;
    Service<Timers>::call().report( true );

    // ---- hand the containers to the device --------------------------------------------------------------------------------
    // ORBC_DEVICES=0,1,2,3: the cell is split over these GPUs (contiguous ranges of the Morton-ordered Voronoi cells per GPU, the
    // reference's own partition over its workers, util_numa.h:30-45); ORBC_DEVICE=k or nothing: one GPU
    std::vector<int> devices;
    if ( const char * env_devs = std::getenv( "ORBC_DEVICES" ) ) {
        std::stringstream ss( env_devs );
        for ( std::string tok; std::getline( ss, tok, ',' ); ) if ( tok.size() ) devices.push_back( std::atoi( tok.c_str() ) );
    }
    if ( devices.empty() ) devices.push_back( env_dev ? std::atoi( env_dev ) : 0 );
    b200::Device dev( devices );
    dev.time_calls = std::getenv( "ORBC_TIMERS" ) != nullptr;
    dev.upload( lipid, protein, voronoi, cell_lipid, cell_protein );

    auto rebuild = [&]() {
        b200::voronoi_update( dev, param );
        b200::cell_update( dev, ORBC_LIPID, param );
        b200::cell_update( dev, ORBC_PROTEIN, param );
    };
    auto forces = [&]() {
        b200::compute_pairwise_fused( dev );
        b200::compute_bonded( dev );
    };
    auto dump = [&]() { b200::save_frame( dev, traj, lipid, protein, cell_lipid, cell_protein, param ); };

    // ---- energy minimisation (openrbc.cpp:88-146) ---------------------------------------------------------------------------
    std::cout << "Opt..." << std::endl;
    Service<Timers>::call()["+optimization"].start();
    for ( int nopt = 0; nopt < param.opt_nstep; ++nopt ) {
        rebuild();
        b200::integrate( dev, b200::clear_force() );
        forces();
        if ( dev.world() == 1 && !fused_opt ) {
            b200::integrate( dev, b200::post_torque() );
            b200::opt_move( dev, param );
            b200::integrate( dev, b200::bounce_back( param ) );
        } else b200::opt_fused( dev, param );                     // the same three operations in one pass, with the halo push
        if ( ( nopt + 1 ) % param.freq_dump == 0 ) dump();
        if ( ( nopt + 1 ) % param.freq_display == 0 )
            display( std::cout, nopt + 1, b200::compute_temperature( dev ), omp_get_wtime() - Service<Timers>::call()["+optimization"].get_start_time() );
    }
    dev.synchronize();
    b200::flush_frames( dev, traj );
    Service<Timers>::call()["+optimization"].stop();
    Service<Timers>::call().report( true );

    // ---- main loop (openrbc.cpp:151-256) --------------------------------------------------------------------------------
    std::cout << "Run ... " << std::endl;
    Service<Timers>::call()["+main-loop"].start();
    // Maxwell velocities from the reference's generator (assign_temperature.h:26-44), drawn on the host.  The minimisation has
    // reordered the device's containers at every rebuild, and the velocity spread depends on mass[type]: bring the host's type /
    // tag arrays into the device's current slot order first, so that slot i receives a velocity drawn for ITS type
    dev.download_ids( protein );
    integrate( assign_temperature( param ), lipid, protein );
    dev.upload_velocities( lipid, protein );
    rebuild();
    b200::integrate( dev, b200::clear_force() );
    forces();
    if ( nose_hoover ) b200::integrate( dev, b200::post_toque_final_update( param ) );
    else b200::integrate( dev, b200::verlet_langevin( param ) );

    param.nstep = 0;
    while ( param.nstep * param.dt < param.t_total ) {
        if ( nose_hoover ) b200::integrate( dev, b200::verlet_initial_bounce_clearforce_update( param ) );
        if ( param.nstep % param.freq_voronoi == 0 ) {
            if ( param.nstep % param.freq_cleanup == 0 ) b200::delete_lipid( dev, lipid, param );
            rebuild();
        }
        forces();
        if ( nose_hoover ) b200::integrate( dev, b200::post_toque_final_update( param ) );
        else b200::integrate( dev, b200::verlet_langevin( param ) );
        ++param.nstep;
        if ( param.nstep % param.freq_dump == 0 ) dump();
        if ( param.nstep % param.freq_display == 0 )
            display( std::cout, param.nstep * param.dt, b200::compute_temperature( dev ), omp_get_wtime() - Service<Timers>::call()["+main-loop"].get_start_time() );
    }
    dev.synchronize();
    b200::flush_frames( dev, traj );

    std::size_t nl = 0, np = 0;
    orbc_size( dev.ctx, ORBC_LIPID, &nl ); orbc_size( dev.ctx, ORBC_PROTEIN, &np );
    unsigned long long launches = 0;
    for ( auto c : dev.ranks ) { unsigned long long l = 0; orbc_launch_count( c, &l ); launches += l; }
    std::stringstream msg;
    msg << param.nstep << " steps * " << nl + np << " particles on " << dev.world() << " B200 (" << launches << " kernel launches).";
    display_timing( std::cout, Service<Timers>::call()["+main-loop"].stop(), msg.str().c_str() );
    return 0;
}
