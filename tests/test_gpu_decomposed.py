"""One cell split over several ranks (openrbc_b200/csrc/multi.cuh) against the single-GPU path on the same inputs.

The ranks are contexts of ONE process (one host thread each) and may share a device: peer stores, epoch-flag barriers,
owned-range kernels, migration and halo push are exactly the code a one-process-per-GPU run executes; only the pointer
exchange differs (raw pointers instead of CUDA IPC handles).  The decomposition must not change results: integer structures
equal, per-particle outputs equal up to the order of the few atomically accumulated protein->lipid reactions.
"""
import os
import threading

import numpy as np
import pytest

from tests.common import GOLDEN, rel_err

pytestmark = pytest.mark.gpu


def load_state(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return {k[3:]: v for k, v in g.items() if k.startswith("in_")}


def make_ranks(st, world, opts=None, **kw):
    import torch
    from openrbc_b200 import Simulation
    ndev = torch.cuda.device_count()
    sims = [Simulation(st, rank=r, world=world, device=r % ndev, **kw) for r in range(world)]
    for s in sims:
        for k, v in (opts or {}).items():
            s.set_option(k, v)
    blobs = [s.mg_export() for s in sims]
    for s in sims:
        s.mg_connect(blobs)
    return sims


def on_all(sims, fn):
    """Run fn(sim) for every rank concurrently (the ranks wait for each other on the device)."""
    errs = []

    def work(s):
        try:
            fn(s)
            s.synchronize()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(s,)) for s in sims]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in th), "a rank is stuck"
    if errs:
        raise errs[0]


def gathered(sims, s, fields):
    """Owned slots of every rank, concatenated in rank order = the whole container in global slot order."""
    parts = {f: [] for f in fields}
    end_prev = 0
    for sim in sims:
        b, e = sim.owned_range(s)
        assert b == end_prev, "owned ranges must tile the container"
        end_prev = e
        d = sim.download(s, fields)
        for f in fields:
            parts[f].append(d[f][b:e])
    assert end_prev == sims[0].size(s)
    return {f: np.concatenate(parts[f]) for f in fields}


@pytest.mark.parametrize("name", ["vesicle_ico0", "sphere_r12"])
@pytest.mark.parametrize("world", [2, 3])
def test_forces_and_rebuild_call_by_call(name, world):
    from openrbc_b200 import Simulation
    st = load_state(name)
    one = Simulation(st, kBT=0.0)
    sims = make_ranks(st, world, kBT=0.0)
    # forces on the uploaded state: every rank computes its own cells from the (complete) uploaded halo
    one.compute_pairwise_fused(); one.compute_bonded()
    on_all(sims, lambda s: (s.compute_pairwise_fused(), s.compute_bonded()))
    for s in (0, 1):
        ref, got = one.download(s, "ft"), gathered(sims, s, "ft")
        assert rel_err(got["f"], ref["f"]) < 2e-6 and rel_err(got["t"], ref["t"]) < 2e-6
    # one Langevin step (halo push), then a rebuild with a Morton renumbering of the cells (migration between ranks)
    for sim in [one] + sims:
        sim.nstep = 24
    one.verlet_langevin(); one.rebuild()
    on_all(sims, lambda s: (s.verlet_langevin(), s.rebuild()))
    for what in ("centroids", "cell_start_l", "cell_start_p", "morton_keys"):
        ref = one.dump(what)
        for sim in sims:
            np.testing.assert_array_equal(sim.dump(what), ref, err_msg=f"{what} rank {sim.rank}")
    for s in (0, 1):
        ref, got = one.download(s, "xvno", ids=True), gathered(sims, s, "xvno")
        for f in "xvno":
            np.testing.assert_allclose(got[f], ref[f], rtol=0, atol=1e-6 * (1 + np.abs(ref[f]).max(initial=0.0)), err_msg=f)
    # forces on the migrated state: halo copies and the tag -> slot map must have followed the particles
    one.compute_pairwise_fused(); one.compute_bonded()
    on_all(sims, lambda s: (s.compute_pairwise_fused(), s.compute_bonded()))
    for s in (0, 1):
        ref, got = one.download(s, "ft"), gathered(sims, s, "ft")
        assert rel_err(got["f"], ref["f"]) < 1e-5 and rel_err(got["t"], ref["t"]) < 1e-5
    for sim in [one] + sims:
        sim.close()


@pytest.mark.parametrize("lists", [2, 0])
@pytest.mark.parametrize("name,world", [("vesicle_ico0", 2), ("vesicle_ico0", 4), ("sphere_r12", 4)])
def test_free_running_loop(name, world, lists):
    """orbc_run_langevin with thermal noise over several rebuilds (Morton step included): the counter-based generator is
    keyed by the global slot, so a decomposed run follows the single-GPU trajectory."""
    from openrbc_b200 import Simulation
    st = load_state(name)
    one = Simulation(st, kBT=0.22)
    sims = make_ranks(st, world, opts={"nl_reuse": lists}, kBT=0.22)  # hit lists forced on / off (automatic: on up to four ranks, off beyond)
    for sim in [one] + sims:
        sim.nstep = 20
    one.run_langevin(8)
    on_all(sims, lambda s: s.run_langevin(8))
    for what in ("cell_start_l", "cell_start_p"):
        ref = one.dump(what)
        for sim in sims:
            np.testing.assert_array_equal(sim.dump(what), ref, err_msg=f"{what} rank {sim.rank}")
    for s in (0, 1):
        ref, got = one.download(s, "xvno"), gathered(sims, s, "xvno")
        for f in "xvno":
            np.testing.assert_allclose(got[f], ref[f], rtol=0, atol=2e-5 * (1 + np.abs(ref[f]).max(initial=0.0)), err_msg=f)
    t = sum(sim.compute_temperature() for sim in sims)
    assert abs(t - one.compute_temperature()) < 1e-6 * t
    for sim in [one] + sims:
        sim.close()


@pytest.mark.parametrize("world", [2, 3])
def test_delete_lipid(world):
    """cleanup.h:29-91 on a decomposed run: same survivors, same slots, same partition as on one GPU."""
    from openrbc_b200 import Simulation
    st = load_state("vesicle_ico0")
    one = Simulation(st, kBT=0.22)
    sims = make_ranks(st, world, kBT=0.22)
    one.run_langevin(3)
    on_all(sims, lambda s: s.run_langevin(3))
    n_one = one.delete_lipid(1.3)
    out = {}
    on_all(sims, lambda s: out.__setitem__(s.rank, s.delete_lipid(1.3)))
    assert 0 < n_one < len(st["lx"]) and all(v == n_one for v in out.values()), (n_one, out)
    ref = one.dump("cell_start_l")
    for sim in sims:
        assert sim.size(0) == n_one
        np.testing.assert_array_equal(sim.dump("cell_start_l"), ref)
    ref, got = one.download(0, "xvno"), gathered(sims, 0, "xvno")
    for f in "xvno":
        np.testing.assert_allclose(got[f], ref[f], rtol=0, atol=2e-6 * (1 + np.abs(ref[f]).max(initial=0.0)), err_msg=f)
    # and the loop goes on from there
    one.run_langevin(3)
    on_all(sims, lambda s: s.run_langevin(3))
    ref, got = one.download(0, "xvno"), gathered(sims, 0, "xvno")
    for f in "xvno":
        np.testing.assert_allclose(got[f], ref[f], rtol=0, atol=2e-5 * (1 + np.abs(ref[f]).max(initial=0.0)), err_msg=f)
    for sim in [one] + sims:
        sim.close()


@pytest.mark.parametrize("name,world", [("vesicle_ico0", 2), ("vesicle_ico0", 4)])
def test_nose_hoover_loop(name, world):
    """orbc_run_nh (integrate_nh.h fused pair) on a decomposed run: the partial kinetic energies of the ranks are exchanged and
    summed in rank order, so every rank carries the same friction zeta as the single-GPU run."""
    from openrbc_b200 import Simulation
    st = load_state(name)
    one = Simulation(st, kBT=0.22)
    sims = make_ranks(st, world, kBT=0.22)
    for sim in [one] + sims:
        sim.nstep = 22
        sim.zeta = 0.01
    one.run_nh(6)
    on_all(sims, lambda s: s.run_nh(6))
    for sim in sims:
        assert abs(sim.zeta - one.zeta) <= 1e-6 * abs(one.zeta) and sim.Q.value == one.Q.value, (sim.zeta, one.zeta)
    for s in (0, 1):
        ref, got = one.download(s, "xvno"), gathered(sims, s, "xvno")
        for f in "xvno":
            np.testing.assert_allclose(got[f], ref[f], rtol=0, atol=2e-5 * (1 + np.abs(ref[f]).max(initial=0.0)), err_msg=f)
    # call by call: the kinetic energy returned on every rank is the global one
    ke_one = one.nh_initial_fused()
    out = {}
    on_all(sims, lambda s: out.__setitem__(s.rank, s.nh_initial_fused()))
    assert all(abs(v - ke_one) <= 1e-9 * ke_one for v in out.values()), (ke_one, out)
    for sim in [one] + sims:
        sim.close()


@pytest.mark.parametrize("world", [2, 4])
def test_volume_constraint(world):
    """constrain_volume.h:26-83 on a decomposed run (BASELINE configs[2]): every rank computes the normals and the volume share of
    its own cells, the shares are exchanged and summed in rank order — same volume and same forces as on one GPU, call by
    call (twice: the scratch normals persist) and inside the free-running loop."""
    from openrbc_b200 import Simulation
    st = load_state("vesicle_ico0")
    one = Simulation(st, kBT=0.22)
    sims = make_ranks(st, world, kBT=0.22)
    for k in range(2):
        v_one = one.constrain_volume(3.15, 0.05)
        out = {}
        on_all(sims, lambda s: out.__setitem__(s.rank, s.constrain_volume(3.15, 0.05)))
        assert all(abs(v - v_one) <= 1e-6 * abs(v_one) for v in out.values()), (v_one, out)
        for s in (0, 1):
            ref, got = one.download(s, "f"), gathered(sims, s, "f")
            assert rel_err(got["f"], ref["f"]) < 1e-6, (k, s)
    for sim in [one] + sims:
        sim.clear_force()
        sim.set_volume_constraint(True, 3.15, 0.05)
        sim.nstep = 22
    one.run_langevin(6)
    on_all(sims, lambda s: s.run_langevin(6))
    for s in (0, 1):
        ref, got = one.download(s, "xvno"), gathered(sims, s, "xvno")
        for f in "xvno":
            np.testing.assert_allclose(got[f], ref[f], rtol=0, atol=2e-5 * (1 + np.abs(ref[f]).max(initial=0.0)), err_msg=f)
    for sim in [one] + sims:
        sim.close()


@pytest.mark.parametrize("name,world", [("vesicle_ico0", 2), ("vesicle_ico0", 3), ("sphere_r12", 4)])
def test_minimisation_loop(name, world):
    """orbc_run_minimize (openrbc.cpp:88-133) decomposed: a rebuild with Morton renumbering on every iteration."""
    from openrbc_b200 import Simulation
    st = load_state(name)
    one = Simulation(st, kBT=0.0)
    sims = make_ranks(st, world, kBT=0.0)
    one.run_minimize(3)
    on_all(sims, lambda s: s.run_minimize(3))
    for what in ("cell_start_l", "cell_start_p", "centroids"):
        ref = one.dump(what)
        for sim in sims:
            np.testing.assert_array_equal(sim.dump(what), ref, err_msg=f"{what} rank {sim.rank}")
    for s in (0, 1):
        ref, got = one.download(s, "xnft"), gathered(sims, s, "xnft")
        for f in "xn":
            np.testing.assert_allclose(got[f], ref[f], rtol=0, atol=2e-6 * (1 + np.abs(ref[f]).max(initial=0.0)), err_msg=f)
        assert rel_err(got["f"], ref["f"]) < 1e-5 and rel_err(got["t"], ref["t"]) < 1e-5
    for sim in [one] + sims:
        sim.close()


def test_frames_of_the_ranks_combine():
    """save_frame on a decomposed run: each rank's image holds the titles and its own slots, zeros elsewhere; OR-ed together
    they are the single-GPU frame."""
    from openrbc_b200 import Simulation
    st = load_state("vesicle_ico0")
    one = Simulation(st, kBT=0.22)
    sims = make_ranks(st, 3, kBT=0.22)
    for sim in [one] + sims:
        sim.nstep = 22
    one.run_langevin(4)
    on_all(sims, lambda s: s.run_langevin(4))
    want = one.save_frame(15)
    out = {}
    on_all(sims, lambda s: out.__setitem__(s.rank, s.save_frame(15).copy()))
    merged = np.zeros_like(want)
    for r in sorted(out):
        merged |= out[r]
    # positions / velocities follow the single-GPU trajectory to rounding, everything else is equal: compare section by section
    n = one.size(0) + one.size(1)
    off = 36 + 8 * n
    np.testing.assert_array_equal(merged[:off], want[:off])                      # titles, nstep, NATOM, IDENTITY
    for width in (12, 12, 12, 4):                                                # POSITION VELOCITY ROTATION VORONOI
        np.testing.assert_array_equal(merged[off:off + 8], want[off:off + 8])
        a = merged[off + 8:off + 8 + width * n]; b = want[off + 8:off + 8 + width * n]
        if width == 4:
            np.testing.assert_array_equal(a, b)
        else:
            np.testing.assert_allclose(a.view(np.float32), b.view(np.float32), rtol=0, atol=2e-5 * (1 + np.abs(b.view(np.float32)).max()))
        off += 8 + width * n
    np.testing.assert_array_equal(merged[off:], want[off:])
    for sim in [one] + sims:
        sim.close()


@pytest.mark.parametrize("world", [2, 3])
def test_reupload_of_owned_rows_only(world):
    """orbc_upload_range: connected ranks re-upload only their own rows; orbc_mg_export fetches the halo copies from their owners
    over the peer mappings.  The run that follows must equal the run after whole uploads (and the single-GPU run)."""
    from openrbc_b200 import Simulation
    st = load_state("vesicle_ico0")
    one = Simulation(st, kBT=0.0)
    sims = make_ranks(st, world, kBT=0.0)
    on_all(sims, lambda s: s.run_langevin(3))                 # leave the first state behind: slots, ranges and halos have moved
    garbage = dict(st)
    for k in ("lx", "ln", "px", "pn"):
        garbage[k] = st[k] + np.float32(3.0)                  # what the other ranks' rows must NOT be taken from
    def reload(s):
        mine = dict(garbage)
        cb, ce = __import__("openrbc_b200").engine.cell_range(len(st["centroids"]), s.rank, world)
        for p, cs in (("l", st["cs_l"]), ("p", st["cs_p"])):
            b, e = int(cs[cb]), int(cs[ce])
            for f in "xvno":
                a = np.array(mine[p + f], copy=True); a[b:e] = st[p + f][b:e]; mine[p + f] = a
        s.upload(mine, owned_only=True)
        s.mg_export()
        s.nstep = 0
    on_all(sims, reload)
    one.run_langevin(4)
    on_all(sims, lambda s: s.run_langevin(4))
    for s in (0, 1):
        ref, got = one.download(s, "xvno"), gathered(sims, s, "xvno")
        for f in "xvno":
            np.testing.assert_allclose(got[f], ref[f], rtol=0, atol=2e-6 * (1 + np.abs(ref[f]).max(initial=0.0)), err_msg=f)
    for sim in [one] + sims:
        sim.close()


@pytest.mark.parametrize("name", ["vesicle_ico0", "sphere_r12"])
def test_launch_bounds_below_the_container_size(name):
    """With the production slack (8192) the launch bound of the owned-particle kernels equals the container size on these small
    systems, which hides any kernel that is launched over the bound but indexed from slot 0.  A small slack makes rank 1 own
    slots BEYOND its launch bound: clear_force, the call-by-call Nose-Hoover kernels and the fused minimiser must still cover them."""
    from openrbc_b200 import Simulation
    st = load_state(name)
    one = Simulation(st, kBT=0.22)
    sims = make_ranks(st, 2, opts={"debug_own_slack": 32}, kBT=0.22)
    n_l = len(st["lx"])
    assert sims[1].owned_range(0)[1] == n_l and n_l // 2 + n_l // 8 + 32 < n_l           # rank 1 owns the tail, the bound stops short of it
    # dirty forces everywhere, then clear_force: every slot of every rank must be clean afterwards
    for sim in [one] + sims:
        for s in (0, 1):
            if sim.size(s):
                sim.set_field(s, "f", np.full((sim.size(s), 3), 7.0, np.float32)); sim.set_field(s, "t", np.full((sim.size(s), 3), -3.0, np.float32))
    one.clear_force()
    on_all(sims, lambda s: s.clear_force())
    for sim in sims:
        for s in (0, 1):
            d = sim.download(s, "ft")
            b, e = sim.owned_range(s)
            assert not d["f"][b:e].any() and not d["t"][b:e].any()
    # one Nose-Hoover step call by call, then two minimiser iterations
    def nh_step(s):
        s.nh_initial_fused(); s.rebuild(); s.compute_pairwise_fused(); s.compute_bonded(); s.nh_final_fused(); s.nstep += 1
    nh_step(one)
    on_all(sims, nh_step)
    for s in (0, 1):
        ref, got = one.download(s, "xvno"), gathered(sims, s, "xvno")
        for f in "xvno":
            np.testing.assert_allclose(got[f], ref[f], rtol=0, atol=2e-6 * (1 + np.abs(ref[f]).max(initial=0.0)), err_msg=f)
    assert all(abs(s.zeta - one.zeta) <= 1e-6 * abs(one.zeta) for s in sims)
    one.run_minimize(2)
    on_all(sims, lambda s: s.run_minimize(2))
    for s in (0, 1):
        ref, got = one.download(s, "xn"), gathered(sims, s, "xn")
        for f in "xn":
            np.testing.assert_allclose(got[f], ref[f], rtol=0, atol=2e-6 * (1 + np.abs(ref[f]).max(initial=0.0)), err_msg=f)
    for sim in [one] + sims:
        sim.close()


def test_delete_lipid_that_deletes_nothing_leaves_the_partition_alone():
    """cleanup.h:62: `if ( size_new < n )` — without a stray the reference neither compacts nor re-partitions, and the next
    voronoi.update averages over the OLD membership.  A decomposed run must not re-partition either (it used to)."""
    from openrbc_b200 import Simulation
    st = load_state("vesicle_ico0")
    one = Simulation(st, kBT=0.22)
    sims = make_ranks(st, 2, kBT=0.22)
    one.run_langevin(1)                                       # one step: the partition is now stale (rebuild at step 0 only)
    on_all(sims, lambda s: s.run_langevin(1))
    before = one.dump("cell_start_l").copy()
    assert one.delete_lipid(1e9) == len(st["lx"])
    out = {}
    on_all(sims, lambda s: out.__setitem__(s.rank, s.delete_lipid(1e9)))
    assert all(v == len(st["lx"]) for v in out.values())
    for sim in [one] + sims:
        np.testing.assert_array_equal(sim.dump("cell_start_l"), before)
    one.run_langevin(3)
    on_all(sims, lambda s: s.run_langevin(3))
    np.testing.assert_array_equal(sims[0].dump("cell_start_l"), one.dump("cell_start_l"))
    ref, got = one.download(0, "xvno"), gathered(sims, 0, "xvno")
    for f in "xvno":
        np.testing.assert_allclose(got[f], ref[f], rtol=0, atol=2e-5 * (1 + np.abs(ref[f]).max(initial=0.0)), err_msg=f)
    for sim in [one] + sims:
        sim.close()
