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


def make_ranks(st, world, **kw):
    import torch
    from openrbc_b200 import Simulation
    ndev = torch.cuda.device_count()
    sims = [Simulation(st, rank=r, world=world, device=r % ndev, **kw) for r in range(world)]
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


@pytest.mark.parametrize("name,world", [("vesicle_ico0", 2), ("vesicle_ico0", 4), ("sphere_r12", 4)])
def test_free_running_loop(name, world):
    """orbc_run_langevin with thermal noise over several rebuilds (Morton step included): the counter-based generator is
    keyed by the global slot, so a decomposed run follows the single-GPU trajectory."""
    from openrbc_b200 import Simulation
    st = load_state(name)
    one = Simulation(st, kBT=0.22)
    sims = make_ranks(st, world, kBT=0.22)
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
