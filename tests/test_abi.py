"""CPU-side checks of the product boundary: the C-ABI library builds, loads and exports exactly the entry points that
include/orbc_b200.h declares; host-only helpers give the reference's numbers.  No GPU compute is attempted here."""
import os
import re

import numpy as np
import pytest

import openrbc_b200 as orbc
from openrbc_b200 import engine
from tests.common import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(orbc.library_path()):
        orbc.build_library()
    return orbc.load_library()


def test_header_and_library_agree(lib):
    header = open(os.path.join(ROOT, "include", "orbc_b200.h")).read()
    declared = set(re.findall(r"ORBC_API\s+[\w\s\*]+?\b(orbc_\w+)\s*\(", header))
    assert declared == set(engine.EXPORTS), declared ^ set(engine.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/orbc_b200.h but not exported"


def test_every_entry_point_cites_the_reference():
    header = open(os.path.join(ROOT, "include", "orbc_b200.h")).read()
    for needle in ("compute_pairwise_fused.h:238-320", "compute_bonded.h:89-146", "integrate_nh.h:29-37", "voronoi.h:77-86",
                   "voronoi.h:153-163", "cleanup.h:29-91", "compute_temperature.h:23-29", "constrain_volume.h:26-83", "trajectory.h:61-105",
                   "voronoi.h:54-75", "openrbc.cpp:88-133", "openrbc.cpp:229", "openrbc.cpp:189-256", "integrate_langevin.h:99-149"):
        assert needle in header, needle


def test_forcefield_table_matches_reference_bytes(lib):
    """orbc_forcefield_canonical (host code of the product) against the table dumped from the reference (golden)."""
    g = np.load(os.path.join(GOLDEN, "vesicle_ico0.npz"))
    assert orbc.forcefield_canonical().as_array().tobytes() == g["forcefield"].tobytes()


def test_zeta_update_host_helper(lib):
    import ctypes as C
    from oracle import port
    q1, q2 = C.c_float(0.0), C.c_float(0.0)
    a = lib.orbc_nh_zeta_update(0.03, C.byref(q1), 1e-2, 0.22, 1234.5, 3544)
    b = port.lib().orc_nh_zeta_update(C.c_float(0.03), C.byref(q2), C.c_double(1e-2), C.c_float(0.22), C.c_double(1234.5), 3544)
    assert a == b and q1.value == q2.value


def test_unfused_zeta_update_host_helper(lib):
    """verlet_nh_update's destructor (integrate_nh.h:74-77): its 1.5f * n * kBT is a float product, one ulp away from the fused
    kernel's; the library's host helper against the oracle's restatement over a few values."""
    import ctypes as C
    from oracle import port
    lib.orbc_nh_zeta_update_unfused.restype = C.c_float
    lib.orbc_nh_zeta_update_unfused.argtypes = [C.c_float, C.POINTER(C.c_float), C.c_double, C.c_float, C.c_double, C.c_long]
    for zeta, ke, n in ((0.03, 1234.5, 3544), (0.0, 75341.25, 174102), (-0.2, 1.0e6, 3205506)):
        q1, q2 = C.c_float(0.0), C.c_float(0.0)
        a = lib.orbc_nh_zeta_update_unfused(zeta, C.byref(q1), 1e-2, 0.22, ke, n)
        b = port.lib().orc_nh_zeta_update_unfused(C.c_float(zeta), C.byref(q2), C.c_double(1e-2), C.c_float(0.22), C.c_double(ke), n)
        assert a == b and q1.value == q2.value, (zeta, ke, n)


def test_no_cpu_fallback_without_device(lib):
    """Creating a context must fail loudly when no CUDA device is present (it must never fall back to the CPU)."""
    import ctypes as C
    ctx = C.c_void_p()
    rc = lib.orbc_create(C.byref(ctx), 0)
    if rc == 0:
        lib.orbc_destroy(ctx)
        pytest.skip("a CUDA device is present")
    assert rc < 0 and b"no CUDA device" in lib.orbc_last_error()
