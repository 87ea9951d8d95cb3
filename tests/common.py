"""Shared helpers for the test-suite: reference worlds (oracle/_ref), small systems, tolerances."""
import os
import tempfile

import numpy as np
import pytest

from oracle import ref as refmod

HAVE_REF = refmod.available("strict")
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_strict.so not built (needs /root/reference)")

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_meshdir = None


def mesh_prefix(subdiv):
    """Write an icosphere mesh in OpenRBC's text format and return its prefix."""
    global _meshdir
    from openrbc_b200.meshgen import icosphere, write_mesh
    if _meshdir is None:
        _meshdir = tempfile.mkdtemp(prefix="orbc_mesh_")
    prefix = os.path.join(_meshdir, f"ico{subdiv}")
    if not os.path.exists(prefix + ".vert.txt"):
        write_mesh(prefix, *icosphere(subdiv))
    return prefix


def ref_sphere(radius=20.0, extra=()):
    """Reference world: random lipid sphere (`-i lipid`, init_random.h:27) + 64 Lloyd iterations."""
    r = refmod.Ref("strict", threads=1, args=["-i", "lipid", *extra])
    r.init_lipid_sphere(radius)
    r.voronoi_init(64)
    return r


def ref_vesicle(subdiv=1, extra=()):
    """Reference world: membrane + cytoskeleton built by init_rbc.h from an icosphere mesh."""
    r = refmod.Ref("strict", threads=1, args=["-i", "trimesh", "-m", mesh_prefix(subdiv), *extra])
    r.init_trimesh()
    r.voronoi_init(64)
    return r


def rel_err(a, b):
    """max_i |a_i - b_i| / (|b_i| + rms(b)) over rows — the bound of SURVEY.md Appendix B.2."""
    a = np.asarray(a, np.float64).reshape(-1, 3)
    b = np.asarray(b, np.float64).reshape(-1, 3)
    if len(b) == 0:
        return 0.0
    rms = np.sqrt((b * b).sum(1).mean())
    den = np.linalg.norm(b, axis=1) + rms
    den[den == 0] = 1.0
    return float((np.linalg.norm(a - b, axis=1) / den).max())
