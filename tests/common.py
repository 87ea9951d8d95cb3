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


# offsets into the float table of oracle.ref.Ref.forcefield() (layout of include/orbc_b200.h: orbc_forcefield)
FF_CUTSQLP, FF_CUTSQPP, FF_LJ_CUTSQ, FF_CUTSQLL = 18, 78, 150, 267


def branch_hits(st, g, ff):
    """Pairs the reference's driver evaluates in each branch of compute_pairwise_fused.h:90-236 on the state `st` (brute force over
    all pairs, restricted to the driver's candidate sets: cell of j inside the r<6 / r<8 / r<9 CENTROID stencil of the cell of i,
    stencils taken from the CSR arrays st{6,8,9}_{ptr,idx} of `g`).  Unordered pairs; small systems only (dense N x N)."""
    ff = np.asarray(ff, np.float64)
    nc = len(st["cs_l"]) - 1

    def adjacency(k):
        a = np.zeros((nc, nc), bool)
        ptr, idx = g[f"st{k}_ptr"], g[f"st{k}_idx"]
        for c in range(nc):
            a[c, idx[ptr[c]:ptr[c + 1]]] = True
        return a

    def d2(a, b):
        # fp32 like the driver (the counts sit far from the cutoffs in the fixtures, rounding does not move them)
        d = a[:, None, :].astype(np.float32) - b[None, :, :].astype(np.float32)
        return (d * d).sum(-1)

    cell_l = np.repeat(np.arange(nc), np.diff(st["cs_l"]))
    cell_p = np.repeat(np.arange(nc), np.diff(st["cs_p"]))
    ty = np.asarray(st["ptype"], np.int64)
    out = {}
    r2 = d2(st["lx"], st["lx"])
    ok = adjacency(6)[cell_l][:, cell_l] & (r2 > 1e-5)
    out["ll"] = int((np.triu(ok & (r2 < ff[FF_CUTSQLL]), 1)).sum())
    if len(ty):
        r2 = d2(st["px"], st["lx"])
        ok = adjacency(8)[cell_p][:, cell_l] & (r2 > 1e-5)
        poly = ok & (r2 < ff[FF_CUTSQLP + ty][:, None])
        lj = ok & ~poly & (r2 < ff[FF_LJ_CUTSQ + ty][:, None])
        out["pl_poly"], out["pl_lj"] = int(poly.sum()), int(lj.sum())
        r2 = d2(st["px"], st["px"])
        t12 = ty[:, None] + 6 * ty[None, :]
        ok = adjacency(9)[cell_p][:, cell_p] & (r2 > 1e-5)
        rep = np.triu(ok & (r2 < ff[FF_CUTSQPP + t12]), 1)
        lj = np.triu(ok & ~(r2 < ff[FF_CUTSQPP + t12]) & (r2 < ff[FF_LJ_CUTSQ + t12]), 1)
        out["pp_rep"], out["pp_lj"] = int(rep.sum()), int(lj.sum())
        out["pp_rep_type_pairs"] = len(set(map(tuple, np.sort(np.stack([ty[np.nonzero(rep)[0]], ty[np.nonzero(rep)[1]]], 1), 1).tolist())))
        out["pp_lj_type_pairs"] = len(set(map(tuple, np.sort(np.stack([ty[np.nonzero(lj)[0]], ty[np.nonzero(lj)[1]]], 1), 1).tolist())))
    return out
