"""Synthetic flat lipid-bilayer patches for the size sweep of BASELINE.json configs[3] ("Config 4").

Not in the reference: lipids sit on a jittered hexagonal lattice (spacing a = sqrt(2 / (sqrt(3) rho)), rho = 1.05 as
runtime_parameter.h:44) in z = const sheets of at most 1000 x 1000 length units, sheets stacked 10 units apart (further than
the largest centroid stencil, 9, so sheets never interact), directors +z, v = o = 0, no proteins.  The initial Voronoi
partition is blocks of 5 x 3 lattice sites (15 lipids per cell, the reference uses N/14 cells); the first rebuilds of a run
relax it (each rebuild is one Lloyd iteration: nearest-centroid assignment + centroid update).  Cells are stored in the
reference's Morton order (reorder_morton.h:25-42: quantum 0.5, offset 2000, 11/11/10 bits).

Host-side input generator only; nothing here runs on the device path.
"""
import numpy as np

RHO = 1.05
BX, BY = 5, 3          # lattice sites per cell along x / y


def _spread3(v):
    """bit_space3 of reorder_morton.h:25-31 (keeps bits 0..10)."""
    v = v.astype(np.uint64) & 0x7FF
    out = np.zeros_like(v)
    for b in range(11):
        out |= ((v >> b) & 1) << (3 * b)
    return out


def morton_keys(p):
    q = (2.0 * p.astype(np.float32)).astype(np.float64) + 2000.0
    u = q.astype(np.uint64)
    k = _spread3(u[:, 0]) | (_spread3(u[:, 1]) << 1) | (_spread3(u[:, 2]) << 2)
    return (k & 0xFFFFFFFF).astype(np.uint32)


def flat_patch_state(n_lipids, seed=1234):
    a = (2.0 / (3.0 ** 0.5 * RHO)) ** 0.5
    per_sheet_max = int(RHO * 1000 * 1000)
    n_sheets = max(1, -(-n_lipids // per_sheet_max))
    per_sheet = n_lipids // n_sheets
    ny_c = max(1, int(round((per_sheet / (BX * BY)) ** 0.5 * (BX / (BY * 3 ** 0.5 / 2)) ** 0.5)))   # cells along y so the sheet is ~square
    nx_c = max(1, per_sheet // (BX * BY * ny_c))
    nx, ny = nx_c * BX, ny_c * BY
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    x = (ix + 0.5 * (iy & 1)) * a
    y = iy * (a * 3 ** 0.5 / 2)
    x -= x.mean(); y -= y.mean()
    cell_in_sheet = (ix // BX) * ny_c + (iy // BY)
    rng = np.random.default_rng(seed)
    xs, cs = [], []
    z0 = -5.0 * (n_sheets - 1)
    for s in range(n_sheets):
        p = np.stack([x.ravel(), y.ravel(), np.full(x.size, z0 + 10.0 * s)], 1)
        p += rng.uniform(-0.05, 0.05, p.shape)
        xs.append(p.astype(np.float32)); cs.append(cell_in_sheet.ravel() + s * nx_c * ny_c)
    lx = np.concatenate(xs); cell = np.concatenate(cs)
    n_cells = n_sheets * nx_c * ny_c
    order = np.argsort(cell, kind="stable")
    lx, cell = lx[order], cell[order]
    cnt = np.bincount(cell, minlength=n_cells)
    cen = np.stack([np.bincount(cell, lx[:, d].astype(np.float64), n_cells) for d in range(3)], 1) / cnt[:, None]
    cen = cen.astype(np.float32)
    # cells in Morton order, particles sorted by (new) cell
    perm = np.argsort(morton_keys(cen), kind="stable")
    inv = np.empty(n_cells, np.int64); inv[perm] = np.arange(n_cells)
    cen = np.ascontiguousarray(cen[perm])
    newcell = inv[cell]
    order = np.argsort(newcell, kind="stable")
    lx = np.ascontiguousarray(lx[order])
    cs_l = np.zeros(n_cells + 1, np.int32); cs_l[1:] = np.cumsum(np.bincount(newcell, minlength=n_cells))
    n = len(lx)
    ln = np.zeros((n, 3), np.float32); ln[:, 2] = 1.0
    z3 = np.zeros((n, 3), np.float32)
    e3 = np.zeros((0, 3), np.float32)
    return dict(lx=lx, lv=z3, ln=ln, lo=z3.copy(), px=e3, pv=e3, pn=e3, po=e3, ptype=np.zeros(0, np.int32), ptag=np.zeros(0, np.int32),
                bonds=np.zeros((0, 3), np.int32), centroids=cen, cs_l=cs_l, cs_p=np.zeros(n_cells + 1, np.int32))
