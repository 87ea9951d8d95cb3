"""TEST INFRASTRUCTURE — ctypes wrapper + step composition for the C restatement (oracle/orbc_oracle.c).

`World` strings the port's functions together in the order of the reference's main loop
(src/openrbc.cpp:189-256) on plain numpy arrays.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import this module; the product never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborbc_oracle.so")


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("orbc_oracle.c", "orbc_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
    return _SO


class ForceField(C.Structure):
    _fields_ = [(n, C.c_float * k) for n, k in (
        ("mass", 6), ("radius", 6), ("cutlp", 6), ("cutsqlp", 6), ("replp", 6), ("attlp", 6), ("alphalp", 6),
        ("cutpp", 36), ("cutsqpp", 36), ("reppp", 36), ("lj_cutsq", 36), ("lj_lj1", 36), ("lj_lj2", 36),
        ("r0", 4), ("K", 4))] + [(n, C.c_float) for n in ("cutll", "cutsqll", "repll", "attll", "alphall")]

    def as_array(self):
        return np.frombuffer(bytes(self), np.float32).copy()


class MT19937(C.Structure):
    _fields_ = [("idata", C.c_uint32 * 624), ("rdata", C.c_float * 624), ("state", C.c_uint32 * 624), ("ipos", C.c_int), ("rpos", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.orc_morton_encode.restype = C.c_uint32
        L.orc_morton_encode.argtypes = [C.c_float] * 3
        L.orc_temperature.restype = C.c_double
        L.orc_constrain_volume.restype = C.c_float
        L.orc_delete_lipid_mask.restype = C.c_long
        L.orc_nh_zeta_update.restype = C.c_float
        L.orc_nh_zeta_update_unfused.restype = C.c_float
        L.orc_mt_uint.restype = C.c_uint32
        L.orc_mt_u01.restype = C.c_float
        L.orc_uint2u11.restype = C.c_float
        L.orc_uint2u11.argtypes = [C.c_uint32]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f3(a):
    return np.ascontiguousarray(a, np.float32).reshape(-1, 3)


def i1(a):
    return np.ascontiguousarray(a, np.int32).reshape(-1)


def forcefield():
    ff = ForceField()
    lib().orc_forcefield_canonical(C.byref(ff))
    return ff


# ---- thin functional wrappers --------------------------------------------------------------------
def update_centroid(cell_start, x):
    cs = i1(cell_start)
    out = np.empty((len(cs) - 1, 3), np.float32)
    lib().orc_update_centroid(len(cs) - 1, _p(cs), _p(f3(x)), _p(out))
    return out


def morton_encode(x, y, z):
    return lib().orc_morton_encode(float(x), float(y), float(z))


def morton_perm(pts):
    pts = f3(pts)
    perm = np.empty(len(pts), np.int32)
    keys = np.empty(len(pts), np.uint32)
    lib().orc_morton_perm(len(pts), _p(pts), _p(perm), _p(keys))
    return perm, keys


def assign_nearest(x, centroids):
    x = f3(x)
    c = f3(centroids)
    aff = np.empty(len(x), np.int32)
    tie = np.empty(len(x), np.int32)
    nt = lib().orc_assign_nearest(C.c_long(len(x)), _p(x), len(c), _p(c), _p(aff), _p(tie))
    return aff, tie.astype(bool), nt


def partition(aff, n_cells):
    aff = i1(aff)
    cs = np.empty(n_cells + 1, np.int32)
    cells = np.empty(len(aff), np.int32)
    li = np.empty(len(aff), np.int32)
    lib().orc_partition(C.c_long(len(aff)), n_cells, _p(aff), _p(cs), _p(cells), _p(li))
    return cs, cells, li


def voronoi_init(x, n_cells, n_iterate=64, others=()):
    """Restatement of VoronoiDiagram::init (voronoi.h:54-75): initial guess = every delta-th particle from delta / 2 (wrapping),
    then n_iterate rounds of Morton sort of the centroids, partition, centroid update through `cells`; the container (x and the
    arrays in `others`) is reordered at the rounds k = 0, 1, 2, 4, 8, ... and once more at the end (:70,74).
    Returns (centroids, cell_start, x, others, ties)."""
    x = f3(x).copy()
    others = [np.ascontiguousarray(a).copy() for a in others]
    n = len(x)
    delta = n // n_cells
    idx = (delta // 2 + np.arange(n_cells, dtype=np.int64) * delta) % n
    c = np.ascontiguousarray(x[idx])
    ties = 0
    for k in range(n_iterate):
        perm, _ = morton_perm(c)
        c = np.ascontiguousarray(c[perm])
        aff, _, nt = assign_nearest(x, c)
        ties += nt
        cs, cells, _ = partition(aff, n_cells)
        c = update_centroid(cs, np.ascontiguousarray(x[cells]))      # the sum runs over cells[j] in ascending j (:131-134)
        if (k & (~k + 1)) == k:
            x = np.ascontiguousarray(x[cells]); others = [np.ascontiguousarray(a[cells]) for a in others]
    x = np.ascontiguousarray(x[cells]); others = [np.ascontiguousarray(a[cells]) for a in others]
    return c, cs, x, others, ties


def stencil(centroids, cell, rmax, cap=4096):
    c = f3(centroids)
    out = np.empty(cap, np.int32)
    n = lib().orc_stencil(len(c), _p(c), int(cell), C.c_float(rmax), _p(out), cap)
    return out[:n].copy()


def langevin_noise(mt, n):
    out = np.empty((n, 3), np.float32)
    lib().orc_langevin_noise(C.byref(mt), C.c_long(n), _p(out))
    return out


def mt_init(seed):
    g = MT19937()
    lib().orc_mt_init(C.byref(g), C.c_uint32(seed & 0xFFFFFFFF))
    return g


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return list(o)


def philox_noise(seed, step, species, n):
    out = np.empty((n, 3), np.float32)
    lib().orc_philox_noise(C.c_uint64(seed), C.c_uint32(step), C.c_uint32(species), C.c_long(n), _p(out))
    return out


BOX_LO, BOX_HI = -1000.0, 1000.0  # runtime_parameter.h:111-115


class World:
    """Port-side mirror of the reference's containers + Voronoi state (numpy arrays)."""

    def __init__(self, st, dt=1e-2, kBT=0.22, eta=0.01):
        self.ff = forcefield()
        g = lambda k: f3(st[k]).copy()
        self.lx, self.lv, self.ln, self.lo = g("lx"), g("lv"), g("ln"), g("lo")
        self.px, self.pv, self.pn, self.po = g("px"), g("pv"), g("pn"), g("po")
        self.lf = np.zeros_like(self.lx); self.lt = np.zeros_like(self.lx)
        self.pf = np.zeros_like(self.px); self.pt = np.zeros_like(self.px)
        self.ptype = i1(st["ptype"]).copy()
        self.ptag = i1(st["ptag"]).copy()
        self.bonds = np.ascontiguousarray(st["bonds"], np.int32).reshape(-1, 3).copy()
        self.centroids = f3(st["centroids"]).copy()
        self.cs_l = i1(st["cs_l"]).copy()
        self.cs_p = i1(st["cs_p"]).copy()
        self.n_cells = len(self.centroids)
        self.dt, self.kBT, self.eta = dt, kBT, eta
        self.zeta, self.Q = 0.0, C.c_float(0.0)
        self.box = (BOX_LO, BOX_HI)
        self.nstep = 0
        self.freq_sort_ctrd = 24
        self.cell_normal = np.zeros((self.n_cells, 3), np.float32)
        self.counters = np.zeros(8, np.int64)
        self.ties = 0
        self._tag2idx()

    def _tag2idx(self):
        size = int(self.ptag.max()) + 1 if len(self.ptag) else 1
        self.tag2idx = np.empty(size, np.int32)
        lib().orc_build_tag2idx(C.c_long(len(self.ptag)), _p(self.ptag), _p(self.tag2idx), C.c_long(size))

    # voronoi.h:77-86
    def voronoi_update(self):
        self.centroids = update_centroid(self.cs_l, self.lx)
        if self.nstep % self.freq_sort_ctrd == 0:
            perm, _ = morton_perm(self.centroids)
            self.centroids = np.ascontiguousarray(self.centroids[perm])
            self.last_perm = perm

    # voronoi.h:153-163 + reorder.h:73-149
    def cell_update(self, s):
        x = self.lx if s == 0 else self.px
        aff, tie, nt = assign_nearest(x, self.centroids)
        self.ties += nt
        cs, cells, li = partition(aff, self.n_cells)
        if s == 0:
            self.cs_l, self.cells_l, self.aff_l, self.tie_l = cs, cells, aff, tie
            self.lx, self.lv, self.ln, self.lo = (np.ascontiguousarray(a[cells]) for a in (self.lx, self.lv, self.ln, self.lo))
        else:
            self.cs_p, self.cells_p, self.aff_p, self.tie_p = cs, cells, aff, tie
            self.px, self.pv, self.pn, self.po = (np.ascontiguousarray(a[cells]) for a in (self.px, self.pv, self.pn, self.po))
            self.ptype = np.ascontiguousarray(self.ptype[cells]); self.ptag = np.ascontiguousarray(self.ptag[cells])
            self._tag2idx()

    def rebuild(self):
        self.voronoi_update(); self.cell_update(0); self.cell_update(1)

    def clear_force(self):
        for a in (self.lf, self.lt, self.pf, self.pt):
            a[...] = 0

    def compute_pairwise_fused(self):
        lib().orc_pairwise_fused(C.byref(self.ff), self.n_cells, _p(self.centroids),
                                 C.c_long(len(self.lx)), _p(self.lx), _p(self.ln), _p(self.cs_l), _p(self.lf), _p(self.lt),
                                 C.c_long(len(self.px)), _p(self.px), _p(self.pn), _p(self.ptype), _p(self.cs_p), _p(self.pf), _p(self.pt),
                                 _p(self.counters))

    def compute_bonded(self):
        lib().orc_bonded(C.byref(self.ff), C.c_long(len(self.bonds)), _p(self.bonds), _p(self.tag2idx), _p(self.px), _p(self.pf))

    def _each(self):
        yield (self.lx, self.lv, self.lf, self.ln, self.lo, self.lt, None)
        yield (self.px, self.pv, self.pf, self.pn, self.po, self.pt, self.ptype)

    def post_torque(self):
        for x, v, f, n, o, t, ty in self._each():
            lib().orc_post_torque(C.c_long(len(x)), _p(n), _p(t))

    def bounce_back(self):
        for x, v, f, n, o, t, ty in self._each():
            lib().orc_bounce_back(C.c_long(len(x)), _p(x), _p(v), C.c_double(self.box[0]), C.c_double(self.box[1]))

    def verlet_langevin(self, noise_l=None, noise_p=None):
        for (x, v, f, n, o, t, ty), nz in zip(self._each(), (noise_l, noise_p)):
            nz = None if nz is None else f3(nz)
            lib().orc_verlet_langevin(C.byref(self.ff), C.c_long(len(x)), _p(x), _p(v), _p(f), _p(n), _p(o), _p(t), _p(ty),
                                      C.c_double(self.dt), C.c_float(self.eta), C.c_float(self.kBT), _p(nz))

    def nh_initial_fused(self):
        ke = C.c_double(0.0); n = 0
        for x, v, f, nn, o, t, ty in self._each():
            lib().orc_nh_initial_fused(C.byref(self.ff), C.c_long(len(x)), _p(x), _p(v), _p(f), _p(nn), _p(o), _p(t), _p(ty),
                                       C.c_double(self.dt), C.c_float(self.zeta), C.c_double(self.box[0]), C.c_double(self.box[1]), C.byref(ke))
            n += len(x)
        self.last_ke = ke.value
        self.zeta = lib().orc_nh_zeta_update(C.c_float(self.zeta), C.byref(self.Q), C.c_double(self.dt), C.c_float(self.kBT), ke, n)
        return ke.value

    def nh_final_fused(self):
        ke = C.c_double(0.0); n = 0
        for x, v, f, nn, o, t, ty in self._each():
            lib().orc_nh_final_fused(C.byref(self.ff), C.c_long(len(x)), _p(v), _p(f), _p(nn), _p(o), _p(t), _p(ty),
                                     C.c_double(self.dt), C.c_float(self.zeta), C.byref(ke))
            n += len(x)
        self.last_ke = ke.value
        self.zeta = lib().orc_nh_zeta_update(C.c_float(self.zeta), C.byref(self.Q), C.c_double(self.dt), C.c_float(self.kBT), ke, n)
        return ke.value

    def nh_final(self):                                          # verlet_nh_final (unfused), integrate_nh.h:155-176
        for x, v, f, nn, o, t, ty in self._each():
            lib().orc_nh_final(C.byref(self.ff), C.c_long(len(x)), _p(v), _p(f), _p(o), _p(t), _p(ty), C.c_double(self.dt), C.c_float(self.zeta))

    def nh_update(self):                                         # verlet_nh_update + its destructor, integrate_nh.h:66-94
        ke = C.c_double(0.0); n = 0
        for x, v, f, nn, o, t, ty in self._each():
            lib().orc_nh_update(C.byref(self.ff), C.c_long(len(x)), _p(v), _p(ty), C.byref(ke))
            n += len(x)
        self.last_ke = ke.value
        self.zeta = lib().orc_nh_zeta_update_unfused(C.c_float(self.zeta), C.byref(self.Q), C.c_double(self.dt), C.c_float(self.kBT), ke, n)
        return ke.value

    def opt_move(self, dr_opt=5e-2, dn_opt=5e-2):
        for x, v, f, n, o, t, ty in self._each():
            lib().orc_opt_move(C.byref(self.ff), C.c_long(len(x)), _p(x), _p(n), _p(f), _p(t), _p(ty),
                               C.c_double(self.dt), C.c_double(dr_opt), C.c_double(dn_opt))

    def compute_temperature(self):
        return lib().orc_temperature(C.byref(self.ff), C.c_long(len(self.lv)), _p(self.lv), C.c_long(len(self.pv)), _p(self.pv), _p(self.ptype))

    def constrain_volume(self, target, strength):
        return lib().orc_constrain_volume(C.byref(self.ff), self.n_cells, _p(self.centroids), _p(self.cell_normal),
                                          C.c_long(len(self.lx)), _p(self.ln), _p(self.cs_l), _p(self.lf),
                                          C.c_long(len(self.px)), _p(self.ptype), _p(self.cs_p), _p(self.pf),
                                          C.c_float(target), C.c_float(strength))

    def delete_lipid(self, stray_tolerance):
        keep = np.zeros(len(self.lx), np.int32)
        kept = lib().orc_delete_lipid_mask(self.n_cells, _p(self.centroids), _p(self.cs_l), _p(self.lx), C.c_float(stray_tolerance), _p(keep))
        if kept < len(self.lx):
            m = keep.astype(bool)
            self.lx, self.lv, self.ln, self.lo = (np.ascontiguousarray(a[m]) for a in (self.lx, self.lv, self.ln, self.lo))
            self.lf = np.zeros_like(self.lx); self.lt = np.zeros_like(self.lx)
            self.cell_update(0)  # cleanup.h:85
        return kept

    # openrbc.cpp:189-256, default build
    def step_langevin(self, freq_voronoi=2, noise=None):
        if self.nstep % freq_voronoi == 0:
            self.rebuild()
        self.compute_pairwise_fused()
        self.compute_bonded()
        nl, npr = noise if noise is not None else (None, None)
        self.verlet_langevin(nl, npr)
        self.nstep += 1

    def step_nh(self, freq_voronoi=2):
        self.nh_initial_fused()
        if self.nstep % freq_voronoi == 0:
            self.rebuild()
        self.compute_pairwise_fused()
        self.compute_bonded()
        self.nh_final_fused()
        self.nstep += 1


def frame_bytes(nstep, dump_field, tag_base, lx, lv, ln, lf, aff_l, px, pv, pn, pf, aff_p, ptype, ptag):
    """Restatement of save_frame (trajectory.h:61-105): FRAMEBEG nstep(int32) NATOM n(size_t) IDENTITY (tag, type)*
    [POSITION] [VELOCITY] [ROTATION] [VORONOI] [FORCE] FRAMEEND; titles NUL-padded to 8 bytes (:29-33); lipids first;
    lipid tag = base + i, type 0 (container.h:122-130); DumpField bits runtime_parameter.h:30-36."""
    def title(t):
        return t.encode().ljust(8, b"\0")
    nl, npr = len(lx), len(px)
    out = [title("FRAMEBEG"), np.int32(nstep).tobytes(), title("NATOM"), np.uint64(nl + npr).tobytes(), title("IDENTITY")]
    ident = np.empty((nl + npr, 2), np.int32)
    ident[:nl, 0] = tag_base + np.arange(nl); ident[:nl, 1] = 0
    ident[nl:, 0] = ptag; ident[nl:, 1] = ptype
    out.append(ident.tobytes())
    def sec(name, a, b, dt=np.float32):
        out.append(title(name)); out.append(np.ascontiguousarray(a, dt).tobytes()); out.append(np.ascontiguousarray(b, dt).tobytes())
    if dump_field & 1: sec("POSITION", lx, px)
    if dump_field & 8: sec("VELOCITY", lv, pv)
    if dump_field & 2: sec("ROTATION", ln, pn)
    if dump_field & 4: sec("VORONOI", aff_l, aff_p, np.int32)
    if dump_field & 16: sec("FORCE", lf, pf)
    out.append(title("FRAMEEND"))
    return b"".join(out)
