"""ctypes binding of include/orbc_b200.h and the `Simulation` host mirror.

`Simulation` keeps the vocabulary of the reference (containers `lipid` / `protein`, Voronoi cells,
centroid stencils, functor-kernel names of integrate()) so that the parity tests read like the
reference's own main loop (src/openrbc.cpp:189-256).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liborbc_b200.so")

LIPID, PROTEIN = 0, 1
CLEAR_FORCE, POST_TORQUE, BOUNCE_BACK, VERLET_LANGEVIN, NH_INITIAL_FUSED, NH_FINAL_FUSED, NH_FINAL, NH_UPDATE = range(8)
OPT_MOVE, OPT_FUSED = 9, 10
STENCIL_STRIDE = 64
DUMP = dict(centroids=0, cell_start_l=1, cell_start_p=2, cells_l=3, cells_p=4, aff_l=5, aff_p=6, morton_keys=7, morton_perm=8,
            stencil_counts=9, stencil=10, tag2idx=11, counters=12, nl_stats=13)

PROF = dict(pair_lipid=0, pair_protein=1, bonded=2, integrate=3, rebuild=4)

EXPORTS = [
    "orbc_create", "orbc_destroy", "orbc_last_error", "orbc_synchronize", "orbc_set_stream", "orbc_forcefield_canonical",
    "orbc_set_forcefield", "orbc_upload", "orbc_upload_range", "orbc_upload_bonds", "orbc_voronoi_upload", "orbc_set_field", "orbc_voronoi_update",
    "orbc_cell_update", "orbc_rebuild", "orbc_delete_lipid", "orbc_compute_pairwise_fused", "orbc_compute_bonded",
    "orbc_constrain_volume", "orbc_integrate", "orbc_nh_zeta_update", "orbc_nh_zeta_update_unfused", "orbc_compute_temperature", "orbc_run_langevin",
    "orbc_run_nh", "orbc_download", "orbc_size", "orbc_n_cells", "orbc_debug_dump", "orbc_debug_noise", "orbc_event_record",
    "orbc_event_elapsed_ms", "orbc_launch_count", "orbc_profile_enable", "orbc_profile_read", "orbc_set_option",
    "orbc_voronoi_init", "orbc_set_volume_constraint", "orbc_run_minimize", "orbc_frame_bytes", "orbc_save_frame", "orbc_save_frame_begin", "orbc_save_frame_end",
    "orbc_profile_kernels", "orbc_profile_kernels_report", "orbc_mg_init", "orbc_mg_blob_bytes", "orbc_mg_cell_range", "orbc_mg_export", "orbc_mg_connect", "orbc_mg_range",
]


class OrbcError(RuntimeError):
    pass


class ForceField(C.Structure):
    _fields_ = [(n, C.c_float * k) for n, k in (
        ("mass", 6), ("radius", 6), ("cutlp", 6), ("cutsqlp", 6), ("replp", 6), ("attlp", 6), ("alphalp", 6),
        ("cutpp", 36), ("cutsqpp", 36), ("reppp", 36), ("lj_cutsq", 36), ("lj_lj1", 36), ("lj_lj2", 36),
        ("r0", 4), ("K", 4))] + [(n, C.c_float) for n in ("cutll", "cutsqll", "repll", "attll", "alphall")]

    def as_array(self):
        return np.frombuffer(bytes(self), np.float32).copy()


class StepParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("kBT", C.c_float), ("eta", C.c_float), ("zeta", C.c_float),
                ("box_lo", C.c_double), ("box_hi", C.c_double), ("dr_opt", C.c_double), ("dn_opt", C.c_double),
                ("nstep", C.c_int), ("seed", C.c_uint64), ("noise_lipid", C.c_void_p), ("noise_protein", C.c_void_p)]


class StepResult(C.Structure):
    _fields_ = [("ke", C.c_double), ("n", C.c_long)]


def library_path():
    return _SO


def build_library(verbose=False):
    """Compile csrc/ for sm_100a into the in-tree liborbc_b200.so (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc")], stdout=out)
    return _SO


_lib = None


def load_library():
    """Load liborbc_b200.so; fails loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise OrbcError(f"{_SO} is missing: build it with `make -C openrbc_b200/csrc` (python __graft_entry__.py build); "
                            "openrbc_b200 has no CPU fallback")
        lib = C.CDLL(_SO)
        lib.orbc_last_error.restype = C.c_char_p
        lib.orbc_nh_zeta_update.restype = C.c_float
        lib.orbc_nh_zeta_update.argtypes = [C.c_float, C.POINTER(C.c_float), C.c_double, C.c_float, C.c_double, C.c_long]
        lib.orbc_nh_zeta_update_unfused.restype = C.c_float
        lib.orbc_nh_zeta_update_unfused.argtypes = [C.c_float, C.POINTER(C.c_float), C.c_double, C.c_float, C.c_double, C.c_long]
        lib.orbc_constrain_volume.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p]
        lib.orbc_delete_lipid.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
        lib.orbc_debug_noise.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_size_t, C.c_void_p]
        lib.orbc_upload.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t] + [C.c_void_p] * 6
        lib.orbc_upload_bonds.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        lib.orbc_voronoi_upload.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orbc_set_field.argtypes = [C.c_void_p, C.c_int, C.c_char, C.c_size_t, C.c_void_p]
        lib.orbc_download.argtypes = [C.c_void_p, C.c_int, C.c_size_t] + [C.c_void_p] * 10
        lib.orbc_debug_dump.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        lib.orbc_integrate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.orbc_run_langevin.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.orbc_run_nh.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        for name in ("orbc_synchronize", "orbc_compute_pairwise_fused", "orbc_compute_bonded"):
            getattr(lib, name).argtypes = [C.c_void_p]
        lib.orbc_upload_range.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t] + [C.c_void_p] * 6
        lib.orbc_voronoi_update.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.orbc_cell_update.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.orbc_rebuild.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.orbc_compute_temperature.argtypes = [C.c_void_p, C.c_void_p]
        lib.orbc_event_record.argtypes = [C.c_void_p, C.c_int]
        lib.orbc_event_elapsed_ms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.orbc_launch_count.argtypes = [C.c_void_p, C.c_void_p]
        lib.orbc_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        lib.orbc_profile_enable.argtypes = [C.c_void_p, C.c_int]
        lib.orbc_profile_read.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.orbc_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        lib.orbc_destroy.argtypes = [C.c_void_p]
        lib.orbc_destroy.restype = None
        lib.orbc_profile_kernels.argtypes = [C.c_void_p, C.c_int]
        lib.orbc_profile_kernels_report.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.orbc_voronoi_init.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.orbc_set_volume_constraint.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        lib.orbc_run_minimize.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.orbc_frame_bytes.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.orbc_save_frame.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        lib.orbc_save_frame_begin.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.orbc_save_frame_end.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orbc_mg_init.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.orbc_mg_blob_bytes.restype = C.c_size_t
        lib.orbc_mg_cell_range.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.orbc_mg_export.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.orbc_mg_connect.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.orbc_mg_range.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def cell_range(n_cells, rank, world):
    """Cells [begin, end) owned by `rank` of `world` (util_numa.h:41-42)."""
    b, e = C.c_int(), C.c_int()
    lib = load_library()
    if lib.orbc_mg_cell_range(n_cells, rank, world, C.byref(b), C.byref(e)) != 0:
        raise OrbcError(lib.orbc_last_error().decode())
    return b.value, e.value


def forcefield_canonical():
    ff = ForceField()
    load_library().orbc_forcefield_canonical(C.byref(ff))
    return ff


def _f3(a):
    return np.ascontiguousarray(a, np.float32).reshape(-1, 3)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Simulation:
    """Device-resident copy of the reference's containers and Voronoi state, driven through the C ABI.

    `st` is the state dictionary produced by the reference's own initialisation (oracle.ref.Ref.state(),
    tools/make_states.py) or by the synthetic generators: lx lv ln lo, px pv pn po (N x 3 float32), ptype, ptag,
    bonds (B x 3: type, tag_i, tag_j), centroids (C x 3), cs_l, cs_p (C + 1).
    """

    def __init__(self, st, dt=1e-2, kBT=0.22, eta=0.01, seed=0xBAD5EED, device=0, box=(-1000.0, 1000.0), rank=0, world=1):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        self._ck(self.lib.orbc_create(C.byref(self.ctx), int(device)))
        self.rank, self.world = rank, world
        if world > 1:
            self._ck(self.lib.orbc_mg_init(self.ctx, rank, world))
        self.dt, self.kBT, self.eta, self.seed = dt, kBT, eta, seed
        self.box = box
        self.zeta, self.Q = 0.0, C.c_float(0.0)
        self.nstep = 0
        self.freq_sort_ctrd, self.freq_sort_bond, self.freq_voronoi = 24, 120, 2
        self.dr_opt = self.dn_opt = 5e-2
        self.last_ke = 0.0
        if st is not None:
            self.upload(st)

    # ---- plumbing ---------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise OrbcError(f"orbc error {rc}: {self.lib.orbc_last_error().decode()}")

    def close(self):
        if self.ctx:
            self.lib.orbc_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def params(self, noise_l=None, noise_p=None):
        p = StepParams()
        p.dt, p.kBT, p.eta, p.zeta = self.dt, self.kBT, self.eta, self.zeta
        p.box_lo, p.box_hi = self.box
        p.dr_opt, p.dn_opt = self.dr_opt, self.dn_opt
        p.nstep, p.seed = self.nstep, self.seed
        self._keep = (None if noise_l is None else _f3(noise_l), None if noise_p is None else _f3(noise_p))
        p.noise_lipid = None if self._keep[0] is None else self._keep[0].ctypes.data
        p.noise_protein = None if self._keep[1] is None else self._keep[1].ctypes.data
        return p

    # ---- upload / download ------------------------------------------------------------------------------
    def upload(self, st, owned_only=False):
        """Hand the host's containers to the device.  owned_only (a connected rank of a decomposed run re-uploading its system):
        only the rows of this rank's own cells travel, mg_export() then fetches the halo from the owners."""
        import time
        t = [time.perf_counter()]
        lx, lv, ln, lo = (_f3(st["l" + f]) for f in "xvno")
        px, pv, pn, po = (_f3(st["p" + f]) for f in "xvno")
        ty = np.ascontiguousarray(st["ptype"], np.int32)
        tg = np.ascontiguousarray(st["ptag"], np.int32)
        if owned_only and self.world > 1:
            cb, ce = cell_range(len(st["centroids"]), self.rank, self.world)
            bl, el = int(st["cs_l"][cb]), int(st["cs_l"][ce]); bp, ep = int(st["cs_p"][cb]), int(st["cs_p"][ce])
            self._ck(self.lib.orbc_upload_range(self.ctx, LIPID, len(lx), bl, el - bl, 3, _p(lx), _p(lv), _p(ln), _p(lo), None, None)); t.append(time.perf_counter())
            self._ck(self.lib.orbc_upload_range(self.ctx, PROTEIN, len(px), bp, ep - bp, 3, _p(px), _p(pv), _p(pn), _p(po), _p(ty), _p(tg))); t.append(time.perf_counter())
        else:
            self._ck(self.lib.orbc_upload(self.ctx, LIPID, len(lx), 3, _p(lx), _p(lv), _p(ln), _p(lo), None, None)); t.append(time.perf_counter())
            self._ck(self.lib.orbc_upload(self.ctx, PROTEIN, len(px), 3, _p(px), _p(pv), _p(pn), _p(po), _p(ty), _p(tg))); t.append(time.perf_counter())
        bd = np.ascontiguousarray(st["bonds"], np.int32).reshape(-1, 3)
        self._ck(self.lib.orbc_upload_bonds(self.ctx, len(bd), _p(bd))); t.append(time.perf_counter())
        if "centroids" in st:
            c = _f3(st["centroids"])
            csl = np.ascontiguousarray(st["cs_l"], np.int32)
            csp = np.ascontiguousarray(st["cs_p"], np.int32)
            self._ck(self.lib.orbc_voronoi_upload(self.ctx, len(c), _p(c), _p(csl), _p(csp))); t.append(time.perf_counter())
        # wall time of the four calls (lipids, proteins, bonds, Voronoi state), for the phase breakdown of bench.py
        self.last_upload_ms = [round((b - a) * 1e3, 2) for a, b in zip(t[:-1], t[1:])]
        rows_l, rows_p = (el - bl, ep - bp) if (owned_only and self.world > 1) else (len(lx), len(px))
        self.last_upload_bytes = 48 * (rows_l + rows_p) + ty.nbytes + tg.nbytes + bd.nbytes + sum(
            np.asarray(st[k]).nbytes for k in ("centroids", "cs_l", "cs_p") if k in st)

    def size(self, s):
        n = C.c_size_t()
        self._ck(self.lib.orbc_size(self.ctx, s, C.byref(n)))
        return n.value

    @property
    def n_cells(self):
        n = C.c_int()
        self._ck(self.lib.orbc_n_cells(self.ctx, C.byref(n)))
        return n.value

    def download(self, s, fields="xvnoft", ids=False, affiliation=False):
        n = self.size(s)
        out = {f: np.empty((n, 3), np.float32) for f in fields}
        ptr = [_p(out[f]) if f in out else None for f in "xvnoft"]
        aff = np.empty(n, np.int32) if affiliation else None
        ty = np.empty(n, np.int32) if ids else None
        tg = np.empty(n, np.int32) if ids else None
        nn = C.c_size_t()
        self._ck(self.lib.orbc_download(self.ctx, s, 3, *ptr, _p(aff), _p(ty), _p(tg), C.byref(nn)))
        if affiliation:
            out["affiliation"] = aff
        if ids:
            out["type"], out["tag"] = ty, tg
        return out

    def download_into(self, s, x=None, v=None, n=None, o=None, f=None, t=None, affiliation=False):
        """Download selected fields into caller-owned (e.g. pinned) N x 3 float32 arrays; returns the bytes copied."""
        cnt = self.size(s)
        aff = affiliation if isinstance(affiliation, np.ndarray) else (np.empty(cnt, np.int32) if affiliation else None)   # a caller-owned (pinned) int32 array, or True
        arrs = [x, v, n, o, f, t]
        for a in arrs:
            assert a is None or (a.dtype == np.float32 and a.flags.c_contiguous and len(a) >= cnt)
        nn = C.c_size_t()
        self._ck(self.lib.orbc_download(self.ctx, s, 3, *[_p(a) for a in arrs], _p(aff), None, None, C.byref(nn)))
        if self.world > 1:                 # a rank fills (and moves) only the rows it owns
            b, e = self.owned_range(s)
            cnt = e - b
        return sum(12 * cnt for a in arrs if a is not None) + (4 * cnt if aff is not None else 0)

    def get(self, s, field):
        return self.download(s, field)[field]

    def set_field(self, s, field, a):
        a = _f3(a)
        assert len(a) == self.size(s)
        self._ck(self.lib.orbc_set_field(self.ctx, s, field.encode(), 3, _p(a)))

    def dump(self, what):
        nc = self.n_cells
        shape, dt = {
            "centroids": ((nc, 3), np.float32), "cell_start_l": ((nc + 1,), np.int32), "cell_start_p": ((nc + 1,), np.int32),
            "cells_l": ((self.size(0),), np.int32), "cells_p": ((self.size(1),), np.int32),
            "aff_l": ((self.size(0),), np.int32), "aff_p": ((self.size(1),), np.int32),
            "morton_keys": ((nc,), np.uint32), "morton_perm": ((nc,), np.int32), "stencil_counts": ((nc, 3), np.int32),
            "stencil": ((nc, STENCIL_STRIDE), np.int32), "counters": ((8,), np.uint64), "nl_stats": ((4,), np.uint32),
        }[what]
        out = np.empty(shape, dt)
        self._ck(self.lib.orbc_debug_dump(self.ctx, DUMP[what], _p(out), out.nbytes))
        return out

    def stencils(self):
        """Per-cell centroid stencils as three lists of sorted arrays: r < 9, r < 8, r < 6."""
        cnt = self.dump("stencil_counts")
        st = self.dump("stencil")
        s9 = [np.sort(st[c, :cnt[c, 2]]) for c in range(len(cnt))]
        s8 = [np.sort(st[c, :cnt[c, 1]]) for c in range(len(cnt))]
        s6 = [np.sort(st[c, :cnt[c, 0]]) for c in range(len(cnt))]
        return s9, s8, s6

    def noise(self, nstep, species, n):
        out = np.empty((n, 3), np.float32)
        self._ck(self.lib.orbc_debug_noise(self.ctx, self.seed, nstep, species, n, _p(out)))
        return out

    def set_option(self, name, value):
        self._ck(self.lib.orbc_set_option(self.ctx, name.encode(), float(value)))

    def synchronize(self):
        self._ck(self.lib.orbc_synchronize(self.ctx))

    # ---- one cell over several GPUs ----------------------------------------------------------------------------
    def mg_export(self):
        """This rank's connection blob (bytes): raw pointers + CUDA IPC handles of the arrays its peers write into."""
        n = self.lib.orbc_mg_blob_bytes()
        buf = C.create_string_buffer(n)
        self._ck(self.lib.orbc_mg_export(self.ctx, buf, n))
        return buf.raw

    def mg_connect(self, blobs):
        """blobs: the mg_export() of every rank, in rank order."""
        assert len(blobs) == self.world
        n = self.lib.orbc_mg_blob_bytes()
        joined = b"".join(blobs)
        self._ck(self.lib.orbc_mg_connect(self.ctx, joined, n))

    def owned_range(self, s):
        b, e = C.c_size_t(), C.c_size_t()
        self._ck(self.lib.orbc_mg_range(self.ctx, s, C.byref(b), C.byref(e)))
        return b.value, e.value

    def voronoi_init(self, n_cells, n_iterate=64):               # voronoi.init(lipid, cell_lipid, param, 64)     openrbc.cpp:71
        self._ck(self.lib.orbc_voronoi_init(self.ctx, int(n_cells), int(n_iterate)))

    # ---- the reference's hot-path calls -------------------------------------------------------------------
    def voronoi_update(self):                                    # voronoi.update(lipid, cell_lipid, param)      openrbc.cpp:202
        self._ck(self.lib.orbc_voronoi_update(self.ctx, self.nstep, self.freq_sort_ctrd))

    def cell_update(self, s):                                    # cell_*.update(container, voronoi, param)      :203-204
        self._ck(self.lib.orbc_cell_update(self.ctx, s, self.nstep, self.freq_sort_bond))

    def rebuild(self):
        self._ck(self.lib.orbc_rebuild(self.ctx, self.nstep, self.freq_sort_ctrd, self.freq_sort_bond))

    def compute_pairwise_fused(self):                            # :219
        self._ck(self.lib.orbc_compute_pairwise_fused(self.ctx))

    def compute_bonded(self):                                    # :225
        self._ck(self.lib.orbc_compute_bonded(self.ctx))

    def integrate(self, kernel, noise_l=None, noise_p=None, want_result=False):
        p = self.params(noise_l, noise_p)
        res = StepResult()
        self._ck(self.lib.orbc_integrate(self.ctx, kernel, C.byref(p), C.byref(res) if want_result else None))
        return res

    def clear_force(self):
        self.integrate(CLEAR_FORCE)

    def post_torque(self):
        self.integrate(POST_TORQUE)

    def bounce_back(self):
        self.integrate(BOUNCE_BACK)

    def verlet_langevin(self, noise_l=None, noise_p=None):      # :233
        self.integrate(VERLET_LANGEVIN, noise_l, noise_p)

    def _zeta(self, res):
        self.last_ke = res.ke
        self.zeta = self.lib.orbc_nh_zeta_update(self.zeta, C.byref(self.Q), self.dt, self.kBT, res.ke, res.n)

    def nh_initial_fused(self):                                  # :193, destructor updates zeta
        res = self.integrate(NH_INITIAL_FUSED, want_result=True)
        self._zeta(res)
        return res.ke

    def nh_final_fused(self):                                    # :236
        res = self.integrate(NH_FINAL_FUSED, want_result=True)
        self._zeta(res)
        return res.ke

    def nh_final(self):                                          # verlet_nh_final (unfused), integrate_nh.h:155-176
        self.integrate(NH_FINAL)

    def nh_update(self):                                         # verlet_nh_update + its destructor, integrate_nh.h:66-94
        res = self.integrate(NH_UPDATE, want_result=True)
        self.last_ke = res.ke
        self.zeta = self.lib.orbc_nh_zeta_update_unfused(self.zeta, C.byref(self.Q), self.dt, self.kBT, res.ke, res.n)
        return res.ke

    def opt_move(self):                                          # :114-131
        self.integrate(OPT_MOVE)

    def opt_fused(self):                                         # :110-133 in one pass (post_torque + mover + bounce_back)
        self.integrate(OPT_FUSED)

    def compute_temperature(self):                               # :253
        t = C.c_double()
        self._ck(self.lib.orbc_compute_temperature(self.ctx, C.byref(t)))
        return t.value

    def constrain_volume(self, target, strength):                # :229
        v = C.c_float()
        self._ck(self.lib.orbc_constrain_volume(self.ctx, target, strength, C.byref(v)))
        return v.value

    def set_volume_constraint(self, on, target=3.15, strength=0.05):
        """Apply constrain_volume inside run_langevin / run_nh at the place of openrbc.cpp:229."""
        self._ck(self.lib.orbc_set_volume_constraint(self.ctx, int(on), target, strength))

    # ---- save_frame (trajectory.h:61-105) ---------------------------------------------------------------------
    def frame_bytes(self, dump_field=7):
        n = C.c_size_t()
        self._ck(self.lib.orbc_frame_bytes(self.ctx, dump_field, C.byref(n)))
        return n.value

    def save_frame(self, dump_field=7, tag_base=1, out=None):
        """One frame in the .orbc byte layout (np.uint8 array); `out` may be a caller-owned (pinned) buffer."""
        n = self.frame_bytes(dump_field)
        buf = out if out is not None else np.empty(n, np.uint8)
        got = C.c_size_t()
        self._ck(self.lib.orbc_save_frame(self.ctx, self.nstep, dump_field, tag_base, _p(buf), buf.nbytes, C.byref(got)))
        return buf[:got.value]

    def save_frame_begin(self, dump_field=7, tag_base=1):
        self._ck(self.lib.orbc_save_frame_begin(self.ctx, self.nstep, dump_field, tag_base))

    def save_frame_end(self):
        """View (np.uint8) of the oldest frame in flight, inside the library's pinned buffer: valid until the next-but-one begin."""
        ptr, n = C.c_void_p(), C.c_size_t()
        self._ck(self.lib.orbc_save_frame_end(self.ctx, C.byref(ptr), C.byref(n)))
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_ubyte)), shape=(n.value,))

    def delete_lipid(self, stray_tolerance):                     # :201
        n = C.c_size_t()
        self._ck(self.lib.orbc_delete_lipid(self.ctx, stray_tolerance, C.byref(n)))
        return n.value

    # ---- whole-loop entry points ----------------------------------------------------------------------
    def run_langevin(self, n_steps):
        p = self.params()
        self._ck(self.lib.orbc_run_langevin(self.ctx, C.byref(p), n_steps, self.freq_voronoi, self.freq_sort_ctrd))
        self.nstep += n_steps

    def run_minimize(self, n_steps):                             # openrbc.cpp:88-133
        p = self.params()
        self._ck(self.lib.orbc_run_minimize(self.ctx, C.byref(p), n_steps, self.freq_sort_ctrd))

    def run_nh(self, n_steps):
        p = self.params()
        z = C.c_float(self.zeta)
        self._ck(self.lib.orbc_run_nh(self.ctx, C.byref(p), n_steps, self.freq_voronoi, self.freq_sort_ctrd, C.byref(z), C.byref(self.Q)))
        self.zeta = z.value
        self.nstep += n_steps

    def step_langevin(self, noise=None):
        """One iteration of the reference's while-loop (openrbc.cpp:189-244), call for call."""
        if self.nstep % self.freq_voronoi == 0:
            self.rebuild()
        self.compute_pairwise_fused()
        self.compute_bonded()
        nl, npr = noise if noise is not None else (None, None)
        self.verlet_langevin(nl, npr)
        self.nstep += 1

    def step_langevin_cv_checked(self, target, strength):
        """The same iteration with constrain_volume at the place of openrbc.cpp:229 (BASELINE.json configs[2])."""
        if self.nstep % self.freq_voronoi == 0:
            self.rebuild()
        self.compute_pairwise_fused()
        self.compute_bonded()
        self.constrain_volume(target, strength)
        self.verlet_langevin()
        self.nstep += 1
        self.synchronize()

    def step_nh_checked(self):
        """One iteration of the Nose-Hoover branch of the loop (openrbc.cpp:193-240), call for call, then the status read-back."""
        self.nh_initial_fused()                      # :193 (its destructor updates zeta from the kinetic energy)
        if self.nstep % self.freq_voronoi == 0:
            self.rebuild()
        self.compute_pairwise_fused()
        self.compute_bonded()
        self.nh_final_fused()                        # :236
        self.nstep += 1
        self.synchronize()

    def step_langevin_checked(self):
        """One loop iteration call by call, then wait for it and read the device status back (16 B D2H)."""
        self.step_langevin()
        self.synchronize()

    # ---- timing ------------------------------------------------------------------------------------------
    def event_record(self, slot):
        self._ck(self.lib.orbc_event_record(self.ctx, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_float()
        self._ck(self.lib.orbc_event_elapsed_ms(self.ctx, a, b, C.byref(ms)))
        return ms.value

    def profile_enable(self, on=True):
        self._ck(self.lib.orbc_profile_enable(self.ctx, int(on)))

    def profile_read(self, cls):
        """(total device ms, launches) of one kernel class since the last read; cls is a key of PROF."""
        ms, cnt = C.c_double(), C.c_ulonglong()
        self._ck(self.lib.orbc_profile_read(self.ctx, PROF[cls], C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    def profile_kernels(self, on=True):
        self._ck(self.lib.orbc_profile_kernels(self.ctx, int(on)))

    def kernel_report(self):
        """[(kernel, launches, total_us)] since profile_kernels(True) / the last report, sorted by total time."""
        buf = C.create_string_buffer(1 << 16)
        self._ck(self.lib.orbc_profile_kernels_report(self.ctx, buf, len(buf)))
        rows = [ln.rsplit(None, 2) for ln in buf.value.decode().splitlines()]
        return [(r[0], int(r[1]), float(r[2])) for r in rows]

    def launch_count(self):
        n = C.c_ulonglong()
        self._ck(self.lib.orbc_launch_count(self.ctx, C.byref(n)))
        return n.value
