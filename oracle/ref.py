"""TEST INFRASTRUCTURE — ctypes wrapper around oracle/_ref/libref_{strict,fast}.so.

The shared objects are the UNMODIFIED reference headers (/root/reference/src/*.h) compiled behind
oracle/ref_harness.cpp by oracle/Makefile.  They travel to the GPU box prebuilt (oracle/_ref is
git-ignored but not gpurun-ignored).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

F32P = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
I32P = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

# ids shared with include/orbc_b200.h (orbc_integrator)
CLEAR_FORCE, POST_TORQUE, BOUNCE_BACK, VERLET_LANGEVIN, NH_INITIAL_FUSED, NH_FINAL_FUSED, NH_FINAL, NH_UPDATE, ASSIGN_TEMPERATURE = range(9)


def available(kind="strict"):
    return os.path.exists(os.path.join(_HERE, "_ref", f"libref_{kind}.so"))


class Ref:
    """One reference world.  A process must keep a single thread count for its lifetime (the
    reference keeps function-local statics sized by omp_get_max_threads())."""

    def __init__(self, kind="strict", threads=1, args=()):
        path = os.path.join(_HERE, "_ref", f"libref_{kind}.so")
        self.lib = lib = C.CDLL(path)
        lib.ref_size.restype = C.c_long
        lib.ref_n_bonds.restype = C.c_long
        lib.ref_delete_lipid.restype = C.c_long
        lib.ref_compute_temperature.restype = C.c_double
        lib.ref_get_param.restype = C.c_double
        lib.ref_run_langevin.restype = C.c_double
        lib.ref_run_nh.restype = C.c_double
        lib.ref_run_opt.restype = C.c_double
        lib.ref_timer.restype = C.c_double
        lib.ref_uint2u11.restype = C.c_float
        lib.ref_morton_encode.restype = C.c_uint
        lib.ref_morton_encode.argtypes = [C.c_float] * 3
        lib.ref_set_param.argtypes = [C.c_char_p, C.c_double]
        lib.ref_get_param.argtypes = [C.c_char_p]
        lib.ref_constrain_volume.argtypes = [C.c_float, C.c_float]
        lib.ref_init_lipid_sphere.argtypes = [C.c_float]
        argv = [b"openrbc"] + [str(a).encode() for a in args]
        arr = (C.c_char_p * len(argv))(*argv)
        # the reference prints its parameter table to stdout; silence it
        self._quiet(lambda: lib.ref_create(len(argv), arr, int(threads)))

    @staticmethod
    def _quiet(fn):
        import sys
        sys.stdout.flush()
        saved = os.dup(1)
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)
        try:
            return fn()
        finally:
            os.dup2(saved, 1)
            os.close(devnull)
            os.close(saved)

    # ---- init -------------------------------------------------------------------------------
    def init_lipid_sphere(self, r):
        self._quiet(lambda: self.lib.ref_init_lipid_sphere(float(r)))

    def init_trimesh(self):
        self._quiet(lambda: self.lib.ref_init_trimesh())

    def voronoi_init(self, n_iter=64):
        return self._quiet(lambda: self.lib.ref_voronoi_init(int(n_iter)))

    # ---- params -----------------------------------------------------------------------------
    def set_param(self, name, v):
        assert self.lib.ref_set_param(name.encode(), float(v)) == 0, name

    def get_param(self, name):
        return self.lib.ref_get_param(name.encode())

    # ---- state ------------------------------------------------------------------------------
    def size(self, s):
        return self.lib.ref_size(s)

    @property
    def n_cells(self):
        return self.lib.ref_n_cells()

    def get(self, s, name):
        out = np.empty((self.size(s), 3), np.float32)
        self.lib.ref_get(s, name.encode(), out.ctypes.data_as(C.c_void_p))
        return out

    def set(self, s, name, a):
        a = np.ascontiguousarray(a, np.float32)
        assert a.shape == (self.size(s), 3)
        self.lib.ref_set(s, name.encode(), a.ctypes.data_as(C.c_void_p))

    def protein_ids(self):
        n = self.size(1)
        t = np.empty(n, np.int32)
        g = np.empty(n, np.int32)
        self.lib.ref_get_protein_ids(t.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p))
        return t, g

    def bonds(self):
        b = np.empty((self.lib.ref_n_bonds(), 3), np.int32)
        self.lib.ref_get_bonds(b.ctypes.data_as(C.c_void_p))
        return b

    def centroids(self):
        c = np.empty((self.n_cells, 3), np.float32)
        self.lib.ref_get_centroids(c.ctypes.data_as(C.c_void_p))
        return c

    def cell_array(self, s, what):
        idx = {"cell_start": 0, "cells": 1, "affiliation": 2, "local_index": 3}[what]
        n = self.n_cells + 1 if idx == 0 else self.size(s)
        a = np.empty(n, np.int32)
        self.lib.ref_get_cell_array(s, idx, a.ctypes.data_as(C.c_void_p))
        return a

    def lipid_tag_base(self):
        return self.lib.ref_lipid_tag_base()

    def state(self):
        """Snapshot of everything the hot path reads (numpy arrays)."""
        st = {}
        for s, p in ((0, "l"), (1, "p")):
            for f in "xvno":
                st[p + f] = self.get(s, f)
        st["ptype"], st["ptag"] = self.protein_ids()
        st["bonds"] = self.bonds()
        if self.n_cells:
            st["centroids"] = self.centroids()
            st["cs_l"] = self.cell_array(0, "cell_start")
            st["cs_p"] = self.cell_array(1, "cell_start")
        return st

    def load_state(self, st):
        z = lambda a: np.ascontiguousarray(a, np.float32)
        v = C.c_void_p
        lx, lv, ln, lo = (z(st["l" + f]) for f in "xvno")
        self.lib.ref_set_lipids(C.c_long(len(lx)), lx.ctypes.data_as(v), lv.ctypes.data_as(v), ln.ctypes.data_as(v), lo.ctypes.data_as(v))
        px, pv, pn, po = (z(st["p" + f]) for f in "xvno")
        ty = np.ascontiguousarray(st["ptype"], np.int32)
        tg = np.ascontiguousarray(st["ptag"], np.int32)
        bd = np.ascontiguousarray(st["bonds"], np.int32).reshape(-1, 3)
        self.lib.ref_set_proteins(C.c_long(len(px)), px.ctypes.data_as(v), pv.ctypes.data_as(v), pn.ctypes.data_as(v), po.ctypes.data_as(v),
                                  ty.ctypes.data_as(v), tg.ctypes.data_as(v), C.c_long(len(bd)), bd.ctypes.data_as(v))
        if "centroids" in st:
            c = z(st["centroids"])
            csl = np.ascontiguousarray(st["cs_l"], np.int32)
            csp = np.ascontiguousarray(st["cs_p"], np.int32)
            self.lib.ref_set_voronoi(len(c), c.ctypes.data_as(v), csl.ctypes.data_as(v), csp.ctypes.data_as(v))

    # ---- hot-path calls -----------------------------------------------------------------------
    def voronoi_update(self):
        self.lib.ref_voronoi_update()

    def cell_update(self, s):
        self.lib.ref_cell_update(s)

    def update_particle_affiliation(self, s):
        self.lib.ref_update_particle_affiliation(s)

    def compute_pairwise_fused(self):
        self.lib.ref_compute_pairwise_fused()

    def compute_bonded(self):
        self.lib.ref_compute_bonded()

    def compute_temperature(self):
        return self.lib.ref_compute_temperature()

    def constrain_volume(self, target, strength):
        self._quiet(lambda: self.lib.ref_constrain_volume(float(target), float(strength)))

    def delete_lipid(self):
        return self.lib.ref_delete_lipid()

    def integrate(self, kernel):
        assert self.lib.ref_integrate(int(kernel)) == 0

    def opt_move(self):
        self.lib.ref_opt_move()

    def save_frame(self, dump_field=7):
        """Bytes of one frame as the reference's save_frame writes it (trajectory.h:61-105)."""
        import tempfile
        fd, path = tempfile.mkstemp(suffix=".orbc")
        os.close(fd)
        try:
            assert self.lib.ref_save_frame(path.encode(), int(dump_field)) == 0
            with open(path, "rb") as f:
                return f.read()
        finally:
            os.unlink(path)

    def stencil(self, cell, rmax, cap=256):
        out = np.empty(cap, np.int32)
        n = self.lib.ref_get_stencil(int(cell), C.c_float(rmax), out.ctypes.data_as(C.c_void_p), cap)
        assert n <= cap
        return out[:n].copy()

    def stencil_refined(self, cell, cap=256):
        o9, o8, o6 = (np.empty(cap, np.int32) for _ in range(3))
        n = np.zeros(3, np.int32)
        v = C.c_void_p
        self.lib.ref_get_stencil_refined(int(cell), o9.ctypes.data_as(v), o8.ctypes.data_as(v), o6.ctypes.data_as(v), cap, n.ctypes.data_as(v))
        return o9[:n[0]].copy(), o8[:n[1]].copy(), o6[:n[2]].copy()

    def morton_encode(self, x, y, z):
        return self.lib.ref_morton_encode(float(x), float(y), float(z))

    def reorder_morton(self, pts):
        p = np.ascontiguousarray(pts, np.float32).copy()
        self.lib.ref_reorder_morton(C.c_long(len(p)), p.ctypes.data_as(C.c_void_p))
        return p

    def uint2u11(self, u):
        return self.lib.ref_uint2u11(C.c_uint(int(u)))

    def run_langevin(self, n, cleanup=False):
        return self.lib.ref_run_langevin(int(n), int(cleanup))

    def run_nh(self, n, cleanup=False):
        return self.lib.ref_run_nh(int(n), int(cleanup))

    def run_opt(self, n):
        return self.lib.ref_run_opt(int(n))

    def timer(self, name):
        return self.lib.ref_timer(name.encode())

    def forcefield(self):
        n = self.lib.ref_forcefield_floats()
        out = np.empty(n, np.float32)
        self.lib.ref_get_forcefield(out.ctypes.data_as(C.c_void_p))
        return out
