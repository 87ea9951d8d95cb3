"""Closed triangle meshes in the three-file text format OpenRBC's `-i trimesh -m <prefix>` reads
(`<prefix>.vert.txt`, `.bond.txt`, `.face.txt`; 0-based indices; reference: src/init_rbc.h:45-68,
fixture example-large/rbc.*.txt).  Host-side utility for tests and synthetic benchmark inputs;
not on the device path.

Face winding follows the reference's fixture: counter-clockwise seen from outside, so that the
reference's face normal cross(v0-v1, v2-v1) (init_rbc.h:92) points the same way as for rbc.*.txt.
"""
import numpy as np


def icosphere(subdiv):
    """Unit icosphere: V = 10*4^k + 2 vertices, E = 30*4^k edges, F = 20*4^k faces."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    verts = [np.array(p, float) / np.linalg.norm(p) for p in v]
    faces = [tuple(x) for x in f]
    for _ in range(subdiv):
        cache = {}

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        nf = []
        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = nf
    V = np.array(verts)
    F = np.array(faces, np.int64)
    # make every face counter-clockwise seen from outside
    n = np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]])
    flip = (n * V[F].mean(1)).sum(1) < 0
    F[flip] = F[flip][:, [0, 2, 1]]
    E = np.unique(np.sort(np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]]), axis=1), axis=0)
    return V, E, F


def biconcave(V, r0=3.91, c0=0.207, c1=2.003, c2=-1.123):
    """Map unit-sphere points onto the Evans-Fung biconcave RBC profile (radius r0 in the mesh's own
    units; OpenRBC rescales any mesh so that the mean edge is 80 nm, init_rbc.h:73-86)."""
    rho = np.sqrt(V[:, 0] ** 2 + V[:, 1] ** 2)
    s = np.clip(rho, 0.0, 1.0)
    z = 0.5 * np.sqrt(np.clip(1.0 - s * s, 0.0, None)) * (c0 + c1 * s ** 2 + c2 * s ** 4)
    out = np.empty_like(V)
    out[:, 0] = r0 * V[:, 0]
    out[:, 1] = r0 * V[:, 1]
    out[:, 2] = r0 * z * np.sign(V[:, 2])
    return out


def write_mesh(prefix, V, E, F):
    np.savetxt(prefix + ".vert.txt", V, fmt="%.6f", delimiter="\t")
    np.savetxt(prefix + ".bond.txt", E, fmt="%d", delimiter="\t")
    np.savetxt(prefix + ".face.txt", F, fmt="%d", delimiter="\t")
