"""Shared helpers for the tests (mesh cases, seeds)."""
import numpy as np

from mesh_navigation_b200 import synth


def mesh_case(n, terrain, seed=42):
    pos, faces = synth.grid_mesh(n, n, terrain=terrain, seed=seed)
    return pos, faces


def face_of_vertex(faces, v):
    return int(np.where((faces == v).any(1))[0][0])


def centre_seed(pos, faces, frac=(0.5, 0.5)):
    n_extent = pos[:, :2].max(0)
    v = synth.nearest_vertex(pos, [n_extent[0] * frac[0], n_extent[1] * frac[1], float(pos[:, 2].mean())])
    f = face_of_vertex(faces, v)
    return v, f, pos[faces[f]].mean(0).astype(np.float32)


def rel_err(got, ref):
    fin = np.isfinite(ref)
    assert (np.isfinite(got) == fin).all(), "reached sets differ"
    r = np.zeros_like(ref, dtype=np.float64)
    r[fin] = np.abs(got[fin].astype(np.float64) - ref[fin]) / np.maximum(ref[fin], 1e-30)
    return r


def disc_lethals(pos, n_discs, radius, seed=7):
    """Synthetic obstacle lethals: vertices inside n discs at PCG32(seed) positions."""
    V = pos.shape[0]
    centres = synth.pcg32_stream(seed, n_discs, V)
    mask = np.zeros(V, dtype=bool)
    xy = pos[:, :2]
    for c in centres:
        d = xy - xy[c]
        lo = np.abs(d).max(1) <= radius
        idx = np.where(lo)[0]
        mask[idx[(d[idx] ** 2).sum(1) <= radius * radius]] = True
    return np.where(mask)[0].astype(np.uint32)


def delaunay_mesh(n_points, seed=3, extent=6.0, with_hub=True):
    """Irregular test mesh: Delaunay triangulation of random points (vertex degrees 3..12+), optionally with a
    hub vertex connected to a ring of 24 (degree 24 > the 8 ELL slots and > the 12-entry replay buffer)."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    xy = rng.random((n_points, 2)) * extent
    if with_hub:
        c = np.array([extent * 0.37, extent * 0.61])
        keep = np.linalg.norm(xy - c, axis=1) > 0.30
        ang = np.linspace(0, 2 * np.pi, 24, endpoint=False)
        ring = c + 0.22 * np.stack([np.cos(ang), np.sin(ang)], 1) * (1 + 0.05 * rng.random(24))[:, None]
        xy = np.concatenate([xy[keep], ring, c[None, :]])
    tri = Delaunay(xy)
    faces = tri.simplices.astype(np.uint32)
    # consistent CCW orientation
    a, b, c3 = xy[faces[:, 0]], xy[faces[:, 1]], xy[faces[:, 2]]
    area = (b[:, 0] - a[:, 0]) * (c3[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c3[:, 0] - a[:, 0])
    flip = area < 0
    faces[flip] = faces[flip][:, [0, 2, 1]]
    faces = faces[np.abs(area) > 1e-9]
    z = 0.4 * np.sin(xy[:, 0] * 1.3) * np.cos(xy[:, 1] * 0.9)
    pos = np.concatenate([xy, z[:, None]], 1).astype(np.float32)
    return pos, faces
