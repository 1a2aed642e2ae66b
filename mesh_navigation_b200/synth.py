"""Deterministic synthetic triangle meshes for the parity tests and the bench.

The reference ships no mesh (it loads HDF5/assimp files through lvr2,
mesh_map/src/mesh_map.cpp:149-452), so the workloads of BASELINE.json are the
synthetic meshes SURVEY.md section 8d defines:

* regular grid nx x ny, spacing h = 0.1 m, vertex (i, j) -> index j*nx + i,
  every quad split by the same diagonal (interior degree 6), CCW faces;
* deterministic xy jitter of +-0.2 h from splitmix64(seed ^ index) so that
  exact float ties have measure ~0;
* planar: z = 0;  terrain: z = 2.0 m * fBm(x / 20 m, y / 20 m), 4 octaves,
  lacunarity 2, gain 0.5, classic Perlin gradient noise with a permutation
  table shuffled by the seed.

This is bench/test input generation only; it is not on the hot path.
"""
from __future__ import annotations

import numpy as np

_MASK64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser on uint64 arrays."""
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def _u01(bits: np.ndarray) -> np.ndarray:
    """uint64 -> float64 in [0, 1)."""
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


class _Perlin:
    """Classic 2-D Perlin gradient noise (Ken Perlin 1985 formulation)."""

    def __init__(self, seed: int):
        rng_bits = splitmix64(np.arange(256, dtype=np.uint64) ^ np.uint64(seed))
        # Fisher-Yates with a deterministic bit stream
        perm = np.arange(256, dtype=np.int64)
        for i in range(255, 0, -1):
            j = int(rng_bits[i] % np.uint64(i + 1))
            perm[i], perm[j] = perm[j], perm[i]
        self.perm = np.concatenate([perm, perm])
        ang = 2.0 * np.pi * np.arange(256) / 256.0
        self.gx = np.cos(ang)
        self.gy = np.sin(ang)

    @staticmethod
    def _fade(t):
        return t * t * t * (t * (t * 6.0 - 15.0) + 10.0)

    def noise(self, x: np.ndarray, y: np.ndarray) -> np.ndarray:
        xi = np.floor(x).astype(np.int64)
        yi = np.floor(y).astype(np.int64)
        xf = x - xi
        yf = y - yi
        xi &= 255
        yi &= 255
        p = self.perm
        h00 = p[p[xi] + yi]
        h10 = p[p[xi + 1] + yi]
        h01 = p[p[xi] + yi + 1]
        h11 = p[p[xi + 1] + yi + 1]
        n00 = self.gx[h00] * xf + self.gy[h00] * yf
        n10 = self.gx[h10] * (xf - 1) + self.gy[h10] * yf
        n01 = self.gx[h01] * xf + self.gy[h01] * (yf - 1)
        n11 = self.gx[h11] * (xf - 1) + self.gy[h11] * (yf - 1)
        u = self._fade(xf)
        v = self._fade(yf)
        nx0 = n00 + u * (n10 - n00)
        nx1 = n01 + u * (n11 - n01)
        return nx0 + v * (nx1 - nx0)


def fbm(x: np.ndarray, y: np.ndarray, seed: int, octaves: int = 4,
        lacunarity: float = 2.0, gain: float = 0.5) -> np.ndarray:
    pn = _Perlin(seed)
    amp = 1.0
    freq = 1.0
    out = np.zeros_like(x, dtype=np.float64)
    for _ in range(octaves):
        out += amp * pn.noise(x * freq, y * freq)
        amp *= gain
        freq *= lacunarity
    return out


def grid_mesh(nx: int, ny: int, *, h: float = 0.1, jitter: float = 0.2,
              seed: int = 42, terrain: bool = False, z_scale: float = 2.0,
              wavelength: float = 20.0):
    """Return (positions float32 [V,3], faces uint32 [F,3]).

    Row chunks keep the peak memory at 50 M vertices reasonable.
    """
    V = nx * ny
    idx = np.arange(V, dtype=np.uint64)
    i = (idx % np.uint64(nx)).astype(np.float64)
    j = (idx // np.uint64(nx)).astype(np.float64)
    if jitter > 0.0:
        bx = splitmix64(idx ^ np.uint64(seed))
        by = splitmix64(bx)
        jx = (_u01(bx) * 2.0 - 1.0) * jitter * h
        jy = (_u01(by) * 2.0 - 1.0) * jitter * h
        del bx, by
    else:
        jx = jy = 0.0
    x = i * h + jx
    y = j * h + jy
    del i, j, jx, jy, idx
    pos = np.empty((V, 3), dtype=np.float32)
    pos[:, 0] = x
    pos[:, 1] = y
    if terrain:
        chunk = 1 << 22
        for s in range(0, V, chunk):
            e = min(V, s + chunk)
            pos[s:e, 2] = z_scale * fbm(x[s:e] / wavelength, y[s:e] / wavelength, seed)
    else:
        pos[:, 2] = 0.0
    del x, y

    qi = np.arange(nx - 1, dtype=np.uint32)
    qj = np.arange(ny - 1, dtype=np.uint32)
    v00 = (qj[:, None] * np.uint32(nx) + qi[None, :]).reshape(-1)
    faces = np.empty((v00.size, 2, 3), dtype=np.uint32)
    faces[:, 0, 0] = v00
    faces[:, 0, 1] = v00 + 1
    faces[:, 0, 2] = v00 + nx + 1
    faces[:, 1, 0] = v00
    faces[:, 1, 1] = v00 + nx + 1
    faces[:, 1, 2] = v00 + nx
    return pos, faces.reshape(-1, 3)


def pcg32_stream(seed: int, n: int, bound: int) -> np.ndarray:
    """n bounded draws from PCG32 (XSH-RR), sequence 54 (O'Neill's demo stream)."""
    mask = (1 << 64) - 1
    mult = 6364136223846793005
    inc = (54 << 1) | 1
    state = 0
    state = (state * mult + inc) & mask
    state = (state + seed) & mask
    state = (state * mult + inc) & mask
    out = np.empty(n, dtype=np.int64)
    for k in range(n):
        old = state
        state = (old * mult + inc) & mask
        xorshifted = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        r = ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & 0xFFFFFFFF
        out[k] = r % bound
    return out


def batch_goal_vertices(V: int, n: int, seed: int = 1234, blocked=None) -> np.ndarray:
    """First n distinct vertex ids from PCG32(seed) that are not blocked."""
    goals = []
    seen = set()
    draws = pcg32_stream(seed, 4 * n + 64, V)
    for v in draws:
        v = int(v)
        if v in seen:
            continue
        if blocked is not None and blocked[v]:
            continue
        seen.add(v)
        goals.append(v)
        if len(goals) == n:
            break
    if len(goals) < n:
        raise RuntimeError("not enough goal draws")
    return np.asarray(goals, dtype=np.uint32)


def nearest_vertex(pos: np.ndarray, p) -> int:
    d = pos.astype(np.float64) - np.asarray(p, dtype=np.float64)[None, :]
    return int(np.argmin(np.einsum("ij,ij->i", d, d)))


def disc_lethals_grid(pos: np.ndarray, nx: int, ny: int, n_discs: int, radius: float, h: float = 0.1, seed: int = 7) -> np.ndarray:
    """Synthetic obstacle lethals of SURVEY 8d config 3 on a grid_mesh: vertices within `radius` (xy) of n_discs
    centre vertices drawn from PCG32(seed).  Same set as the exhaustive scan (tests/util.disc_lethals), but only the
    index window around each centre is examined."""
    V = nx * ny
    centres = pcg32_stream(seed, n_discs, V)
    k = int(np.ceil(radius / h)) + 2                      # jitter is +-0.2 h per vertex
    di, dj = np.meshgrid(np.arange(-k, k + 1), np.arange(-k, k + 1), indexing="xy")
    ci, cj = centres % nx, centres // nx
    ii = ci[:, None] + di.ravel()[None, :]
    jj = cj[:, None] + dj.ravel()[None, :]
    ok = (ii >= 0) & (ii < nx) & (jj >= 0) & (jj < ny)
    idx = np.where(ok, jj * nx + ii, 0)
    d = pos[idx, :2] - pos[centres, None, :2]
    inside = ok & (np.abs(d).max(2) <= radius) & ((d ** 2).sum(2) <= radius * radius)
    return np.unique(idx[inside]).astype(np.uint32)
