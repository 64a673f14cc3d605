"""NumPy restatement of the seeded synthetic sources (TEST INFRASTRUCTURE).

Bit-for-bit the law of ``source_kernel`` in pyrayt_b200/csrc/prt_kernels.cu /
pyrayt_b200/sources.py: uint64 hashing, then only + - * / sqrt in the same
order, so the device-generated RaySet can be checked for exact equality and the
CPU baseline can trace the very same rays.
"""
import numpy as np

_C1 = np.uint64(0x9E3779B97F4A7C15)
_C2 = np.uint64(0xD1B54A32D192ED03)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def u01(seed: int, i: np.ndarray, k: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = np.uint64(seed) ^ (i.astype(np.uint64) * _C1 + np.uint64(k) * _C2)
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def unit_disk(seed: int, i: np.ndarray, k0: int):
    a = np.zeros(i.shape)
    b = np.zeros(i.shape)
    todo = np.ones(i.shape, dtype=bool)
    for k in range(0, 64, 2):
        if not todo.any():
            break
        x = 2 * u01(seed, i, k0 + k) - 1
        y = 2 * u01(seed, i, k0 + k + 1) - 1
        r2 = x * x + y * y
        ok = todo & (r2 <= 1.0) & (r2 > 1e-12)
        a = np.where(ok, x, a)
        b = np.where(ok, y, b)
        todo &= ~ok
    return a, b


def generate(kind: int, seed: int, origin, p, n: int, first_index: int = 0) -> np.ndarray:
    """(13, n) float64 RaySet array (pyrayt/_pyrayt.py:13-44 layout)."""
    i = np.arange(first_index, first_index + n, dtype=np.uint64)
    p = list(p) + [0.0] * (16 - len(p))
    rays = np.zeros((13, n))
    rays[0], rays[1], rays[2] = origin
    rays[3] = 1.0
    rays[11] = 1.0
    rays[12] = i.astype(np.float64)
    if kind == 1:
        a, b = unit_disk(seed, i, 0)
        rays[1] = origin[1] + p[0] * a
        rays[2] = origin[2] + p[0] * b
        f = (i % np.uint64(3)).astype(np.int64)
        rays[4] = np.asarray([p[2], p[4], p[6]])[f]
        rays[5] = np.asarray([p[3], p[5], p[7]])[f]
        w = ((i // np.uint64(3)) % np.uint64(3)).astype(np.int64)
        rays[10] = np.asarray(p[8:11])[w]
        rays[9] = p[11]
    elif kind == 2:
        ct = 1 - u01(seed, i, 0) * (1 - p[0])
        st = np.sqrt(1 - ct * ct)
        a, b = unit_disk(seed, i, 1)
        r = np.sqrt(a * a + b * b)
        rays[4] = ct
        rays[5] = st * (a / r)
        rays[6] = st * (b / r)
        rays[10] = p[1]
        rays[9] = p[2]
    elif kind == 3:
        a, b = unit_disk(seed, i, 0)
        u = p[0] * a
        v = p[0] * b
        rays[4] = -np.sqrt(1 - (u * u + v * v))
        rays[5] = u
        rays[6] = v
        rays[10] = p[1]
        rays[9] = p[2]
    else:
        raise ValueError(kind)
    return rays


def from_source(src, n: int, first_index: int = 0, total=None) -> np.ndarray:
    """Same rays as ``pyrayt_b200.sources.SyntheticSource.generate`` /
    ``ReferenceSourceSet.generate`` (host array)."""
    if hasattr(src, "windows"):  # ReferenceSourceSet: concatenated reference sources, ids renumbered
        out = np.empty((13, n))
        for s, first, count, off in src.windows(n, first_index, total):
            full = reference_source(s.kind, list(s.p), int(s.p[2]))
            out[:, off:off + count] = full[:, first:first + count]
        return out
    return generate(src.kind, src.seed, tuple(src.origin), list(src.p), n, first_index)


def reference_source(kind: int, p, n: int, seed: int = 0, origin=(0.0, 0.0, 0.0)) -> np.ndarray:
    """NumPy restatement of pyrayt.components Line/Circle/Cone/WedgeOfRays.generate_rays
    (pyrayt/components.py:481-613) for the descriptor layout of pyrayt_b200.sources.from_reference:
    the same NumPy expressions as the reference, so it is bit-identical to it."""
    rays = np.zeros((2, 4, n))
    rays[0, 3] = 1
    if kind == 10:
        if n > 1:
            rays[0, 1] = np.linspace(-p[0] / 2, p[0] / 2, n)
        rays[1, 0] = 1
    elif kind == 11:
        theta = np.linspace(0, 2 * np.pi, n)
        rays[0, 1] = p[0] / 2 * np.sin(theta)
        rays[0, 2] = p[0] / 2 * np.cos(theta)
        rays[1, 0] = 1
    elif kind == 12:
        if n > 1:
            angles = 2 * np.pi * np.arange(0, n) / n
            rays[1, 1] = np.sin(p[0]) * np.sin(angles)
            rays[1, 2] = np.sin(p[0]) * np.cos(angles)
        rays[1, 0] = np.cos(p[0])
    elif kind == 13:
        angles = np.linspace(-p[0] / 2, p[0] / 2, n)
        rays[1, 0] = np.cos(angles)
        rays[1, 1] = np.sin(angles)
    elif kind == 14:
        # Lamp._local_ray_generation (pyrayt/components.py:637-654) with u01(seed, ray id, k) as the uniforms
        rid = (np.float64(p[3]) + np.arange(n)).astype(np.uint64)
        theta = np.arccos(1 - u01(seed, rid, 0) * (1 - np.cos(p[0])))
        phi = u01(seed, rid, 1) * (2 * 3.141592653589793)
        rays[0, 1] = origin[0] * (u01(seed, rid, 2) - 0.5)
        rays[0, 2] = origin[1] * (u01(seed, rid, 3) - 0.5)
        rays[1, 0] = np.cos(theta)
        rays[1, 1] = np.sin(theta) * np.cos(phi)
        rays[1, 2] = np.sin(theta) * np.sin(phi)
        intensity = 100.0 * np.cos(theta)
    else:
        raise ValueError(kind)
    M = np.eye(4)
    M[:3] = np.asarray(p[4:16]).reshape(3, 4)
    rays = np.matmul(M, rays)
    rays[1] /= np.linalg.norm(rays[1], axis=0)
    out = np.zeros((13, n))
    out[:8] = rays.reshape(8, n)
    out[9] = intensity if kind == 14 else 100.0
    out[10] = p[1]
    out[11] = 1.0
    out[12] = p[3] + np.arange(n)
    return out
