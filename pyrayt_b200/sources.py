"""Seeded synthetic sources generated on the device (K3).

The benchmark configs need up to 2^28 rays; building them with NumPy on the host
and copying 104 B/ray over PCIe would dominate the run, so the rays are written
straight into a device RaySet by ``prt_generate_source``.  The law is a
counter-based integer hash followed only by + - * / sqrt, which makes it
restatable bit-for-bit on the CPU (``oracle/sources_np.py`` is that restatement
and the tests compare the two).

u(i, k) = mix64(seed ^ (i*0x9E3779B97F4A7C15 + k*0xD1B54A32D192ED03)) >> 11, scaled by 2^-53
mix64   = splitmix64 finaliser.

kind 1  collimated field fan (config 4): start points uniform on a disk of radius R in
        the plane x = origin.x (rejection-sampled from the square); ray i belongs to
        field i % 3 (direction (cos a_f, sin a_f, 0)) and wavelength (i // 3) % 3.
kind 2  point source, directions uniform in solid angle within a cone about +x (config 2).
kind 3  point source, Lambertian within a cone about -x (config 5; Malley's method).
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass, field
from typing import Sequence

from . import _lib


@dataclass
class SyntheticSource:
    kind: int
    seed: int
    origin: Sequence[float]
    p: Sequence[float] = field(default_factory=lambda: [0.0] * 16)

    def desc(self) -> _lib.PrtSourceDesc:
        d = _lib.PrtSourceDesc()
        d.kind = self.kind
        d.seed = self.seed
        for k in range(3):
            d.origin[k] = float(self.origin[k])
        pp = list(self.p) + [0.0] * (16 - len(self.p))
        for k in range(16):
            d.p[k] = float(pp[k])
        return d

    def generate(self, n: int, device: int = 0, first_index: int = 0, out=None):
        """Returns a (13, n) float64 CUDA tensor in the reference RaySet layout."""
        import torch

        lib = _lib.load()
        if out is None:
            out = torch.empty((_lib.RAY_ROWS, n), dtype=torch.float64, device=torch.device("cuda", device))
        d = self.desc()
        with torch.cuda.device(device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
            _lib.check(lib.prt_generate_source(ctypes.byref(d), out.data_ptr(), n, int(out.stride(0)) if n else 0,
                                               first_index, stream), "prt_generate_source")
        return out


def field_fan(seed: int, x_start: float, radius: float, field_deg=(0.0, 2.0, 5.0),
              wavelengths=(0.486, 0.588, 0.656), intensity: float = 100.0) -> SyntheticSource:
    p = [radius, 0.0]
    for a in field_deg:
        p += [math.cos(math.radians(a)), math.sin(math.radians(a))]
    p += list(wavelengths) + [intensity]
    return SyntheticSource(1, seed, (x_start, 0.0, 0.0), p)


def solid_angle_cone(seed: int, apex, half_angle_deg: float, wavelength: float = 0.633,
                     intensity: float = 100.0) -> SyntheticSource:
    return SyntheticSource(2, seed, apex, [math.cos(math.radians(half_angle_deg)), wavelength, intensity])


def lambertian_cone(seed: int, apex, half_angle_deg: float, wavelength: float = 0.633,
                    intensity: float = 100.0) -> SyntheticSource:
    return SyntheticSource(3, seed, apex, [math.sin(math.radians(half_angle_deg)), wavelength, intensity])
