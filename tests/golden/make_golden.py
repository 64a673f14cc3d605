"""Generate the golden fixtures by running the UNMODIFIED PyRayT reference.

Runs only in the build container (needs /root/reference); the GPU box has no
reference, so the outputs are committed:

    tests/golden/<case>.npz       rays (13,N), frame (15,rows) from the reference, generation_limit
    tests/golden/<case>.scene.json  the flattened scene (pyrayt_b200.scene.FlatScene.to_json)

The reference is run with ``np.argsort`` defaulting to kind="stable" (the
behaviour of its pinned numpy 1.20.2) -- see oracle/ref_shim.py.

    python tests/golden/make_golden.py            # regenerate everything
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim, sources_np  # noqa: E402
from pyrayt_b200 import sources as dev_sources  # noqa: E402
from pyrayt_b200 import workloads  # noqa: E402
from pyrayt_b200.scene import flatten  # noqa: E402

pyrayt = ref_shim.load()
import pyrayt.components as pc  # noqa: E402
import pyrayt.materials as matl  # noqa: E402
import tinygfx.g3d as cg  # noqa: E402


from oracle import ref_scenes  # noqa: E402  (scene builders + the reference's trace() on a fixed RaySet)

reference_trace = ref_scenes.reference_trace


def cone(n, half_deg, apex, seed, wavelength=0.633):
    return sources_np.from_source(dev_sources.solid_angle_cone(seed, apex, half_deg, wavelength), n)


# ---------------------------------------------------------------- BASELINE.json configs

def config1_collimator():
    """examples/convex_collimator.py as shipped."""
    lens = pc.biconvex_lens(2, 2, 0.25, aperture=1)
    p = (1.5 - 1) * (1 / 2 - 1 / -2 + (1.5 - 1) * 0.25 / (1.5 * 2 * -2))
    src = pc.ConeOfRays(cone_angle=6).move_x(-1 / p)
    baffle = pc.baffle((1, 1)).move_x(1)
    return [lens, baffle], np.array(src.generate_rays(50)), 100


config2_scene = ref_scenes.config2_scene


def config2_tutorial():
    """tutorial condenser lens + aperture stop + detector, filled 10 degree cone, seed 1."""
    return config2_scene(), sources_np.from_source(workloads.CONFIG2_SOURCE, 4096), 100


config3_scene = ref_scenes.config3_scene


def config3_rays(per_source):
    srcs = [pc.LineOfRays(spacing=0.1, wavelength=x).move_x(-0.5).rotate_y(-3) for x in np.linspace(0.44, 0.75, 11)]
    rays = np.hstack([np.array(s.generate_rays(per_source)) for s in srcs])
    rays[12] = np.arange(rays.shape[1])
    return rays


def config3_prism():
    """examples/chromatic_dispersion.py geometry, 11 wavelengths."""
    return config3_scene(), config3_rays(96), 10


config4_scene = ref_scenes.config4_scene


CONFIG4_SOURCE = workloads.CONFIG4_SOURCE


def config4_stack():
    return config4_scene(), sources_np.from_source(CONFIG4_SOURCE, 1536), 64


config5_scene = ref_scenes.config5_scene


CONFIG5_SOURCE = workloads.CONFIG5_SOURCE


def config5_cavity():
    return config5_scene(), sources_np.from_source(CONFIG5_SOURCE, 1024), 32


# ---------------------------------------------------------------- extra coverage

def facing_mirrors():
    """test_core.py:54-66 style: two facing mirrors bounce until the generation limit."""
    m1 = pc.plane_mirror(0.2, aperture=(2, 2)).move_x(2)
    m2 = pc.plane_mirror(0.2, aperture=2.0).move_x(-2)
    return [m1, m2], cone(256, 5.0, (0.0, 0.0, 0.0), 3), 10


def curved_mirrors():
    pm = pc.parabolic_mirror(5, 1, aperture=4)
    sm = pc.spherical_mirror(10, 1, aperture=3).move_x(-8).rotate_z(180)
    rays = cone(768, 30.0, (0.0, 0.0, 0.0), 4)
    rays[4:7] *= -1
    return [pm, sm, pc.baffle((6, 6)).move_x(12)], rays, 20


def thick_lens_zoo():
    """Every thick_lens flavour of int_test_thick_lenses.py plus a rectangular aperture."""
    comps = [
        pc.thick_lens(np.inf, np.inf, 0.5, aperture=2.0).move_x(0),
        pc.thick_lens(5, 8, 0.4, aperture=2.0).move_x(2),          # meniscus
        pc.thick_lens(6, -6, 0.6, aperture=2.0).move_x(4),         # biconvex
        pc.thick_lens(6, np.inf, 0.5, aperture=2.0).move_x(6),     # plano-convex
        pc.thick_lens(-6, 6, 0.3, aperture=2.0).move_x(8),         # biconcave
        pc.thick_lens(np.inf, 6, 0.3, aperture=(2.0, 1.5)).move_x(10),  # plano-concave, rectangular
        pc.plano_convex_lens(4, 0.5, aperture=2.0).move_x(12),
        pc.baffle((4, 4)).move_x(14),
    ]
    rays = cone(1024, 8.0, (-4.0, 0.0, 0.0), 7)
    return comps, rays, 40


def nested_csg():
    """Right-nested tree, UNION (incl. the disjoint-union bounding-box quirk Q4), scaled/rotated leaves."""
    glass = matl.glass["ideal"]
    a = cg.Sphere(1.0, material=glass).move_x(0.0)
    b = cg.Sphere(1.0, material=glass).move_x(0.8)
    c = cg.Cylinder(0.7, -2, 2, material=glass).rotate_y(90)
    right_nested = cg.csg.union(a, cg.csg.intersect(b, c))
    d = cg.Sphere(0.5, material=matl.mirror).move(4.0, 0.0, 0.0)
    e = cg.Sphere(0.5, material=matl.mirror).move(4.0, 2.0, 0.0)  # disjoint: invisible (Q4)
    disjoint = cg.csg.union(d, e)
    f = cg.Cuboid.from_sides(1.0, 1.0, 1.0, material=glass).scale(1.0, 2.0, 0.5).rotate_z(25).rotate_x(10).move(-3, 0, 0)
    g = cg.Sphere(0.6, material=glass).scale(1.0, 1.5, 1.0).move(-3.2, 0.3, 0.0)
    diff = cg.csg.difference(f, g)
    h = cg.csg.difference(cg.csg.intersect(cg.Sphere(1.2, material=glass), cg.Cuboid.from_sides(2, 2, 2, material=glass)),
                          cg.csg.union(cg.Cylinder(0.3, -3, 3, material=glass), cg.Cylinder(0.3, -3, 3, material=glass).rotate_x(90)))
    h.move(0, 4, 0)
    walls = [pc.baffle((20, 20)).move_x(9), pc.baffle((20, 20)).move_x(-9),
             pc.baffle((20, 20)).rotate_z(90).move_y(9), pc.baffle((20, 20)).rotate_z(90).move_y(-9)]
    rng = np.random.default_rng(11)
    n = 1536
    rays = np.zeros((13, n))
    rays[0:3] = rng.uniform(-6, 6, (3, n)) * np.array([[1.0], [1.0], [0.15]])
    v = rng.normal(size=(3, n)) * np.array([[1.0], [1.0], [0.2]])
    rays[4:7] = v / np.linalg.norm(v, axis=0)
    rays[3] = 1
    rays[9] = 100
    rays[10] = rng.uniform(0.45, 0.7, n)
    rays[11] = 1
    rays[12] = np.arange(n)
    return [right_nested, disjoint, diff, h] + walls, rays, 12


def stop_ties():
    """Wide-angle rays on a bare aperture: Plane's [t,t] pair against the cylinder hits (Q3)."""
    stop = pc.aperture((1, 1), 0.5).move_x(0.5)
    det = pc.baffle((3, 3)).move_x(1)
    return [stop, det], cone(2048, 80.0, (0.0, 0.0, 0.0), 9), 10


def axis_aligned_edge_cases():
    """Exactly axis-parallel, near-axis (Q2) and grid rays on lens + cube mirror + plane."""
    lens = pc.biconvex_lens(2, 2, 0.25, aperture=1)
    cube = pc.plane_mirror(0.2, aperture=(1.0, 1.0)).move(3, 0, 0)
    det = pc.baffle((2, 2)).move_x(-3)
    n_side = 15
    ys, zs = np.meshgrid(np.linspace(-0.6, 0.6, n_side), np.linspace(-0.6, 0.6, n_side))
    n = ys.size
    blocks = []
    for tilt in (0.0, 1e-9, 1e-5, 5e-5, 2e-4, 1e-3):
        r = np.zeros((13, n))
        r[0] = -2
        r[1] = ys.ravel()
        r[2] = zs.ravel()
        r[3] = 1
        r[4] = np.cos(tilt)
        r[5] = np.sin(tilt)
        r[9] = 100
        r[10] = 0.55
        r[11] = 1
        blocks.append(r)
    rays = np.hstack(blocks)
    rays[12] = np.arange(rays.shape[1])
    return [lens, cube, det], rays, 12


CASES = {
    "config1_collimator": config1_collimator,
    "config2_tutorial": config2_tutorial,
    "config3_prism": config3_prism,
    "config4_stack": config4_stack,
    "config5_cavity": config5_cavity,
    "facing_mirrors": facing_mirrors,
    "curved_mirrors": curved_mirrors,
    "thick_lens_zoo": thick_lens_zoo,
    "nested_csg": nested_csg,
    "stop_ties": stop_ties,
    "axis_aligned_edge_cases": axis_aligned_edge_cases,
}


def main(argv):
    names = argv or list(CASES)
    for name in names:
        comps, rays, gl = CASES[name]()
        scene = flatten(comps)
        frame = reference_trace(rays, comps, gl)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rays=rays, frame=frame,
                            generation_limit=np.int64(gl))
        with open(os.path.join(HERE, name + ".scene.json"), "w") as fh:
            fh.write(scene.to_json())
        if name.startswith("config"):  # the benchmark workloads also ship with the package
            with open(os.path.join(ROOT, "pyrayt_b200", "data", name + ".scene.json"), "w") as fh:
                fh.write(scene.to_json())
        per_ray = frame.shape[1] / max(1, rays.shape[1])
        print(f"{name:28s} rays {rays.shape[1]:6d} rows {frame.shape[1]:7d} ({per_ray:5.2f}/ray) "
              f"leaves {scene.n_leaves:3d} nodes {scene.n_nodes:3d} max gen {int(frame[0].max()) if frame.size else -1}")


if __name__ == "__main__":
    main(sys.argv[1:])
