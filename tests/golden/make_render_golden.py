"""Golden fixtures for the renderers' nearest-hit loop, from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  For every case the reference's
``EdgeRender`` is stepped through INITIALIZE and PROPAGATE (tinygfx/g3d/renderers.py:62-94) and the
camera rays, hit distances and hit surfaces are stored, plus the finished edge canvas:

    tests/golden/render_<case>.npz          rays (2,4,N), distance (N,), surface (N,), canvas (v,h,4), resolution
    tests/golden/render_<case>.scene.json   the flattened scene

    python tests/golden/make_render_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from pyrayt_b200.scene import flatten  # noqa: E402

pyrayt = ref_shim.load()
import pyrayt.components as pc  # noqa: E402
import pyrayt.materials as matl  # noqa: E402
import tinygfx.g3d as cg  # noqa: E402
from tinygfx.g3d import renderers  # noqa: E402
from tinygfx.g3d.world_objects import OrthographicCamera  # noqa: E402


def system():
    lens = pc.thick_lens(20, -20, 4, aperture=12, material=matl.glass["BK7"])
    neg = pc.thick_lens(-25, 25, 2, aperture=12, material=matl.glass["SF5"]).move_x(8)
    stop = pc.aperture((14, 14), 6.0).move_x(14)
    mirror = pc.plane_mirror(2, aperture=(8, 8)).rotate_z(30).move(20, 2, 1)
    det = pc.baffle((14, 14)).move_x(26)
    return [lens, neg, stop, mirror, det]


def top_view():
    """The camera of renderers._draw_xy (renderers.py:285-292): above the system, looking down -z."""
    comps = system()
    box = np.hstack([c.bounding_volume.bounding_points[:3] for c in comps])
    mins, maxes = box.min(axis=1), box.max(axis=1)
    origin = (maxes + mins) / 2
    origin[2] = 1.5 * maxes[2]
    h_span, v_span = 1.5 * (maxes[:2] - mins[:2])
    cam = OrthographicCamera(120, h_span, v_span / h_span)
    cam.rotate_y(90).rotate_z(90).move(*origin[:3])
    return comps, cam


def inside_view():
    """A camera in the middle of the system looking along +x: half of the components are behind it,
    so pixels whose hits are all negative take the renderers' unfiltered-slot-0 branch."""
    comps = system()
    cam = OrthographicCamera(96, 18.0, 0.75)
    cam.move(11.0, 0.3, -0.2)
    return comps, cam


def oblique_view():
    comps = system()
    cam = OrthographicCamera(80, 40.0, 0.6)
    cam.rotate_y(25).rotate_z(-35).move(-6.0, 9.0, 7.0)
    return comps, cam


CASES = {"top_view": top_view, "inside_view": inside_view, "oblique_view": oblique_view}


def main():
    for name, make in CASES.items():
        comps, cam = make()
        r = renderers.EdgeRender(cam, comps)
        with ref_shim.stable_argsort(), np.errstate(all="ignore"):
            r._st_initialize()
            r._st_propagate()
            rays = np.array(r._rays, dtype=np.float64)
            dist, surf = np.array(r._hit_distances, dtype=np.float64), np.array(r._hit_surfaces, dtype=np.int64)
            r._st_interact()
            canvas = np.array(r._results)
        np.savez_compressed(os.path.join(HERE, f"render_{name}.npz"), rays=rays, distance=dist, surface=surf,
                            canvas=canvas, resolution=np.array(cam.get_resolution()))
        with open(os.path.join(HERE, f"render_{name}.scene.json"), "w") as fh:
            fh.write(flatten(comps).to_json())
        print(f"{name}: {rays.shape[-1]} pixels, {np.mean(surf >= 0):.2f} hit, "
              f"{int(np.sum(dist < 0))} negative distances, surfaces {sorted(set(surf.tolist()))[:8]}")


if __name__ == "__main__":
    main()
