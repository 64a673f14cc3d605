"""Pin the C oracle to the reference's own known-answer tests (restated, not copied).

Each test names the reference test it restates (paths under /root/reference/test).
"""
import numpy as np
import pytest

from oracle import oracle
from tests import scene_util as su

INF = np.inf
S2 = np.sqrt(2.0)


def hits(ptype, params, origin, direction):
    """primitive.intersect for one ray, sorted like TracerSurface.intersect does."""
    return np.sort(oracle.prim_intersect(ptype, params, origin, direction))


# ---------------------------------------------------------------- test_primitives.py

class TestSphere:  # test_primitives.py TestSphere :95-185
    def test_unit_sphere(self):  # :126-137
        assert np.array_equal(hits(su.SPHERE, [1], (0, 0, 0), (1, 0, 0)), [-1, 1])
        assert np.all(np.isinf(hits(su.SPHERE, [1], (0, 0, 2), (1, 0, 0))))

    def test_sphere_behind_ray(self):  # :139-145
        assert np.array_equal(hits(su.SPHERE, [1], (100, 0, 0), (1, 0, 0)), [-101, -99])

    def test_double_root(self):  # :160-163 tangent ray
        np.testing.assert_allclose(hits(su.SPHERE, [1], (-1, 0, 1), (1, 0, 0)), [1, 1])

    def test_normals(self):  # :165-171
        for p in ((0, 0, -1), (0, 0, 1), (0, 1, 0), (0, -1, 0), (1, 0, 0), (-1, 0, 0)):
            np.testing.assert_allclose(oracle.prim_normal(su.SPHERE, [1], p), p)

    def test_arrayed_normals(self):  # :173-182
        rng = np.random.default_rng(0)
        pts = rng.normal(size=(200, 3))
        pts /= np.linalg.norm(pts, axis=1, keepdims=True)
        for p in pts:
            np.testing.assert_allclose(oracle.prim_normal(su.SPHERE, [1], p), p, atol=1e-15)


class TestParaboloid:  # test_primitives.py TestParaboloid :185-283, f=1, height=3
    P = [1, 3]

    def test_at_origin(self):  # :194-202
        np.testing.assert_allclose(hits(su.PARABOLOID, self.P, (0, 0, 0), (0, 1, 0)), [0, 0], atol=1e-15)
        np.testing.assert_allclose(hits(su.PARABOLOID, self.P, (0, 0, 0), (0, 0, 1)), [0, 3])
        np.testing.assert_allclose(hits(su.PARABOLOID, self.P, (0, 0, 0), (1, 0, 0)), [0, 0], atol=1e-15)

    def test_linear_case(self):  # :204-210
        np.testing.assert_allclose(hits(su.PARABOLOID, self.P, (0, 0, -1), (0, 0, 1)), [1, 4])
        np.testing.assert_allclose(hits(su.PARABOLOID, self.P, (2, 0, 0), (0, 0, -1)), [-3, -1])

    def test_double_root(self):  # :212-218
        np.testing.assert_allclose(hits(su.PARABOLOID, self.P, (0, -2, 1), (0, 1, 0)), [0, 4])
        np.testing.assert_allclose(hits(su.PARABOLOID, self.P, (0, 0, 1), (0, 1, 0)), [-2, 2])

    def test_skew(self):  # :220-229
        assert np.all(np.isinf(hits(su.PARABOLOID, self.P, (-1, 0, 0), (0, 1, 0))))
        assert np.all(np.isinf(hits(su.PARABOLOID, self.P, (-10, 0, 0), (0, 0, 1))))
        assert np.all(np.isinf(hits(su.PARABOLOID, self.P, (0, 0, 1.05 * 3), (1, 1, 0))))

    def test_arrayed(self):  # :231-246
        np.testing.assert_allclose(hits(su.PARABOLOID, self.P, (0, 0, 1), (0, 1, 0)), [-2, 2])
        np.testing.assert_allclose(hits(su.PARABOLOID, self.P, (0, 0, 1), (0, 0, 1)), [-1, 2])

    def test_normals(self):  # :266-283
        np.testing.assert_allclose(oracle.prim_normal(su.PARABOLOID, self.P, (0, 0, 0)), (0, 0, -1))
        np.testing.assert_allclose(oracle.prim_normal(su.PARABOLOID, self.P, (0, 2, 1)), np.array((0, 1, -1)) / S2)
        np.testing.assert_allclose(oracle.prim_normal(su.PARABOLOID, self.P, (0.3, 0.2, 3.0)), (0, 0, 1))  # cap


class TestPlane:  # test_primitives.py TestPlane :286-343
    def test_positive(self):  # :290-297
        np.testing.assert_allclose(hits(su.PLANE, [2, 2], (0, 0, 1), (0, 0, -1)), [1, 1])
        np.testing.assert_allclose(hits(su.PLANE, [2, 2], (0, 0, 1), np.array((-1, 0, -1)) / S2), [S2, S2])

    def test_negative(self):  # :299-302
        np.testing.assert_allclose(hits(su.PLANE, [2, 2], (0, 0, 1), (0, 0, 1)), [-1, -1])

    def test_parallel_and_missed(self):  # :304-313
        assert np.all(np.isinf(hits(su.PLANE, [2, 2], (0, 0, 1), (1, 0, 0))))
        assert np.all(np.isinf(hits(su.PLANE, [2, 2], (2, 0, 1), (0, 0, -1))))

    def test_patch_bounds(self):  # :315-327 width 3, length 2
        np.testing.assert_allclose(hits(su.PLANE, [3, 2], (1.49, 0.99, 1), (0, 0, -1)), [1, 1])
        assert np.all(np.isinf(hits(su.PLANE, [3, 2], (1.51, 1.01, 1), (0, 0, -1))))

    def test_arrayed(self):  # :329-343
        np.testing.assert_allclose(hits(su.PLANE, [2, 2], (0, 0, -1), (0, 1, 1)), [1, 1])
        assert np.all(np.isinf(hits(su.PLANE, [2, 2], (0, 0, -1), (0, 1, 0))))

    def test_returns_same_t_twice(self):  # primitives.py:492
        h = oracle.prim_intersect(su.PLANE, [2, 2], (0.1, 0.2, 1), (0, 0.1, -1))
        assert h[0] == h[1]


UNIT_CUBE = [-1, 1, -1, 1, -1, 1]


class TestCube:  # test_primitives.py TestCube :346-494
    def test_within(self):  # :366-374
        for d in np.eye(3):
            assert np.array_equal(hits(su.CUBE, UNIT_CUBE, (0, 0, 0), d), [-1, 1])

    def test_external(self):  # :376-385
        for k in range(3):
            o = np.zeros(3)
            o[k] = -2
            assert np.array_equal(hits(su.CUBE, UNIT_CUBE, o, np.eye(3)[k]), [1, 3])

    def test_at_angle(self):  # :386-390 hits sqrt(2)*[1,2]
        np.testing.assert_allclose(hits(su.CUBE, UNIT_CUBE, (-2, -1, 0), np.array((1, 1, 0)) / S2), S2 * np.array([1, 2]))

    def test_skew(self):  # :392-395
        assert np.all(np.isinf(hits(su.CUBE, UNIT_CUBE, (-2, 0, 0), (0, 1, 0))))

    def test_nondefault(self):  # :397-409
        spans = [-1, 1, -1, 2, -1, 5]
        for k, ext in enumerate((1, 2, 5)):
            assert np.array_equal(hits(su.CUBE, spans, (0, 0, 0), np.eye(3)[k]), [-1, ext])

    def test_arrayed(self):  # :411-430
        assert np.array_equal(hits(su.CUBE, UNIT_CUBE, (-0.5, 0, 0), (1, 0, 0)), [-0.5, 1.5])
        assert np.all(np.isinf(hits(su.CUBE, UNIT_CUBE, (-2, 0, 0), (0, -1, 0))))

    def test_normals(self):  # :432-444
        for p in ((-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)):
            np.testing.assert_allclose(oracle.prim_normal(su.CUBE, UNIT_CUBE, p), p)

    def test_normals_non_unit(self):  # :446-461
        spans = [-2, 5, -3, 6, -4, 7]
        for p, n in (((-2, 0, 0.5), (-1, 0, 0)), ((5, 1, 0.5), (1, 0, 0)), ((0.5, -3, 0.5), (0, -1, 0)),
                     ((0.5, 6, 1), (0, 1, 0)), ((0.5, 1, -4), (0, 0, -1)), ((0.5, 1, 7), (0, 0, 1))):
            np.testing.assert_allclose(oracle.prim_normal(su.CUBE, spans, p), n)

    def test_offcenter_and_corner(self):  # :463-477
        np.testing.assert_allclose(oracle.prim_normal(su.CUBE, UNIT_CUBE, (-1 + 1e-8, 0.3, 0.7)), (-1, 0, 0))
        np.testing.assert_allclose(oracle.prim_normal(su.CUBE, UNIT_CUBE, (1, 1, 1)), np.ones(3) / np.sqrt(3))

    def test_off_face_is_nan(self):  # SURVEY 9-Q8 (primitives.py:593-599)
        assert np.all(np.isnan(oracle.prim_normal(su.CUBE, UNIT_CUBE, (0.2, 0.3, 0.4))))


class TestCylinder:  # test_primitives.py TestCylinder :497-628, radius 1, z in [-1, 1]
    C = [1, -1, 1, 1]

    def test_sidewalls(self):  # :501-510
        for z in (0, 0.5, -0.5):
            np.testing.assert_allclose(hits(su.CYLINDER, self.C, (-2, 0, z), (1, 0, 0)), [1, 3])

    def test_cap(self):  # :512-519
        np.testing.assert_allclose(hits(su.CYLINDER, self.C, (0, 0, 0), (0, 0, 1)), [-1, 1])

    def test_wall_cap(self):  # :521-528 (direction is not normalised in the reference test)
        np.testing.assert_allclose(hits(su.CYLINDER, self.C, (-2, 0, -1), (1, 0, 1)), [1, 2])

    def test_nondefault(self):  # :530-543
        c = [2, 0, 4, 1]
        assert np.all(np.isinf(hits(su.CYLINDER, c, (0, 0, -1), (1, 0, 0))))
        np.testing.assert_allclose(hits(su.CYLINDER, c, (0, 0, 3), (1, 0, 0)), [-2, 2])
        np.testing.assert_allclose(hits(su.CYLINDER, c, (0, 0, 0), (0, 0, 1)), [0, 4])

    def test_no_intersection_outside(self):  # :545-557
        assert np.all(np.isinf(hits(su.CYLINDER, self.C, (2, 0, 0), (0, 0, 1))))

    def test_normals(self):  # :577-605
        for p in ((-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)):
            np.testing.assert_allclose(oracle.prim_normal(su.CYLINDER, self.C, p), p)
        c = [3, -5, 7, 1]
        for p, n in (((-3, 0, 0), (-1, 0, 0)), ((0, 3, 0), (0, 1, 0)), ((0, 0, -5), (0, 0, -1)), ((0, 0, 7), (0, 0, 1))):
            np.testing.assert_allclose(oracle.prim_normal(su.CYLINDER, c, p), n)

    def test_near_axis_rays_miss(self):  # |d_xy|^2 <= 1e-8 takes the linear branch
        # at 1e-5 rad from the axis the "linear" branch yields a single root -c/b (far away), so the
        # cylinder is missed; exactly on the axis and at 2e-4 rad it is hit normally
        for ang, hit in ((0.0, True), (1e-5, False), (2e-4, True)):
            d = np.array((np.sin(ang), 0, np.cos(ang)))
            got = hits(su.CYLINDER, self.C, (0.2, 0, -3), d)
            assert np.all(np.isfinite(got)) == hit, (ang, got)


# ---------------------------------------------------------------- test_operations.py

class TestReflectRefract:
    V = np.array((1, 1, 0)) / S2
    N = np.array((-1.0, 0, 0))

    def test_reflect(self):  # test_operations.py TestReflections :180-214
        np.testing.assert_allclose(oracle.reflect((1, -1, 0), (0, 1, 0)), (1, 1, 0))
        np.testing.assert_allclose(oracle.reflect((0, 0, 1), (0, 0, 1)), (0, 0, -1))

    def test_into_higher(self):  # :221-232
        out, idx = oracle.refract(self.V, self.N, 1, 1.5)
        th = np.arcsin(1 * S2 / (2 * 1.5))
        assert idx == 1.5
        np.testing.assert_allclose(out, (np.cos(th), np.sin(th), 0), atol=1e-15)

    def test_into_lower(self):  # :234-245
        out, idx = oracle.refract(self.V, self.N, 1.1, 1.0)
        th = np.arcsin(1.1 * S2 / 2)
        assert idx == 1.0
        np.testing.assert_allclose(out, (np.cos(th), np.sin(th), 0), atol=1e-15)

    def test_into_world(self):  # :247-261
        out, idx = oracle.refract(self.V, -self.N, 1.5, 1.5, 1.4)
        th = np.arcsin(1.5 * S2 / (2 * 1.4))
        assert idx == 1.4
        np.testing.assert_allclose(out, (np.cos(th), np.sin(th), 0), atol=1e-15)

    def test_total_internal_reflection(self):  # :263-285
        out, idx = oracle.refract(self.V, -self.N, 1.5, 1.5, 1.0)
        assert idx == 1.5
        np.testing.assert_allclose(out, np.array((-1, 1, 0)) / S2, atol=1e-15)
        out, idx = oracle.refract(self.V, self.N, 1.5, 1.0, 1.0)
        assert idx == 1.5
        np.testing.assert_allclose(out, np.array((-1, 1, 0)) / S2, atol=1e-15)

    def test_arrayed_split(self):  # :287-318
        out, idx = oracle.refract(self.V, self.N, 1.5, 1.0)
        assert idx == 1.5
        out, idx = oracle.refract(self.V, self.N, 1.5, 1.6)
        th = np.arcsin(1.5 * S2 / (2 * 1.6))
        assert idx == 1.6
        np.testing.assert_allclose(out, (np.cos(th), np.sin(th), 0), atol=1e-15)


def test_binomial_root_cases_through_cylinder():
    """test_operations.py TestBinomialRoot :101-177 exercised through Cylinder (its only caller)."""
    tall = [1, -1e9, 1e9, 1]
    # two real roots
    np.testing.assert_allclose(hits(su.CYLINDER, tall, (-3, 0, 0), (1, 0, 0)), [2, 4])
    # no real roots -> inf
    assert np.all(np.isinf(hits(su.CYLINDER, tall, (-3, 2, 0), (1, 0, 0))))
    # a ~ 0 (ray along the axis) inside the radius: side = (-inf, +inf), caps decide
    np.testing.assert_allclose(hits(su.CYLINDER, [1, -1, 1, 1], (0.5, 0, -3), (0, 0, 1)), [2, 4])
    # a ~ 0 outside the radius: miss
    assert np.all(np.isinf(hits(su.CYLINDER, [1, -1, 1, 1], (1.5, 0, -3), (0, 0, 1))))


# ---------------------------------------------------------------- test_csg.py

A1 = [1, 4, 5, 10]
A2 = [0, 2, 3, 5, 6, 7, 8, 9, 11, 12]


@pytest.mark.parametrize("op,expected", [
    (1, (0, 10, 11, 12)),                      # test_csg.py:217-220 union
    (2, (1, 2, 3, 4, 5, 5, 6, 7, 8, 9)),       # :222-225 intersect
    (3, (2, 3, 5, 6, 7, 8, 9, 10)),            # :227-231 difference
])
def test_array_csg_golden_vectors(op, expected):
    out = oracle.array_csg(A1, A2, op)
    want = np.full(14, INF)
    want[: len(expected)] = expected
    assert np.array_equal(out, want)


def _two_spheres(op):
    left = su.Leaf(su.SPHERE, [1], sid=7)
    right = su.Leaf(su.SPHERE, [1], world=su.translate(0, -1, 0), sid=9)
    box = {1: (-1, 1, -2, 1, -1, 1), 2: (-1, 1, -1, 0, -1, 1), 3: (-1, 1, -1, 1, -1, 1)}[op]
    return su.build([su.Node(op, left, right, box)])


def _x_rays(ys):
    n = len(ys)
    rays = np.zeros((2, 4, n))
    rays[0, 0] = -5
    rays[0, 1] = ys
    rays[0, 3] = 1
    rays[1, 0] = 1
    return rays


def _sphere_hits(yc, ys):
    disc = 1 - (np.asarray(ys) - yc) ** 2
    with np.errstate(invalid="ignore"):
        r = np.sqrt(disc)
    return np.where(disc >= 0, 5 - r, INF), np.where(disc >= 0, 5 + r, INF)


def test_csg_union_two_spheres():  # test_csg.py:57-92
    ys = np.linspace(-2, 2, 11)
    h, s = oracle.intersect(_two_spheres(1), 0, _x_rays(ys))
    assert np.all(np.isinf(h[2:]))
    missed = np.all(np.isinf(h), axis=0)
    assert not np.any(missed[(ys > -2) & (ys < 1)])
    assert np.array_equal(h, np.sort(h, axis=0))
    rn, rf = _sphere_hits(-1, ys)
    ln, lf = _sphere_hits(0, ys)
    sel = ys < -0.5
    np.testing.assert_allclose(h[:2, sel], np.vstack((rn, rf))[:, sel])
    assert np.all(s[:2, sel & ~missed] == 9)
    sel = ys > -0.5
    np.testing.assert_allclose(h[:2, sel], np.vstack((ln, lf))[:, sel])
    assert np.all(s[:2, sel & ~missed] == 7)


def test_csg_intersect_two_spheres():  # test_csg.py:116-150
    ys = np.linspace(-2, 2, 11)
    h, s = oracle.intersect(_two_spheres(2), 0, _x_rays(ys))
    assert np.all(np.isinf(h[2:]))
    missed = np.all(np.isinf(h), axis=0)
    assert not np.any(missed[(ys > -1) & (ys < 0)])
    rn, rf = _sphere_hits(-1, ys)
    ln, lf = _sphere_hits(0, ys)
    sel = (ys < -0.5) & ~missed
    np.testing.assert_allclose(h[:2, sel], np.vstack((ln, lf))[:, sel])
    assert np.all(s[:2, sel] == 7)
    sel = (ys > -0.5) & ~missed
    np.testing.assert_allclose(h[:2, sel], np.vstack((rn, rf))[:, sel])
    assert np.all(s[:2, sel] == 9)


def test_csg_difference_two_spheres():  # test_csg.py:174-209
    ys = np.linspace(-2, 2, 101)
    h, s = oracle.intersect(_two_spheres(3), 0, _x_rays(ys))
    assert np.all(np.isinf(h[2:, ys > 0]))
    mid = (ys < 0) & (ys > -0.5)
    assert not np.any(np.isinf(h[2:, mid]))
    missed = np.all(np.isinf(h), axis=0)
    assert np.all(missed[(ys < -0.5) | (ys > 1)])
    ln, lf = _sphere_hits(0, ys)
    rn, rf = _sphere_hits(-1, ys)
    sel = ys > 0
    np.testing.assert_allclose(h[:2, sel], np.vstack((ln, lf))[:, sel])
    assert np.all(s[:2, sel & ~missed] == 7)
    np.testing.assert_allclose(h[[0, 3]][:, mid], np.vstack((ln, lf))[:, mid])
    assert np.all(s[[0, 3]][:, mid] == 7)
    np.testing.assert_allclose(h[1:3, mid], np.vstack((rn, rf))[:, mid])
    assert np.all(s[1:3, mid] == 9)


# ---------------------------------------------------------------- test_world_objects.py / materials

def test_moved_sphere_hits_at_one():  # test_world_objects.py:277-282
    scene = su.build([su.Leaf(su.SPHERE, [1], world=su.translate(2, 0, 0))])
    rays = _x_rays([0.0])
    rays[0, 0] = 0.0
    h, s = oracle.intersect(scene, 0, rays)
    assert h.shape == (2, 1) and s.shape == (2, 1)  # :263-275 shape / ids
    np.testing.assert_allclose(h[:, 0], [1, 3])
    assert np.all(s == 100)


def test_sellmeier_index():  # test_pyrayt_materials.py:114-134
    for coeff in ([1, 0, 0, 1, 0, 0], [0, 1, 0, 0, 1, 0], [0, 0, 1, 0, 0, 1]):
        assert oracle.index_at(3, coeff, 2.0) == pytest.approx(np.sqrt(7 / 3), abs=1e-15)
    assert oracle.index_at(2, [1.6], 0.5) == 1.6


def _plane_material_trace(mat, matp, direction, index=1.0, wavelength=0.633):
    """material.trace on an XYPlane at the origin, via a one-generation trace (rows hold the result)."""
    scene = su.build([su.Leaf(su.PLANE, [2, 2], mat=mat, matp=matp),
                      su.Leaf(su.SPHERE, [50], mat=su.MAT_ABSORBER)])
    d = np.asarray(direction, dtype=np.float64)
    o = -d / np.linalg.norm(d)
    rays = su.make_rays([o], [d], wavelength=wavelength, index=index)
    frame, _ = oracle.trace(scene, rays, 5)
    return frame


def test_absorber_ends_the_ray():  # test_pyrayt_materials.py:15-21
    f = _plane_material_trace(su.MAT_ABSORBER, [], (0, 0, -1))
    assert f.shape[1] == 1 and f[5, 0] == 100


def test_mirror_reflection():  # test_pyrayt_materials.py:29-46
    f = _plane_material_trace(su.MAT_MIRROR, [], (0, 1, -1))
    assert f.shape[1] == 2
    np.testing.assert_allclose(f[12:15, 1], np.array((0, 1, 1)) / S2, atol=1e-15)  # tilt after the mirror


def test_refractor_index_and_snell():  # test_pyrayt_materials.py:56-110
    f = _plane_material_trace(su.MAT_GLASS_CONST, [1.6], (0, 0, -1))
    assert f[3, 1] == 1.6  # entering: index updated
    f = _plane_material_trace(su.MAT_GLASS_CONST, [1.6], (0, 0, 1), index=20)
    assert f[3, 1] == 1.0  # exiting: index set to 1
    f = _plane_material_trace(su.MAT_GLASS_CONST, [1.6], (0, 1, -1))
    ang = np.arctan(abs(f[13, 1] / f[14, 1]))
    assert ang == pytest.approx(np.arcsin(np.sin(np.pi / 4) / 1.6), abs=1e-12)
    f = _plane_material_trace(su.MAT_GLASS_CONST, [1.6], (0, np.sin(0.1), np.cos(0.1)), index=1.6)
    ang = np.arctan(abs(f[13, 1] / f[14, 1]))
    assert ang == pytest.approx(np.arcsin(np.sin(0.1) * 1.6), abs=1e-12)
    f = _plane_material_trace(su.MAT_GLASS_CONST, [1.6], (0, 1, 1), index=1.6)  # TIR
    ang = np.arctan(abs(f[13, 1] / f[14, 1]))
    assert ang == pytest.approx(np.pi / 4, abs=1e-12) and f[3, 1] == 1.6


def test_sellmeier_refraction():  # test_pyrayt_materials.py:136-169
    f = _plane_material_trace(su.MAT_GLASS_SELLMEIER, [1, 0, 0, 1, 0, 0], (0, 0, -1), wavelength=2.0)
    assert f[3, 1] == pytest.approx(np.sqrt(7 / 3), abs=1e-15)
    f = _plane_material_trace(su.MAT_GLASS_SELLMEIER, [1, 0, 0, 1, 0, 0], (0, 1, -1), wavelength=2.0)
    ang = np.arctan(abs(f[13, 1] / f[14, 1]))
    assert ang == pytest.approx(np.arcsin(np.sqrt(3 / 7) * S2 / 2), abs=1e-12)


# ---------------------------------------------------------------- test_core.py

def test_trace_to_absorbing_plane():  # test_core.py:45-52: 10 rows, x1 == 3.0
    scene = su.build([su.Leaf(su.PLANE, [4, 4], mat=su.MAT_ABSORBER, world=su.translate(3, 0, 0) @ su.rot_y(90))])
    ys = np.linspace(-0.5, 0.5, 10)
    rays = su.make_rays(np.stack([np.zeros(10), ys, np.zeros(10)], 1), np.tile([1.0, 0, 0], (10, 1)))
    frame, ctr = oracle.trace(scene, rays, 10)
    assert frame.shape == (15, 10)
    np.testing.assert_allclose(frame[9], 3.0)
    assert np.array_equal(frame[4], np.arange(10))


def test_facing_mirrors_fill_the_generation_limit():  # test_core.py:54-66
    m1 = su.Leaf(su.PLANE, [4, 4], mat=su.MAT_MIRROR, world=su.translate(3, 0, 0) @ su.rot_y(90))
    m2 = su.Leaf(su.PLANE, [4, 4], mat=su.MAT_MIRROR, world=su.translate(-3, 0, 0) @ su.rot_y(90))
    rays = su.make_rays(np.zeros((5, 3)), np.tile([1.0, 0, 0], (5, 1)))
    frame, ctr = oracle.trace(su.build([m1, m2]), rays, 10)
    assert frame.shape[1] == 10 * 5
    assert set(frame[0]) == set(range(10))
    assert ctr["limit_rays"] == 5


def test_rows_are_ordered_by_generation_then_id():  # pyrayt/_pyrayt.py:186,:428-435
    scene, rays, _, gl = __import__("tests.helpers", fromlist=["load_case"]).load_case("thick_lens_zoo")
    frame, _ = oracle.trace(scene, rays, gl)
    key = frame[0] * 1e9 + frame[4]
    assert np.all(np.diff(key) > 0)


def test_empty_input():
    scene = su.build([su.Leaf(su.SPHERE, [1])])
    frame, ctr = oracle.trace(scene, np.zeros((13, 0)), 10)
    assert frame.shape == (15, 0) and ctr["rays"] == 0
