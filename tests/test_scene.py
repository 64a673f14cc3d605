"""The scene flattener on duck-typed objects (and on the live reference when it is present)."""
import numpy as np
import pytest

from oracle import ref_shim
from pyrayt_b200 import scene as sc
from tests import fakes, scene_util as su


def _lens():
    glass = fakes.SellmeierRefractor(1.0, 0.2, 1.1, 0.006, 0.02, 103.0)
    a = fakes.Surface(fakes.Sphere(2.0), glass, su.translate(1.9, 0, 0))
    b = fakes.Surface(fakes.Sphere(2.0), glass, su.translate(-1.9, 0, 0))
    c = fakes.Surface(fakes.Cylinder(0.5, -1, 1), glass, su.rot_y(90))
    inner = fakes.CSG(a, b, 2, (-0.1, 0.1, -1, 1, -1, 1))
    return fakes.CSG(inner, c, 3, (-0.1, 0.1, -0.5, 0.5, -0.5, 0.5)), (a, b, c)


def test_flatten_postfix_layout():
    lens, (a, b, c) = _lens()
    det = fakes.Surface(fakes.Plane(2, 3), fakes._AbsorbingMaterial())
    mir = fakes.Surface(fakes.Cube((-1, -2, -3), (1, 2, 3)), fakes._ReflectingMaterial())
    s = sc.flatten([lens, det, mir])
    assert s.n_components == 3 and s.n_leaves == 5 and s.n_nodes == 7
    assert list(s.comp_node_begin) == [0, 5, 6, 7]
    assert list(s.node_kind) == [0, 0, 2, 0, 3, 0, 0]  # A B INTERSECT C DIFFERENCE | det | mirror
    assert list(s.node_leaf) == [0, 1, -1, 2, -1, 3, 4]
    assert list(s.leaf_type) == [1, 1, 5, 3, 4]
    assert list(s.leaf_sid) == [a.get_id(), b.get_id(), c.get_id(), det.get_id(), mir.get_id()]
    assert list(s.leaf_mat) == [3, 3, 3, 0, 1]
    assert list(s.leaf_nscale) == [1, 1, -1, 1, 1]  # DIFFERENCE inverts its right child (csg.py:87-89)
    assert np.allclose(s.leaf_matp[0], [1.0, 0.2, 1.1, 0.006, 0.02, 103.0])
    assert np.allclose(s.node_aabb[4], (-0.1, 0.1, -0.5, 0.5, -0.5, 0.5))
    assert np.allclose(s.leaf_param[4], (-1, 1, -2, 2, -3, 3))
    assert np.allclose(s.leaf_param[2][:4], (0.5, -1, 1, 1))
    assert np.allclose(s.leaf_obj[0].reshape(4, 4), np.linalg.inv(su.translate(1.9, 0, 0)))
    assert s.component_slots(0) == 6 and s.component_slots(1) == 2


def test_json_round_trip_is_exact():
    lens, _ = _lens()
    s = sc.flatten([lens])
    t = sc.FlatScene.from_json(s.to_json())
    for k in ("comp_node_begin", "node_kind", "node_leaf", "node_aabb", "leaf_type", "leaf_obj", "leaf_param",
              "leaf_nscale", "leaf_sid", "leaf_mat", "leaf_matp"):
        assert np.array_equal(getattr(s, k), getattr(t, k)), k


def test_unsupported_things_are_hard_errors():
    class Torus:
        pass

    class MyGlass(fakes.TracableMaterial):
        def index_at(self, w):
            return 1.5

    with pytest.raises(sc.SceneError):
        sc.flatten([fakes.Surface(Torus(), fakes._AbsorbingMaterial())])
    with pytest.raises(sc.SceneError):
        sc.flatten([fakes.Surface(fakes.Sphere(1), MyGlass())])
    with pytest.raises(sc.SceneError):
        sc.flatten([object()])
    s = fakes.Surface(fakes.Sphere(1), fakes._AbsorbingMaterial())
    with pytest.raises(sc.SceneError):  # the same surface inside two different components
        sc.flatten([fakes.CSG(s, fakes.Surface(fakes.Sphere(1), fakes._AbsorbingMaterial()), 2, (-1, 1) * 3),
                    fakes.CSG(s, fakes.Surface(fakes.Sphere(2), fakes._AbsorbingMaterial()), 2, (-1, 1) * 3)])
    too_many = [fakes.Surface(fakes.Sphere(1), fakes._AbsorbingMaterial()) for _ in range(sc.MAX_LEAVES + 1)]
    with pytest.raises(sc.SceneError):
        sc.flatten(too_many)


def test_repeated_component_is_listed_once():
    """The reference accepts the same component twice; the copy can never win the strict `<`."""
    s = fakes.Surface(fakes.Sphere(1), fakes._AbsorbingMaterial())
    t = fakes.Surface(fakes.Sphere(2), fakes._AbsorbingMaterial())
    flat = sc.flatten([s, t, s])
    assert flat.n_components == 2 and list(flat.leaf_sid) == [s.get_id(), t.get_id()]


def test_material_subclasses_that_override_the_law_are_rejected():
    """A subclass of a reference material that re-defines trace() / index_at() must not be traced as its base."""
    class HalfMirror(fakes._ReflectingMaterial):
        def trace(self, surface, ray_set):
            return ray_set

    class Cauchy(fakes.BasicRefractor):
        def index_at(self, wavelength):
            return 1.5 + 0.004 / wavelength ** 2

    class Renamed(fakes.SellmeierRefractor):  # no override: still the reference's law
        pass

    for bad in (HalfMirror(), Cauchy(1.5)):
        with pytest.raises(sc.SceneError, match="overrides"):
            sc.flatten([fakes.Surface(fakes.Sphere(1), bad)])
    patched = fakes._AbsorbingMaterial()
    patched.trace = lambda surface, ray_set: ray_set  # instance-level override
    with pytest.raises(sc.SceneError, match="overrides"):
        sc.flatten([fakes.Surface(fakes.Sphere(1), patched)])
    flat = sc.flatten([fakes.Surface(fakes.Sphere(1), Renamed(1.0, 0.2, 1.0, 0.006, 0.02, 103.0))])
    assert flat.leaf_mat[0] == sc.MAT_GLASS_SELLMEIER


def test_untraceable_material_is_flagged_not_rejected():
    s = sc.flatten([fakes.Surface(fakes.Sphere(1), fakes.Gooch())])
    assert s.leaf_mat[0] == sc.MAT_UNTRACEABLE


@pytest.mark.reference
@pytest.mark.skipif(not ref_shim.available(), reason="PyRayT reference tree not present")
def test_flatten_live_reference_factories():
    """Tree shapes of SURVEY.md 8(a3) straight from the reference's component factories."""
    pyrayt = ref_shim.load()
    import pyrayt.components as pc

    def kinds(comp):
        s = sc.flatten([comp])
        names = {1: "Sphere", 2: "Paraboloid", 3: "Plane", 4: "Cube", 5: "Cylinder"}
        ops = {1: "U", 2: "I", 3: "D"}
        out = []
        for k, l in zip(s.node_kind, s.node_leaf):
            out.append(names[int(s.leaf_type[l])] if k == 0 else ops[int(k)])
        return out, s

    assert kinds(pc.biconvex_lens(2, 2, 0.25, aperture=1))[0] == ["Sphere", "Sphere", "I", "Cylinder", "I"]
    assert kinds(pc.plano_convex_lens(2, 0.25, aperture=1))[0] == ["Sphere", "Cylinder", "I"]
    assert kinds(pc.thick_lens(6, -6, 0.5, aperture=1))[0] == ["Cylinder", "Sphere", "I", "Sphere", "I"]
    k, s = kinds(pc.thick_lens(-6, 6, 0.5, aperture=1))
    assert k == ["Cylinder", "Sphere", "D", "Sphere", "D"] and list(s.leaf_nscale) == [1, -1, -1]
    assert kinds(pc.thick_lens(6, -6, 0.5, aperture=(1, 1)))[0] == ["Cube", "Sphere", "I", "Sphere", "I"]
    k, s = kinds(pc.spherical_mirror(10, 1, aperture=3))
    assert k == ["Cylinder", "Sphere", "D"] and list(s.leaf_mat) == [0, 1]
    k, s = kinds(pc.parabolic_mirror(5, 1, aperture=4))
    assert k == ["Cylinder", "Paraboloid", "D"] and list(s.leaf_mat) == [0, 1]
    k, s = kinds(pc.equilateral_prism(1, 1))
    assert k == ["Cube", "Cube", "D", "Cube", "D"] and set(s.leaf_mat) == {3}
    k, s = kinds(pc.aperture((1, 1), 0.5))
    assert k == ["Plane", "Cylinder", "D"] and list(s.leaf_mat) == [0, 4]  # Gooch cylinder (Q9)
    assert kinds(pc.baffle((1, 1)))[0] == ["Plane"]
    # object matrices stay affine with an exact last row
    assert np.array_equal(s.leaf_obj.reshape(-1, 4, 4)[:, 3, :], np.tile([0.0, 0, 0, 1], (2, 1)))
