"""Scene flattener: live PyRayT / tinygfx objects -> flat SoA arrays (``prt_scene_desc``).

The reference keeps the scene as a graph of Python objects; the trace kernel
wants it flat.  This module walks the components a ``RayTracer`` was given and
emits, per component, a postfix CSG program over leaf surfaces.  It is
duck-typed on the reference's attributes so the reference does not have to be
importable (the GPU box has no ``/root/reference``):

* CSG node  -- ``_l_child``, ``_r_child``, ``_operation`` (``.value`` 1/2/3 =
  UNION/INTERSECT/DIFFERENCE, tinygfx/g3d/csg.py:7-10) and the world-space
  bounding box ``_aobb.axis_spans`` that ``CSGSurface.intersect`` culls with
  (csg.py:93-128).
* leaf      -- ``TracerSurface``: ``_surface_primitive`` (type name + parameters,
  tinygfx/g3d/primitives.py:223,:309-310,:429-430,:510,:630-634), the
  world->object matrix ``_get_object_transform()`` (world_objects.py:171-178),
  ``_normal_scale`` (:305,:319-323), ``get_id()`` (:33-40) and ``material``
  (pyrayt/materials.py).

Anything the kernel cannot represent is a hard error: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import json
from dataclasses import dataclass, field
from typing import Iterable, List

import numpy as np

# prt_prim / prt_node_kind / prt_material (include/pyrayt_b200.h)
PRIM_CODES = {"Sphere": 1, "Paraboloid": 2, "Plane": 3, "Cube": 4, "Cylinder": 5}
NODE_LEAF, NODE_UNION, NODE_INTERSECT, NODE_DIFFERENCE = 0, 1, 2, 3
MAT_ABSORBER, MAT_MIRROR, MAT_GLASS_CONST, MAT_GLASS_SELLMEIER, MAT_UNTRACEABLE = 0, 1, 2, 3, 4

MAX_SLOTS = 32
MAX_LEAVES = 4096  # PRT_MAX_LEAVES
MAX_NODES = 8192  # PRT_MAX_NODES


class SceneError(ValueError):
    """The scene holds something the B200 path cannot represent."""


class PrtSceneDesc(ctypes.Structure):
    """ctypes mirror of ``prt_scene_desc`` (include/pyrayt_b200.h)."""

    _fields_ = [
        ("n_components", ctypes.c_int32),
        ("n_nodes", ctypes.c_int32),
        ("n_leaves", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("comp_node_begin", ctypes.POINTER(ctypes.c_int32)),
        ("node_kind", ctypes.POINTER(ctypes.c_int32)),
        ("node_leaf", ctypes.POINTER(ctypes.c_int32)),
        ("node_aabb", ctypes.POINTER(ctypes.c_double)),
        ("leaf_type", ctypes.POINTER(ctypes.c_int32)),
        ("leaf_obj", ctypes.POINTER(ctypes.c_double)),
        ("leaf_param", ctypes.POINTER(ctypes.c_double)),
        ("leaf_nscale", ctypes.POINTER(ctypes.c_double)),
        ("leaf_sid", ctypes.POINTER(ctypes.c_int64)),
        ("leaf_mat", ctypes.POINTER(ctypes.c_int32)),
        ("leaf_matp", ctypes.POINTER(ctypes.c_double)),
    ]


def _material_record(material) -> tuple:
    """(kind, 6 params) for a reference material object (pyrayt/materials.py)."""
    if not hasattr(material, "trace"):
        # e.g. the Gooch BLACK cylinder inside components.aperture()
        return MAT_UNTRACEABLE, [0.0] * 6
    # The kernel implements exactly four laws.  A material is recognised by the reference class it derives
    # from, and only if nothing between its own class and that base re-defines trace() / index_at() (a
    # partial mirror, a custom dispersion law, ...): such a material cannot run on the device and there is
    # no CPU fallback, so it is an error rather than being silently traced as its base class.
    known = {"_AbsorbingMaterial": MAT_ABSORBER, "_ReflectingMaterial": MAT_MIRROR,
             "SellmeierRefractor": MAT_GLASS_SELLMEIER, "BasicRefractor": MAT_GLASS_CONST}
    mro = type(material).__mro__
    base = next((k for k in mro if k.__name__ in known), None)
    if base is not None:
        overriding = [k.__name__ for k in mro[: mro.index(base)] if "trace" in vars(k) or "index_at" in vars(k)]
        if any(name in getattr(material, "__dict__", {}) for name in ("trace", "index_at")):
            overriding.append("the instance itself")
        if overriding:
            raise SceneError(
                f"material {type(material).__name__} overrides trace()/index_at() of {base.__name__} "
                f"(in {', '.join(overriding)}); only the reference's own absorber, mirror, BasicRefractor and "
                "SellmeierRefractor laws run on the B200 path (no CPU fallback)")
        kind = known[base.__name__]
        if kind == MAT_GLASS_SELLMEIER:
            return kind, [float(getattr(material, k)) for k in ("b1", "b2", "b3", "c1", "c2", "c3")]
        if kind == MAT_GLASS_CONST:
            return kind, [float(material._refractive_index)] + [0.0] * 5
        return kind, [0.0] * 6
    raise SceneError(
        f"material {type(material).__name__} has a custom trace()/index_at(); only absorber, mirror, "
        "BasicRefractor and SellmeierRefractor run on the B200 path (no CPU fallback)"
    )


def _primitive_record(prim) -> tuple:
    name = type(prim).__name__
    if name not in PRIM_CODES:
        raise SceneError(f"unsupported surface primitive {name}")
    p = [0.0] * 6
    if name == "Sphere":
        p[0] = float(prim._radius)
    elif name == "Paraboloid":
        p[0], p[1] = float(prim._focus), float(prim._height)
    elif name == "Plane":
        p[0], p[1] = float(prim._width), float(prim._length)
    elif name == "Cube":
        spans = np.asarray(prim.axis_spans, dtype=np.float64)
        p = [float(x) for x in spans.reshape(6)]
    elif name == "Cylinder":
        p[0], p[1], p[2] = float(prim._radius), float(prim._h_min), float(prim._h_max)
        p[3] = 1.0 if prim._capped else 0.0
    return PRIM_CODES[name], p


@dataclass
class FlatScene:
    """Host SoA copy of the scene; field names follow ``prt_scene_desc``."""

    comp_node_begin: np.ndarray
    node_kind: np.ndarray
    node_leaf: np.ndarray
    node_aabb: np.ndarray
    leaf_type: np.ndarray
    leaf_obj: np.ndarray
    leaf_param: np.ndarray
    leaf_nscale: np.ndarray
    leaf_sid: np.ndarray
    leaf_mat: np.ndarray
    leaf_matp: np.ndarray
    _keep: list = field(default_factory=list, repr=False)

    @property
    def n_components(self) -> int:
        return len(self.comp_node_begin) - 1

    @property
    def n_nodes(self) -> int:
        return len(self.node_kind)

    @property
    def n_leaves(self) -> int:
        return len(self.leaf_type)

    def component_slots(self, c: int) -> int:
        b, e = self.comp_node_begin[c], self.comp_node_begin[c + 1]
        return 2 * int(np.count_nonzero(self.node_kind[b:e] == NODE_LEAF))

    def validate(self) -> None:
        if self.n_leaves > MAX_LEAVES or self.n_nodes > MAX_NODES:
            raise SceneError(f"scene too large: {self.n_leaves} leaves / {self.n_nodes} nodes")
        if len(np.unique(self.leaf_sid)) != self.n_leaves:
            raise SceneError("a surface appears more than once in the component list")
        for c in range(self.n_components):
            if self.component_slots(c) > MAX_SLOTS:
                raise SceneError(f"component {c} has more than {MAX_SLOTS // 2} leaf surfaces")
            depth = 0
            for k in self.node_kind[self.comp_node_begin[c] : self.comp_node_begin[c + 1]]:
                depth += 1 if k == NODE_LEAF else -1
                if depth < 1:
                    raise SceneError("malformed postfix CSG program")
            if depth != 1:
                raise SceneError("malformed postfix CSG program")
        last = self.leaf_obj.reshape(-1, 4, 4)[:, 3, :]
        if not np.array_equal(last, np.tile([0.0, 0.0, 0.0, 1.0], (self.n_leaves, 1))):
            raise SceneError("projective (non-affine) object transforms are not supported")

    def as_desc(self) -> PrtSceneDesc:
        """ctypes view; the arrays stay owned (and kept alive) by this object."""
        d = PrtSceneDesc()
        d.n_components, d.n_nodes, d.n_leaves = self.n_components, self.n_nodes, self.n_leaves

        def ptr(a, ct):
            return a.ctypes.data_as(ctypes.POINTER(ct))

        d.comp_node_begin = ptr(self.comp_node_begin, ctypes.c_int32)
        d.node_kind = ptr(self.node_kind, ctypes.c_int32)
        d.node_leaf = ptr(self.node_leaf, ctypes.c_int32)
        d.node_aabb = ptr(self.node_aabb, ctypes.c_double)
        d.leaf_type = ptr(self.leaf_type, ctypes.c_int32)
        d.leaf_obj = ptr(self.leaf_obj, ctypes.c_double)
        d.leaf_param = ptr(self.leaf_param, ctypes.c_double)
        d.leaf_nscale = ptr(self.leaf_nscale, ctypes.c_double)
        d.leaf_sid = ptr(self.leaf_sid, ctypes.c_int64)
        d.leaf_mat = ptr(self.leaf_mat, ctypes.c_int32)
        d.leaf_matp = ptr(self.leaf_matp, ctypes.c_double)
        return d

    def fingerprint(self) -> bytes:
        """Cheap identity of the flattened scene (all arrays, byte for byte)."""
        return b"|".join(getattr(self, k).tobytes() for k in (
            "comp_node_begin", "node_kind", "node_leaf", "node_aabb", "leaf_type", "leaf_obj", "leaf_param",
            "leaf_nscale", "leaf_sid", "leaf_mat", "leaf_matp"))

    # ---- fixtures: scenes travel to the GPU box as JSON (floats via repr: exact round trip)
    def to_json(self) -> str:
        out = {}
        for k in self.__dataclass_fields__:
            if k.startswith("_"):
                continue
            a = getattr(self, k)
            out[k] = [float(x).hex() for x in a.reshape(-1)] if a.dtype == np.float64 else [int(x) for x in a.reshape(-1)]
        return json.dumps(out, indent=0)

    @classmethod
    def from_json(cls, text: str) -> "FlatScene":
        raw = json.loads(text)
        dt = {
            "comp_node_begin": np.int32, "node_kind": np.int32, "node_leaf": np.int32, "leaf_type": np.int32,
            "leaf_mat": np.int32, "leaf_sid": np.int64,
        }
        kw = {}
        for k, v in raw.items():
            if k in dt:
                kw[k] = np.asarray(v, dtype=dt[k])
            else:
                kw[k] = np.asarray([float.fromhex(x) for x in v], dtype=np.float64)
        for k, w in (("node_aabb", 6), ("leaf_obj", 16), ("leaf_param", 6), ("leaf_matp", 6)):
            kw[k] = kw[k].reshape(-1, w)
        s = cls(**kw)
        s.validate()
        return s


def flatten(components: Iterable) -> FlatScene:
    """Flatten the component list of a ``RayTracer`` (pyrayt/_pyrayt.py:234-260)."""
    if not hasattr(components, "__iter__"):
        components = (components,)
    comp_begin: List[int] = [0]
    node_kind: List[int] = []
    node_leaf: List[int] = []
    node_aabb: List[List[float]] = []
    leaf_type: List[int] = []
    leaf_obj: List[np.ndarray] = []
    leaf_param: List[List[float]] = []
    leaf_nscale: List[float] = []
    leaf_sid: List[int] = []
    leaf_mat: List[int] = []
    leaf_matp: List[List[float]] = []

    def emit(obj) -> None:
        if hasattr(obj, "_l_child") and hasattr(obj, "_r_child"):
            emit(obj._l_child)
            emit(obj._r_child)
            op = int(getattr(obj._operation, "value", obj._operation))
            if op not in (NODE_UNION, NODE_INTERSECT, NODE_DIFFERENCE):
                raise SceneError(f"unknown CSG operation {obj._operation}")
            node_kind.append(op)
            node_leaf.append(-1)
            node_aabb.append([float(x) for x in np.asarray(obj._aobb.axis_spans, dtype=np.float64).reshape(6)])
            return
        prim = getattr(obj, "_surface_primitive", None)
        if prim is None:
            raise SceneError(f"{type(obj).__name__} is neither a CSGSurface nor a TracerSurface")
        code, params = _primitive_record(prim)
        kind, matp = _material_record(obj.material)
        node_kind.append(NODE_LEAF)
        node_leaf.append(len(leaf_type))
        node_aabb.append([0.0] * 6)
        leaf_type.append(code)
        leaf_obj.append(np.array(obj._get_object_transform(), dtype=np.float64).reshape(16))
        leaf_param.append(params)
        leaf_nscale.append(float(obj._normal_scale))
        leaf_sid.append(int(obj.get_id()))
        leaf_mat.append(kind)
        leaf_matp.append(matp)

    seen = set()
    for comp in components:
        # the same component object listed twice: the reference intersects it twice and the second copy can
        # never win _st_propagate's strict `<` (pyrayt/_pyrayt.py:384), so it is dropped here
        if id(comp) in seen:
            continue
        seen.add(id(comp))
        emit(comp)
        comp_begin.append(len(node_kind))

    scene = FlatScene(
        comp_node_begin=np.asarray(comp_begin, dtype=np.int32),
        node_kind=np.asarray(node_kind, dtype=np.int32),
        node_leaf=np.asarray(node_leaf, dtype=np.int32),
        node_aabb=np.asarray(node_aabb, dtype=np.float64).reshape(-1, 6),
        leaf_type=np.asarray(leaf_type, dtype=np.int32),
        leaf_obj=np.asarray(leaf_obj, dtype=np.float64).reshape(-1, 16),
        leaf_param=np.asarray(leaf_param, dtype=np.float64).reshape(-1, 6),
        leaf_nscale=np.asarray(leaf_nscale, dtype=np.float64),
        leaf_sid=np.asarray(leaf_sid, dtype=np.int64),
        leaf_mat=np.asarray(leaf_mat, dtype=np.int32),
        leaf_matp=np.asarray(leaf_matp, dtype=np.float64).reshape(-1, 6),
    )
    scene.validate()
    return scene
