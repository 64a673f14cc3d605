"""Hand-built flat scenes for tests (no reference objects needed)."""
import numpy as np

from pyrayt_b200.scene import (FlatScene, MAT_ABSORBER, MAT_GLASS_CONST, MAT_GLASS_SELLMEIER, MAT_MIRROR,
                               MAT_UNTRACEABLE, NODE_DIFFERENCE, NODE_INTERSECT, NODE_LEAF, NODE_UNION)

SPHERE, PARABOLOID, PLANE, CUBE, CYLINDER = 1, 2, 3, 4, 5
BIG_BOX = (-1e6, 1e6, -1e6, 1e6, -1e6, 1e6)


def translate(x=0.0, y=0.0, z=0.0):
    m = np.eye(4)
    m[:3, 3] = (x, y, z)
    return m


def rot_y(deg):
    a = np.radians(deg)
    m = np.eye(4)
    m[0, 0] = m[2, 2] = np.cos(a)
    m[2, 0] = -np.sin(a)
    m[0, 2] = np.sin(a)
    return m


def rot_z(deg):
    a = np.radians(deg)
    m = np.eye(4)
    m[0, 0] = m[1, 1] = np.cos(a)
    m[0, 1] = -np.sin(a)
    m[1, 0] = np.sin(a)
    return m


class Leaf:
    def __init__(self, ptype, params, mat=MAT_ABSORBER, matp=(), world=None, sid=None, nscale=1.0):
        self.ptype, self.params, self.mat, self.matp = ptype, list(params), mat, list(matp)
        self.world = np.eye(4) if world is None else np.asarray(world, dtype=np.float64)
        self.sid, self.nscale = sid, nscale


class Node:
    def __init__(self, op, left, right, aabb=BIG_BOX):
        self.op, self.left, self.right, self.aabb = op, left, right, aabb


def union(a, b, aabb=BIG_BOX):
    return Node(NODE_UNION, a, b, aabb)


def intersect(a, b, aabb=BIG_BOX):
    return Node(NODE_INTERSECT, a, b, aabb)


def difference(a, b, aabb=BIG_BOX):
    return Node(NODE_DIFFERENCE, a, b, aabb)


def leaf_box(o: Leaf):
    """World-space AABB of a leaf the way the reference builds it (bounding cube corners -> world)."""
    t, p = o.ptype, o.params
    if t == SPHERE:
        lo, hi = [-p[0]] * 3, [p[0]] * 3
    elif t == PARABOLOID:
        r = np.sqrt(4 * p[0] * p[1])
        lo, hi = [-r, -r, 0], [r, r, p[1]]
    elif t == PLANE:
        lo, hi = [-p[0] / 2, -p[1] / 2, -0.01], [p[0] / 2, p[1] / 2, 0.01]
    elif t == CUBE:
        lo, hi = [p[0], p[2], p[4]], [p[1], p[3], p[5]]
    else:
        lo, hi = [-p[0], -p[0], p[1]], [p[0], p[0], p[2]]
    corners = np.array([[x, y, z, 1.0] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])]).T
    w = o.world @ corners
    return np.stack([w[:3].min(axis=1), w[:3].max(axis=1)], axis=1)  # (3, 2)


def tight_box(o):
    """Conservative box of a (sub)tree: INTERSECT overlap, DIFFERENCE left child, UNION hull."""
    if isinstance(o, Leaf):
        return leaf_box(o)
    l, r = tight_box(o.left), tight_box(o.right)
    if o.op == NODE_DIFFERENCE:
        return l
    if o.op == NODE_INTERSECT:
        return np.stack([np.maximum(l[:, 0], r[:, 0]), np.minimum(l[:, 1], r[:, 1])], axis=1)
    return np.stack([np.minimum(l[:, 0], r[:, 0]), np.maximum(l[:, 1], r[:, 1])], axis=1)


def assign_boxes(o, rng):
    """Give every CSG node a box: tight (pruning provable), huge, or deliberately too small."""
    if isinstance(o, Leaf):
        return
    assign_boxes(o.left, rng)
    assign_boxes(o.right, rng)
    b = tight_box(o)
    if np.any(b[:, 0] >= b[:, 1]):
        o.aabb = BIG_BOX
        return
    mode = int(rng.integers(0, 4))
    if mode == 0:
        o.aabb = BIG_BOX
    elif mode == 3:  # shrunk: the reference-style cull now hides parts of the solid; pruning must stay off
        c, h = b.mean(axis=1), (b[:, 1] - b[:, 0]) / 2
        o.aabb = tuple(np.stack([c - 0.6 * h, c + 0.6 * h], axis=1).reshape(6))
    else:
        o.aabb = tuple(b.reshape(6))


def build(components) -> FlatScene:
    comp_begin, kind, nleaf, aabb = [0], [], [], []
    ltype, lobj, lprm, lns, lsid, lmat, lmatp = [], [], [], [], [], [], []

    def emit(o):
        if isinstance(o, Node):
            emit(o.left)
            emit(o.right)
            kind.append(o.op)
            nleaf.append(-1)
            aabb.append(list(o.aabb))
        else:
            kind.append(NODE_LEAF)
            nleaf.append(len(ltype))
            aabb.append([0.0] * 6)
            ltype.append(o.ptype)
            lobj.append(np.linalg.inv(o.world).reshape(16))
            lprm.append(o.params + [0.0] * (6 - len(o.params)))
            lns.append(o.nscale)
            lsid.append(100 + len(lsid) if o.sid is None else o.sid)
            lmat.append(o.mat)
            lmatp.append(o.matp + [0.0] * (6 - len(o.matp)))

    for c in components:
        emit(c)
        comp_begin.append(len(kind))
    obj = np.asarray(lobj, dtype=np.float64).reshape(-1, 16)
    obj[:, 12:15] = 0.0
    obj[:, 15] = 1.0
    s = FlatScene(
        comp_node_begin=np.asarray(comp_begin, dtype=np.int32), node_kind=np.asarray(kind, dtype=np.int32),
        node_leaf=np.asarray(nleaf, dtype=np.int32), node_aabb=np.asarray(aabb, dtype=np.float64).reshape(-1, 6),
        leaf_type=np.asarray(ltype, dtype=np.int32), leaf_obj=obj,
        leaf_param=np.asarray(lprm, dtype=np.float64).reshape(-1, 6), leaf_nscale=np.asarray(lns, dtype=np.float64),
        leaf_sid=np.asarray(lsid, dtype=np.int64), leaf_mat=np.asarray(lmat, dtype=np.int32),
        leaf_matp=np.asarray(lmatp, dtype=np.float64).reshape(-1, 6))
    s.validate()
    return s


def make_rays(origins, directions, wavelength=0.633, index=1.0, intensity=100.0):
    """(13,N) RaySet array (pyrayt/_pyrayt.py:13-44) from (N,3) origins / directions."""
    o = np.atleast_2d(np.asarray(origins, dtype=np.float64))
    d = np.atleast_2d(np.asarray(directions, dtype=np.float64))
    n = o.shape[0]
    r = np.zeros((13, n))
    r[0:3] = o.T
    r[3] = 1
    r[4:7] = d.T
    r[9] = intensity
    r[10] = wavelength
    r[11] = index
    r[12] = np.arange(n)
    return r


def random_scene_and_rays(seed, n_rays=512):
    """A random but valid scene exercising every primitive / operation / material."""
    rng = np.random.default_rng(seed)
    glass = dict(mat=MAT_GLASS_CONST, matp=[1.5])
    bk7 = dict(mat=MAT_GLASS_SELLMEIER, matp=[1.03961212, 0.231792344, 1.01046945, 6.00069867e-3, 2.00179144e-2, 103.560653])

    def pose():
        m = translate(*rng.uniform(-3, 3, 3)) @ rot_z(rng.uniform(0, 360)) @ rot_y(rng.uniform(0, 360))
        if rng.random() < 0.4:  # non-uniform scale, like WorldObject.scale()
            m = m @ np.diag(list(rng.uniform(0.6, 1.8, 3)) + [1.0])
        return m

    def prim(kw):
        t = int(rng.integers(1, 6))
        if t == SPHERE:
            return Leaf(SPHERE, [rng.uniform(0.5, 1.5)], world=pose(), **kw)
        if t == PARABOLOID:
            return Leaf(PARABOLOID, [rng.uniform(0.3, 1.0), rng.uniform(0.5, 2.0)], world=pose(), **kw)
        if t == PLANE:
            return Leaf(PLANE, [rng.uniform(1, 3), rng.uniform(1, 3)], world=pose(), **kw)
        if t == CUBE:
            lo = -rng.uniform(0.3, 1.2, 3)
            hi = rng.uniform(0.3, 1.2, 3)
            return Leaf(CUBE, [lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]], world=pose(), **kw)
        return Leaf(CYLINDER, [rng.uniform(0.3, 1.0), -rng.uniform(0.3, 1.5), rng.uniform(0.3, 1.5), 1.0], world=pose(), **kw)

    mats = [glass, bk7, dict(mat=MAT_MIRROR), dict(mat=MAT_ABSORBER)]
    comps = []
    for _ in range(int(rng.integers(3, 7))):
        kw = mats[int(rng.integers(0, len(mats)))]
        shape = int(rng.integers(0, 4))
        ops = (union, intersect, difference)
        if shape == 0:
            comps.append(prim(kw))
        elif shape == 1:
            comps.append(ops[int(rng.integers(0, 3))](prim(kw), prim(kw)))
        elif shape == 2:
            comps.append(ops[int(rng.integers(0, 3))](ops[int(rng.integers(0, 3))](prim(kw), prim(kw)), prim(kw)))
        else:  # right-nested
            comps.append(ops[int(rng.integers(0, 3))](prim(kw), ops[int(rng.integers(0, 3))](prim(kw), prim(kw))))
    for c in comps:
        assign_boxes(c, rng)
    # enclose in absorbing walls so most rays end on something
    for ax in range(3):
        for sgn in (-1, 1):
            w = np.eye(4)
            if ax == 0:
                w = rot_y(90)
            elif ax == 1:
                w = rot_z(90) @ rot_y(90)
            w = translate(*(np.eye(3)[ax] * sgn * 8)) @ w
            comps.append(Leaf(PLANE, [20, 20], world=w, mat=MAT_ABSORBER))
    o = rng.uniform(-5, 5, (n_rays, 3))
    d = rng.normal(size=(n_rays, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = make_rays(o, d)
    rays[10] = rng.uniform(0.45, 0.7, n_rays)
    return build(comps), rays


def grazing_and_seam_case():
    """A unit mirror sphere and two absorber planes that share the edge y = 0 at x = 5, with rays aimed
    2e-10 inside / outside the sphere's silhouette, through the plane seam, at the outer plane edge, and
    well away from all of them.  Returns (scene, rays, expected (grazing, seam) flags per ray)."""
    sph = Leaf(SPHERE, [1.0], mat=MAT_MIRROR)
    pl_a = Leaf(PLANE, [2.0, 2.0], world=translate(5, 1.0, 0) @ rot_y(90))
    pl_b = Leaf(PLANE, [2.0, 2.0], world=translate(5, -1.0, 0) @ rot_y(90))
    scene = build([sph, pl_a, pl_b])
    o = [[-3, 1 - 2e-10, 0], [-3, 1 + 2e-10, 0], [-3, 0.5, 0], [3, 3e-10, 0.2], [3, 0.3, 0.2], [3, 2.0 - 1e-10, 0.1]]
    rays = make_rays(o, [[1, 0, 0]] * 6)
    # ray 0 / 1: sphere <-> the plane behind it (a change of surface), ray 3: plane a <-> plane b,
    # ray 5: plane a <-> nothing
    expected = [(0, 1), (0, 1), (0, 0), (0, 1), (0, 0), (1, 0)]
    return scene, rays, expected


def lenslet_array(nx, ny, pitch=2.0, radius=0.9, thickness=0.6, curvature=3.0, index=1.5):
    """nx x ny biconvex lenslets in the plane x = 0 (each cylinder & sphere & sphere, a left-deep tree with
    tight boxes) and an absorbing detector behind them: 3 nx ny + 1 leaves.  Returns (scene, lens centres)."""
    comps, centres = [], []
    for i in range(nx):
        for j in range(ny):
            cy, cz = (i - (nx - 1) / 2) * pitch, (j - (ny - 1) / 2) * pitch
            centres.append((cy, cz))
            cyl = Leaf(CYLINDER, [radius, -thickness / 2, thickness / 2, 1.0], mat=MAT_GLASS_CONST, matp=[index],
                       world=translate(0, cy, cz) @ rot_y(90))
            s_a = Leaf(SPHERE, [curvature], mat=MAT_GLASS_CONST, matp=[index],
                       world=translate(-thickness / 2 + curvature, cy, cz))
            s_b = Leaf(SPHERE, [curvature], mat=MAT_GLASS_CONST, matp=[index],
                       world=translate(thickness / 2 - curvature, cy, cz))
            inner = intersect(cyl, s_a)
            inner.aabb = tuple(tight_box(inner).reshape(6))
            lens = intersect(inner, s_b)
            lens.aabb = tuple(tight_box(lens).reshape(6))
            comps.append(lens)
    size = pitch * max(nx, ny) * 2
    comps.append(Leaf(PLANE, [size, size], world=translate(12.0, 0, 0) @ rot_y(90)))
    return build(comps), np.asarray(centres)


def lenslet_rays(centres, per_lens, seed=0, spread=0.8):
    """Rays from x = -5 aimed at random points of every lenslet, slightly tilted."""
    rng = np.random.default_rng(seed)
    c = np.repeat(centres, per_lens, axis=0)
    n = c.shape[0]
    o = np.column_stack([np.full(n, -5.0), c[:, 0] + rng.uniform(-spread, spread, n), c[:, 1] + rng.uniform(-spread, spread, n)])
    d = np.column_stack([np.ones(n), rng.normal(0, 0.05, n), rng.normal(0, 0.05, n)])
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return make_rays(o, d)
