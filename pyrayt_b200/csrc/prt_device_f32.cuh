// Per-ray device functions of the optional FP32 fast mode (PRT_FLAG_FP32, include/pyrayt_b200.h).
//
// The same path as prt_device.cuh -- world->object transform, primitive intersect (primitives.py:241-741),
// CSG merge (csg.py:13-160: closed form for the left-deep trees the reference's factories build, the streaming
// interpreter of prt_device.cuh for any other tree), nearest hit
// (_pyrayt.py:370-392), normals and material interaction (world_objects.py:401-418, operations.py:86-162,
// materials.py:47-145) -- in single precision, with FMA contraction and the fast division / square root.
// Contract: the frame agrees with the FP64 frame to the tolerance stated in DESIGN.md (1e-5 of the scene
// scale for positions, 1e-5 absolute for unit directions); surface / generation ids agree except for rays
// that pass within that distance of an edge, which bench.py counts.
//
// What cannot carry over from the FP64 path is the reference's way of leaving a surface: it restarts a ray
// 1e-6 beyond the point it hit (_pyrayt.py:190,:449), which single precision cannot even represent at
// |x| ~ 100 (ulp 7.6e-6), and the hit point itself is only known to ~1e-5.  The fast mode therefore names
// the leaf a ray has just interacted with (`self`): a root of THAT leaf closer than kSelfEps x the ray's
// scale is the crossing the ray has just made and is forced to -eps ("just behind the ray"), which is where
// exact arithmetic would have put it.  Index parity in the hit lists -- what array_csg decides from -- is
// untouched.
#pragma once
#include <math.h>
#include <stdint.h>

#include "prt_device.cuh"

namespace prt {
namespace f32 {

#define PRT_INFF (__builtin_huge_valf())

constexpr float kSelfEps = 2e-4f;    // x max(1, |origin|_inf): roots of the leaf just left that are the crossing itself
constexpr float kCullMarginF = 1e-3f;  // x max(1, |origin|_inf): slack of the box pruning (FP64 path: 1e-7 absolute)

// scene records in single precision, converted once per block from the staged FP64 blob
struct LeafF {
  float m[12];
  float prm[6];
  float matp[6];
  float nscale;
  int type, mat, comp;
};
struct CompF {
  float root_box[6];
  float inner_box[6];
};

// OrderEntry (prt_scene.h) in single precision: near faces rounded down, far faces up
struct OrderEntryF {
  float near_u, far_u, pmfar_u;
  int comp;
};

struct SceneViewF {
  const BlobHeader* h;
  const Comp* comps;   // shape, flags, leaves, truth table (integers) from the FP64 blob
  const Leaf* leaves;  // FP64 leaves (surface ids)
  const LeafF* lf;
  const CompF* cf;
  const OrderEntryF* order;  // [6][n_boxed] ray-ordered traversal tables, or nullptr (list-order walk)
  const int* unboxed;        // [n_unboxed]
  const Op* ops;             // preorder programs of the components that need the interpreter
  const float* aabb;         // [n_aabb][6] their node boxes, rounded outwards (nullptr: no such component)
};

typedef HitStackT<float> HitStackF;

// double -> float, rounded down / up (boxes are rounded outwards so that a single-precision box still
// contains the FP64 one)
PRT_HD float to_float_dn(double x) {
#if defined(__CUDA_ARCH__)
  return __double2float_rd(x);
#else
  const float f = (float)x;
  return ((double)f > x) ? nextafterf(f, -PRT_INFF) : f;
#endif
}
PRT_HD float to_float_up(double x) {
#if defined(__CUDA_ARCH__)
  return __double2float_ru(x);
#else
  const float f = (float)x;
  return ((double)f < x) ? nextafterf(f, PRT_INFF) : f;
#endif
}
PRT_HD void convert_leaf(const Leaf& L, LeafF& F) {
  for (int k = 0; k < 12; ++k) F.m[k] = (float)L.m[k];
  for (int k = 0; k < 6; ++k) {
    F.prm[k] = (float)L.prm[k];
    F.matp[k] = (float)L.matp[k];
  }
  F.nscale = (float)L.nscale;
  F.type = L.type;
  F.mat = L.mat;
  F.comp = L.comp;
}
PRT_HD void convert_order(const OrderEntry& E, OrderEntryF& F) {
  F.near_u = to_float_dn(E.near_u);
  F.far_u = to_float_up(E.far_u);
  F.pmfar_u = to_float_up(E.pmfar_u);
  F.comp = E.comp;
}
PRT_HD void convert_comp(const Comp& C, CompF& F) {
  for (int k = 0; k < 6; ++k) {
    F.root_box[k] = (k & 1) ? to_float_up(C.root_box[k]) : to_float_dn(C.root_box[k]);
    F.inner_box[k] = (k & 1) ? to_float_up(C.inner_box[k]) : to_float_dn(C.inner_box[k]);
  }
}

PRT_HD bool isz(float x) { return fabsf(x) <= 1e-8f; }
PRT_HD void sort2(float& a, float& b) {
  const float lo = fminf(a, b), hi = fmaxf(a, b);
  // fmin / fmax drop NaNs; hit parameters here are never NaN unless the ray is (dead anyway)
  a = lo;
  b = hi;
}
PRT_HD float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdividef(a, b);
#else
  return a / b;
#endif
}
PRT_HD float frcp(float b) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));  // one MUFU op, ~1 ulp
  return r;
#else
  return 1.0f / b;
#endif
}
PRT_HD float fsqrt(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));  // MUFU.RSQ x x, ~1 ulp
  return r;
#else
  return sqrtf(x);
#endif
}

// roots of a t^2 + b t + c (a > 0, disc >= 0) in the cancellation-free form: q = -(b + sign(b) sqrt(disc)) / 2,
// roots q / a and c / q
PRT_HD void quad_roots(float a, float b, float c, float disc, float& t0, float& t1) {
  const float root = fsqrt(disc);
  const float q = -0.5f * (b + copysignf(root, b));
  const float r0 = fdiv(q, a);
  const float r1 = (q != 0.0f) ? fdiv(c, q) : r0;
  t0 = fminf(r0, r1);
  t1 = fmaxf(r0, r1);
}

// the z-slab clip of Cylinder (primitives.py:680-712) and Paraboloid (:369-399)
PRT_HD void clip_z(float s0, float s1, float zlo, float zhi, float oz, float dz, float& t0, float& t1) {
  float b0, b1;
  if (isz(dz)) {
    b0 = ((oz >= zlo) && (oz <= zhi)) ? -PRT_INFF : PRT_INFF;
    b1 = PRT_INFF;
  } else {
    const float r = frcp(dz);
    b0 = (zlo - oz) * r;
    b1 = (zhi - oz) * r;
  }
  sort2(b0, b1);
  const float lo = fmaxf(s0, b0), hi = fminf(s1, b1);
  const bool hit = lo <= hi;
  t0 = hit ? lo : PRT_INFF;
  t1 = hit ? hi : PRT_INFF;
}

PRT_HD void slab(float o, float d, float lo, float hi, float& mn, float& mx) {
  if (isz(d)) {
    mn = (o <= hi && o >= lo) ? -PRT_INFF : PRT_INFF;
    mx = PRT_INFF;
    return;
  }
  const float r = frcp(d);
  float h0 = (lo - o) * r, h1 = (hi - o) * r;
  sort2(h0, h1);
  mn = h0;
  mx = h1;
}

// Cube.intersect (primitives.py:516-581) / a world-space bounding box (csg.py:126-128)
PRT_HD void cube_hits(const float* sp, float o0, float o1, float o2, float d0, float d1, float d2, float& t0,
                      float& t1) {
  float mn0, mx0, mn1, mx1, mn2, mx2;
  slab(o0, d0, sp[0], sp[1], mn0, mx0);
  slab(o1, d1, sp[2], sp[3], mn1, mx1);
  slab(o2, d2, sp[4], sp[5], mn2, mx2);
  const float lo = fmaxf(fmaxf(mn0, mn1), mn2), hi = fminf(fminf(mx0, mx1), mx2);
  const bool hit = lo < hi;
  t0 = hit ? lo : PRT_INFF;
  t1 = hit ? hi : PRT_INFF;
}

// a box test with precomputed reciprocals of the world direction (every component of one generation)
struct RayInvF {
  float r0, r1, r2;
  bool ok;  // no direction component is (close to) zero: the reciprocal form is usable
};
PRT_HD void box_hits(const float* sp, float o0, float o1, float o2, float d0, float d1, float d2, const RayInvF& I,
                     float& t0, float& t1) {
  if (!I.ok) {
    cube_hits(sp, o0, o1, o2, d0, d1, d2, t0, t1);
    return;
  }
  const float a0 = (sp[0] - o0) * I.r0, b0 = (sp[1] - o0) * I.r0;
  const float a1 = (sp[2] - o1) * I.r1, b1 = (sp[3] - o1) * I.r1;
  const float a2 = (sp[4] - o2) * I.r2, b2 = (sp[5] - o2) * I.r2;
  const float lo = fmaxf(fmaxf(fminf(a0, b0), fminf(a1, b1)), fminf(a2, b2));
  const float hi = fminf(fminf(fmaxf(a0, b0), fmaxf(a1, b1)), fmaxf(a2, b2));
  const bool hit = lo < hi;
  t0 = hit ? lo : PRT_INFF;
  t1 = hit ? hi : PRT_INFF;
}

// TracerSurface.intersect (world_objects.py:360-383).  `self_eps` > 0: this is the leaf the ray has just
// interacted with; roots within self_eps of the origin are the crossing it has just made (see the header).
PRT_HD void leaf_hits(const LeafF& L, float p0, float p1, float p2, float v0, float v1, float v2, float self_eps,
                      float& t0, float& t1) {
  const float o0 = L.m[0] * p0 + L.m[1] * p1 + L.m[2] * p2 + L.m[3];
  const float o1 = L.m[4] * p0 + L.m[5] * p1 + L.m[6] * p2 + L.m[7];
  const float o2 = L.m[8] * p0 + L.m[9] * p1 + L.m[10] * p2 + L.m[11];
  const float d0 = L.m[0] * v0 + L.m[1] * v1 + L.m[2] * v2;
  const float d1 = L.m[4] * v0 + L.m[5] * v1 + L.m[6] * v2;
  const float d2 = L.m[8] * v0 + L.m[9] * v1 + L.m[10] * v2;
  switch (L.type) {
    case PRT_SPHERE: {  // primitives.py:241-271
      const float r = L.prm[0];
      const float a = d0 * d0 + d1 * d1 + d2 * d2;
      const float b = 2 * (d0 * o0 + d1 * o1 + d2 * o2);
      const float c = (o0 * o0 + o1 * o1 + o2 * o2) - r * r;
      const float disc = b * b - 4 * a * c;
      quad_roots(a, b, c, fmaxf(disc, 0.0f), t0, t1);
      if (!(disc >= 0)) {
        t0 = PRT_INFF;
        t1 = PRT_INFF;
      }
    } break;
    case PRT_CYLINDER: {  // primitives.py:650-712 + operations.py:28-63
      const float r = L.prm[0];
      const float a = d0 * d0 + d1 * d1;
      const float b = 2 * (d0 * o0 + d1 * o1);
      const float c = (o0 * o0 + o1 * o1) - r * r;
      const float disc = b * b - 4 * a * c;
      float s0, s1;
      if (isz(a)) {
        // The ray runs along the axis (|d_xy| <= 1e-4).  binomial_root (operations.py:45-52) then solves the
        // linear equation unless b ~ 0 too, in which case the ray is inside or outside for good.  In single
        // precision an exactly parallel ray (a collimated beam returning along a parabolic mirror's axis)
        // carries rounding noise of a few ulp in d_xy, so "b ~ 0" allows for 16 ulp of it.
        if (fabsf(b) <= 1e-8f + 2e-6f * (fabsf(o0) + fabsf(o1))) {
          s0 = (c <= 0) ? -PRT_INFF : PRT_INFF;
          s1 = PRT_INFF;
        } else {
          s0 = s1 = -c / b;
        }
      } else {
        quad_roots(a, b, c, fmaxf(disc, 0.0f), s0, s1);
        if (!(disc >= 0)) {
          s0 = PRT_INFF;
          s1 = PRT_INFF;
        }
      }
      clip_z(s0, s1, L.prm[1], L.prm[2], o2, d2, t0, t1);
    } break;
    case PRT_PARABOLOID: {  // primitives.py:320-399
      const float f = L.prm[0];
      const float a = d0 * d0 + d1 * d1;
      const float b = 2 * (o0 * d0 + o1 * d1) - 4 * f * d2;
      const float c = (o0 * o0 + o1 * o1) - 4 * f * o2;
      const float disc = b * b - 4 * a * c;
      float s0, s1;
      if (isz(a)) {
        s0 = -c / (b + (isz(b) ? 1.0f : 0.0f));
        s1 = (d2 >= 0) ? PRT_INFF : -PRT_INFF;
        sort2(s0, s1);
      } else {
        quad_roots(a, b, c, fmaxf(disc, 0.0f), s0, s1);
        if (!(disc >= 0)) {
          s0 = PRT_INFF;
          s1 = PRT_INFF;
        }
      }
      clip_z(s0, s1, 0.0f, L.prm[1], o2, d2, t0, t1);
    } break;
    case PRT_PLANE: {  // primitives.py:436-492
      float xmn, xmx, ymn, ymx;
      slab(o0, d0, -0.5f * L.prm[0], 0.5f * L.prm[0], xmn, xmx);
      slab(o1, d1, -0.5f * L.prm[1], 0.5f * L.prm[1], ymn, ymx);
      const float lo = fmaxf(xmn, ymn), hi = fminf(xmx, ymx);
      float t = isz(d2) ? PRT_INFF : fdiv(-o2, d2);
      if (!((t >= lo) & (t <= hi))) t = PRT_INFF;
      t0 = t;
      t1 = t;
    } break;
    case PRT_CUBE:  // primitives.py:516-581
      cube_hits(L.prm, o0, o1, o2, d0, d1, d2, t0, t1);
      break;
    default:
      t0 = PRT_INFF;
      t1 = PRT_INFF;
  }
  if (self_eps > 0.0f) {
    if (fabsf(t0) < self_eps) t0 = -self_eps;
    if (fabsf(t1) < self_eps) t1 = -self_eps;
    sort2(t0, t1);
  }
}

// "is this coordinate on that face": np.isclose's 1e-8 + 1e-5 |v| (primitives.py:408,:594-599,:726-733) plus
// `slack`, the single-precision uncertainty of the hit point itself
PRT_HD bool on_face(float q, float v, float slack) { return fabsf(q - v) <= (1e-8f + 1e-5f * fabsf(v) + slack); }

// TracerSurface.get_world_normals (world_objects.py:401-418) with the primitives' normal().  `slack`: how far
// the hit point may be off the surface (a few ulp of the distances involved; the caller knows them).
PRT_HD void world_normal(const LeafF& L, float p0, float p1, float p2, float slack, float& n0, float& n1,
                         float& n2) {
  const float q0 = L.m[0] * p0 + L.m[1] * p1 + L.m[2] * p2 + L.m[3];
  const float q1 = L.m[4] * p0 + L.m[5] * p1 + L.m[6] * p2 + L.m[7];
  const float q2 = L.m[8] * p0 + L.m[9] * p1 + L.m[10] * p2 + L.m[11];
  float a0, a1, a2;
  switch (L.type) {
    case PRT_SPHERE:
      a0 = q0;
      a1 = q1;
      a2 = q2;
      break;
    case PRT_PARABOLOID:
      if (on_face(q2, L.prm[1], slack)) {
        a0 = 0;
        a1 = 0;
        a2 = 1;
      } else {
        a0 = q0;
        a1 = q1;
        a2 = -2 * L.prm[0];
      }
      break;
    case PRT_PLANE:
      a0 = 0;
      a1 = 0;
      a2 = 1;
      break;
    case PRT_CUBE:
      a0 = on_face(q0, L.prm[1], slack) ? 1.0f : (on_face(q0, L.prm[0], slack) ? -1.0f : 0.0f);
      a1 = on_face(q1, L.prm[3], slack) ? 1.0f : (on_face(q1, L.prm[2], slack) ? -1.0f : 0.0f);
      a2 = on_face(q2, L.prm[5], slack) ? 1.0f : (on_face(q2, L.prm[4], slack) ? -1.0f : 0.0f);
      break;
    default:  // PRT_CYLINDER
      a0 = q0;
      a1 = q1;
      a2 = 0;
      if (L.prm[3] != 0.0f) {
        if (on_face(q2, L.prm[1], slack)) {
          a0 = 0;
          a1 = 0;
          a2 = -1;
        }
        if (on_face(q2, L.prm[2], slack)) {
          a0 = 0;
          a1 = 0;
          a2 = 1;
        }
      }
      break;
  }
  // M_obj^T n_obj, normalise, flip (world_objects.py:411-418); normalising the object normal first
  // (as the reference does) only rescales the vector that is normalised here
  const float w0 = L.m[0] * a0 + L.m[4] * a1 + L.m[8] * a2;
  const float w1 = L.m[1] * a0 + L.m[5] * a1 + L.m[9] * a2;
  const float w2 = L.m[2] * a0 + L.m[6] * a1 + L.m[10] * a2;
  const float s = L.nscale * frcp(fsqrt(w0 * w0 + w1 * w1 + w2 * w2));
  n0 = w0 * s;
  n1 = w1 * s;
  n2 = w2 * s;
}

struct RayStateF {
  float p0, p1, p2, v0, v1, v2, wl, nidx;
  int skip;  // convex component the ray has just left for good, or -1 (as in the FP64 path)
  int self;  // leaf the ray has just interacted with, or -1
};

// component.intersect for one ray by the preorder interpreter (eval_component of prt_device.cuh in single
// precision): any CSG tree.  Returns false when the root box is missed or, for proven boxes, lies behind the ray or
// beyond the best hit so far.
PRT_HD bool eval_component(const SceneViewF& sc, int begin, int end, const RayStateF& r, const RayInvF& inv,
                           float margin, float self_eps, float best_t, HitStackF& S, bool& tie) {
  const float p0 = r.p0, p1 = r.p1, p2 = r.p2, v0 = r.v0, v1 = r.v1, v2 = r.v2;
  int sp = 0;
  int pc = begin;
  while (pc < end) {
    const Op op = sc.ops[pc];
    if (op.kind == OP_ENTER) {
      float b0, b1;
      box_hits(sc.aabb + 6 * op.a, p0, p1, p2, v0, v1, v2, inv, b0, b1);
      if (!(b0 < PRT_INFF)) {  // csg.py:126-133
        if (pc == begin) return false;
        S.len[sp++] = 0;
        pc = op.b;
        continue;
      }
      if (pc == begin && (op.c & 1) && (b1 < -margin || b0 > best_t + margin)) return false;
    } else if (op.kind == OP_LEAF) {
      float t0, t1;
      leaf_hits(sc.lf[op.a], p0, p1, p2, v0, v1, v2, (op.a == r.self) ? self_eps : 0.0f, t0, t1);
      const int b = buf_of(S, sp);
      S.t[b][0] = t0;
      S.t[b][1] = t1;
      S.leaf[b][0] = (unsigned short)op.a;
      S.leaf[b][1] = (unsigned short)op.a;
      S.len[sp++] = (t0 < PRT_INFF) ? ((t1 < PRT_INFF) ? 2 : 1) : 0;
    } else if (op.kind == OP_MERGE_LEAF) {
      float t0, t1;
      leaf_hits(sc.lf[op.b], p0, p1, p2, v0, v1, v2, (op.b == r.self) ? self_eps : 0.0f, t0, t1);
      const int n = (t0 < PRT_INFF) ? ((t1 < PRT_INFF) ? 2 : 1) : 0;
      const int lf = op.b;
      merge_lists(
          S, sp - 1, op.a, n, [&](int j) { return j ? t1 : t0; }, [&](int) { return lf; }, tie);
    } else {  // OP_MERGE
      const int rl = sp - 1;
      const int rb = buf_of(S, rl);
      merge_lists(
          S, sp - 2, op.a, S.len[rl], [&](int j) { return S.t[rb][j]; }, [&](int j) { return (int)S.leaf[rb][j]; },
          tie);
      --sp;
    }
    ++pc;
  }
  return true;
}

// one component for nearest_hit: its first positive kept entry (ct, cl), or ct = +inf.  Components with a proven /
// conservative box (Comp.flags & 5) are skipped when the box lies behind the ray or beyond the best hit so far;
// a CSG component whose root box the ray misses has no hits at all (csg.py:126-133).
// GENERIC = false compiles the interpreter out (scenes of bare surfaces and left-deep trees).
template <bool GENERIC>
PRT_HD void eval_comp(const SceneViewF& sc, int c, const RayStateF& r, const RayInvF& inv, float margin,
                      float self_eps, float best_t, HitStackF* S, float& ct, int& cl, bool& tie) {
  const float p0 = r.p0, p1 = r.p1, p2 = r.p2, v0 = r.v0, v1 = r.v1, v2 = r.v2;
  const Comp& C = sc.comps[c];
  const CompF& F = sc.cf[c];
  const int shape = C.shape;
  ct = PRT_INFF;
  cl = -1;
  if (shape == SHAPE_LEAF) {
    if (C.flags & 4) {  // conservative world box of the bare surface: prune only
      float b0, b1;
      box_hits(F.root_box, p0, p1, p2, v0, v1, v2, inv, b0, b1);
      if (!(b0 < PRT_INFF) || (b1 < -margin) || (b0 > best_t + margin)) return;
    }
    float t0, t1;
    leaf_hits(sc.lf[C.leaf_a], p0, p1, p2, v0, v1, v2, (C.leaf_a == r.self) ? self_eps : 0.0f, t0, t1);
    ct = (t0 > 0) ? t0 : ((t1 > 0) ? t1 : PRT_INFF);
    cl = C.leaf_a;
    return;
  }
  if (shape == SHAPE_GENERIC) {
    if (GENERIC) {
      S->flags = 0;
      if (eval_component(sc, C.begin, C.end, r, inv, margin, self_eps, best_t, *S, tie)) {
        const int b = buf_of(*S, 0);
        const int n = S->len[0];
        for (int q = 0; q < n; ++q) {  // sorted: the first positive entry is the argmin of where(hits > 0)
          const float t = S->t[b][q];
          if (t > 0) {
            ct = t;
            cl = S->leaf[b][q];
            break;
          }
        }
      }
    }
    return;
  }
  // SHAPE_LEFT2 / SHAPE_LEFT3
  float b0, b1;
  box_hits(F.root_box, p0, p1, p2, v0, v1, v2, inv, b0, b1);
  if (!(b0 < PRT_INFF)) return;
  if ((C.flags & 1) && ((b1 < -margin) || (b0 > best_t + margin))) return;
  bool inner_hit = true;
  if (shape == SHAPE_LEFT3) {
    box_hits(F.inner_box, p0, p1, p2, v0, v1, v2, inv, b0, b1);
    inner_hit = b0 < PRT_INFF;
  }
  const int la = C.leaf_a, lb = C.leaf_b, lc = C.leaf_c;
  float a0 = PRT_INFF, a1 = PRT_INFF, q0 = PRT_INFF, q1 = PRT_INFF, c0 = PRT_INFF, c1 = PRT_INFF;
  if (inner_hit) {
    leaf_hits(sc.lf[la], p0, p1, p2, v0, v1, v2, (la == r.self) ? self_eps : 0.0f, a0, a1);
    leaf_hits(sc.lf[lb], p0, p1, p2, v0, v1, v2, (lb == r.self) ? self_eps : 0.0f, q0, q1);
  }
  if (shape == SHAPE_LEFT3) leaf_hits(sc.lf[lc], p0, p1, p2, v0, v1, v2, (lc == r.self) ? self_eps : 0.0f, c0, c1);
  left_deep_first_hit((unsigned)C.tt, a0, a1, q0, q1, c0, c1, la, lb, lc, ct, cl, tie);
}

// nearest hit over all components (_pyrayt.py:376-386): smallest distance, earliest component on ties -- a
// result that does not depend on the visiting order.  With traversal tables (sc.order) the boxed components are
// visited in the order the ray meets them along its dominant axis, as in the FP64 path (prt_device.cuh,
// nearest_hit): everything behind the ray is skipped by bisection and the walk stops at the first box that
// begins beyond the best hit.  Otherwise, and for rays the threshold tests cannot serve, list order.
// ORDERED = false (the kernel variant for scenes the encoder leaves in list order): a plain loop over the
// components, nothing else compiled in.
template <bool ORDERED, bool GENERIC>
PRT_HD void nearest_hit(const SceneViewF& sc, const RayStateF& r, float scale, HitStackF* S, float& best_t,
                        int& best_leaf, bool& tie) {
  best_t = PRT_INFF;
  best_leaf = -1;
  const float v0 = r.v0, v1 = r.v1, v2 = r.v2;
  RayInvF inv;
  inv.ok = !(isz(v0) | isz(v1) | isz(v2));
  inv.r0 = frcp(v0);
  inv.r1 = frcp(v1);
  inv.r2 = frcp(v2);
  const float margin = kCullMarginF * scale;
  const float self_eps = kSelfEps * scale;
  const int nc = sc.h->n_components;
  if (!ORDERED) {
    for (int c = 0; c < nc; ++c) {
      if (c == r.skip) continue;
      float ct;
      int cl;
      eval_comp<GENERIC>(sc, c, r, inv, margin, self_eps, best_t, S, ct, cl, tie);
      if (ct < best_t) {  // strict: the earlier component keeps a tie (_pyrayt.py:384)
        best_t = ct;
        best_leaf = cl;
      }
    }
    return;
  }
  int best_comp = -1;
  // dominant axis (DomAxis of the FP64 path)
  const float a0 = fabsf(v0), a1 = fabsf(v1), a2 = fabsf(v2);
  const int k = (a0 >= a1) ? ((a0 >= a2) ? 0 : 2) : ((a1 >= a2) ? 1 : 2);
  const float a = (k == 0) ? a0 : ((k == 1) ? a1 : a2);
  const float o = (k == 0) ? r.p0 : ((k == 1) ? r.p1 : r.p2);
  const float vk = (k == 0) ? v0 : ((k == 1) ? v1 : v2);
  const int sgn = vk < 0 ? 1 : 0;
  const bool ordered = (sc.order != nullptr) & (fabsf(o) <= 1e6f) & (a >= 0.25f) & (a <= 4.0f);
  // one loop, one call site of eval_comp: phase 0 walks the ray-ordered table, phase 1 the components without
  // a box, phase 2 (instead of both) every component in list order
  const int n0 = ordered ? sc.h->n_boxed : 0, n1 = ordered ? sc.h->n_unboxed : nc;
  const OrderEntryF* tab = ordered ? sc.order + (2 * k + sgn) * n0 : nullptr;
  const float thr_far = (sgn ? -o : o) - 3 * margin * a;  // boxes ending before it lie behind the ray
  float thr_near = PRT_INFF;                               // boxes beginning after it lie beyond the best hit
  int j = 0;
  if (ordered) {
    int hi = n0;
    while (j < hi) {
      const int mid = (j + hi) >> 1;
      if (tab[mid].pmfar_u < thr_far) j = mid + 1; else hi = mid;
    }
  }
  int phase = (j < n0) ? 0 : 1;
  if (phase) j = 0;
  for (;;) {
    int c;
    if (phase == 0) {
      const OrderEntryF e = tab[j];
      const bool beyond = e.near_u > thr_near;  // ray order: every later box begins even further away
      ++j;
      if (beyond | (j >= n0)) {
        phase = 1;
        j = 0;
      }
      if (beyond | (e.far_u < thr_far)) continue;
      c = e.comp;
    } else {
      if (j >= n1) break;
      c = ordered ? sc.unboxed[j] : j;
      ++j;
    }
    if (c == r.skip) continue;
    float ct;
    int cl;
    eval_comp<GENERIC>(sc, c, r, inv, margin, self_eps, best_t, S, ct, cl, tie);
    // smallest distance, then earliest component (== the reference's in-order strict `<`, _pyrayt.py:384)
    if ((ct < best_t) | ((ct == best_t) & (ct < PRT_INFF) & (c < best_comp))) {
      best_t = ct;
      best_leaf = cl;
      best_comp = c;
      thr_near = thr_far + (ct + 5 * margin) * a;
    }
  }
}

struct StepOutF {
  float e0, e1, e2;     // hit point
  float nv0, nv1, nv2;  // direction after the interaction
  float n_next;
  int skip;
};

// _st_interact (_pyrayt.py:394-452) for a ray whose nearest hit is (best_t, best_leaf >= 0); counters as in
// the FP64 step_interact.  Returns true when the ray goes on.
PRT_HD bool step_interact(const SceneViewF& sc, const RayStateF& r, int g, int generation_limit, float vn,
                          float best_t, int best_leaf, StepOutF& o, StepCounters& c) {
  const LeafF& L = sc.lf[best_leaf];
  o.e0 = r.p0 + r.v0 * best_t;
  o.e1 = r.p1 + r.v1 * best_t;
  o.e2 = r.p2 + r.v2 * best_t;
  o.n_next = r.nidx;
  o.skip = -1;
  bool goes_on = true;
  float n0 = 0, n1 = 0, n2 = 0;
  if (L.mat == PRT_MAT_MIRROR || L.mat == PRT_MAT_GLASS_CONST || L.mat == PRT_MAT_GLASS_SELLMEIER)
    world_normal(L, o.e0, o.e1, o.e2,
                 2e-6f * fmaxf(fmaxf(1.0f, best_t * vn), fmaxf(fabsf(o.e0), fmaxf(fabsf(o.e1), fabsf(o.e2)))), n0, n1,
                 n2);
  if (L.mat == PRT_MAT_ABSORBER) {
    o.nv0 = 0;
    o.nv1 = 0;
    o.nv2 = 0;
    goes_on = false;
    c.w1 |= kCtrAbs;
  } else if (L.mat == PRT_MAT_MIRROR) {  // operations.py:105-107
    c.w1 += 1u;
    const float dots = r.v0 * n0 + r.v1 * n1 + r.v2 * n2;
    o.nv0 = r.v0 - 2 * n0 * dots;
    o.nv1 = r.v1 - 2 * n1 * dots;
    o.nv2 = r.v2 - 2 * n2 * dots;
    const float out_dot = o.nv0 * n0 + o.nv1 * n1 + o.nv2 * n2;
    o.skip = ((L.comp >= 0) && (out_dot > 1e-3f) && (sc.comps[L.comp].flags & 2)) ? L.comp : -1;
  } else if (L.mat == PRT_MAT_GLASS_CONST || L.mat == PRT_MAT_GLASS_SELLMEIER) {  // operations.py:110-162
    const float rv = frcp(vn);
    const float u0 = r.v0 * rv, u1 = r.v1 * rv, u2 = r.v2 * rv;
    const float cp = u0 * n0 + u1 * n1 + u2 * n2;
    const bool exiting = cp > 0;
    float n2l = 1.0f;
    if (!exiting) {
      if (L.mat == PRT_MAT_GLASS_CONST) {
        n2l = L.matp[0];
      } else {
        const float w2 = r.wl * r.wl;
        n2l = fsqrt(1 + fdiv(L.matp[0] * w2, w2 - L.matp[3]) + fdiv(L.matp[1] * w2, w2 - L.matp[4]) +
                    fdiv(L.matp[2] * w2, w2 - L.matp[5]));
      }
    } else {
      n0 = -n0;
      n1 = -n1;
      n2 = -n2;
    }
    const float q = fdiv(r.nidx, n2l);
    const float c1 = exiting ? cp : -cp;
    const float rad = 1 - (q * q) * (1 - c1 * c1);
    if (rad > 0) {
      const float k = q * c1 - fsqrt(rad);
      o.nv0 = q * u0 + k * n0;
      o.nv1 = q * u1 + k * n1;
      o.nv2 = q * u2 + k * n2;
      o.n_next = n2l;
    } else {
      const float k = 2 * c1;
      o.nv0 = u0 + k * n0;
      o.nv1 = u1 + k * n1;
      o.nv2 = u2 + k * n2;
    }
    const float s = frcp(fsqrt(o.nv0 * o.nv0 + o.nv1 * o.nv1 + o.nv2 * o.nv2));
    o.nv0 *= s;
    o.nv1 *= s;
    o.nv2 *= s;
    const float out_dot = exiting ? -(o.nv0 * n0 + o.nv1 * n1 + o.nv2 * n2) : -1.0f;
    o.skip = ((L.comp >= 0) && (out_dot > 1e-3f) && (sc.comps[L.comp].flags & 2)) ? L.comp : -1;
  } else {
    c.w1 |= kCtrUntr;
    return false;
  }
  c.w0 += 1u << 16;
  if (g + 1 == generation_limit) {
    c.w1 |= kCtrLim;
    return false;
  }
  return goes_on;
}

PRT_HD void advance_ray(RayStateF& r, const StepOutF& o, int hit_leaf, float ray_offset) {
  r.skip = o.skip;
  r.self = hit_leaf;
  r.nidx = o.n_next;
  r.v0 = o.nv0;
  r.v1 = o.nv1;
  r.v2 = o.nv2;
  r.p0 = o.e0 + ray_offset * o.nv0;
  r.p1 = o.e1 + ray_offset * o.nv1;
  r.p2 = o.e2 + ray_offset * o.nv2;
}

PRT_HD float ray_scale(const RayStateF& r) {
  return fmaxf(1.0f, fmaxf(fabsf(r.p0), fmaxf(fabsf(r.p1), fabsf(r.p2))));
}

// dead / NaN test and generation count of step_speed (prt_device.cuh)
PRT_HD float step_speed(const RayStateF& r, StepCounters& c) {
  const float vn = fsqrt(r.v0 * r.v0 + r.v1 * r.v1 + r.v2 * r.v2);
  if (isz(vn)) return 0.0f;
  if (isnan(r.v0) | isnan(r.v1) | isnan(r.v2)) {
    c.w1 |= kCtrNan;
    return 0.0f;
  }
  c.w0 += 1u;
  return vn;
}

}  // namespace f32
}  // namespace prt
