// Per-ray device functions of the trace kernels (sm_100a).
//
// Everything here is PRT_HD (__host__ __device__) so that tests/emul can run the very
// same per-ray code on the host for debugging; the product library only ever calls
// these from kernels.  Arithmetic is IEEE float64 without FMA contraction (-fmad=false)
// in the order of the reference's NumPy expressions.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/pyrayt_b200.h"
#include "prt_scene.h"

#if defined(__CUDACC__)
#define PRT_HD __host__ __device__ __forceinline__
#define PRT_HD_CALL static __host__ __device__ __noinline__  // one copy in the kernel: keeps the hot loop inside the I-cache
#else
#define PRT_HD inline
#define PRT_HD_CALL inline
#endif

namespace prt {
PRT_HD int popc32(unsigned x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
}  // namespace prt

namespace prt {

#define PRT_INF (__builtin_huge_val())

// np.isclose(x, 0): |x| <= 1e-8  (NaN / inf -> false)
PRT_HD bool isz(double x) { return fabs(x) <= 1e-8; }
// np.isclose(p, v): |p - v| <= 1e-8 + 1e-5 |v|
PRT_HD bool iscl(double p, double v) {
  // equal infinities compare close in NumPy; inf - inf = NaN fails the first test, p == v catches it
  return (fabs(p - v) <= (1e-8 + 1e-5 * fabs(v))) | (p == v);
}
PRT_HD void sort2(double& a, double& b) {
  if (b < a) {
    double t = a;
    a = b;
    b = t;
  }
}

// ---------------------------------------------------------------- exact division by a shared reciprocal
//
// Many quotients of one ray share a denominator (the two roots of a quadratic, the two faces
// of a slab, every bounding box of the scene against the same world-space direction).  With
// r = RN(1/b) the sequence q = RN(a*r); e = fma(-b, q, a); q' = fma(e, r, q) yields RN(a/b)
// (Markstein's theorem), i.e. bit-for-bit what the reference's true division gives, for 3 FP64
// instructions per quotient instead of a full IEEE division.  Denominators outside a safe range
// (and numerators that could make the residual inexact) fall back to a real division.
struct Rcp {
  double b, r;
  unsigned lim;  // width of the numerator-exponent window in which the 3-instruction form is exact (0: never)
};

PRT_HD unsigned hi_word(double x) {
#if defined(__CUDA_ARCH__)
  return (unsigned)__double2hiint(x);
#else
  unsigned long long u;
  memcpy(&u, &x, 8);
  return (unsigned)(u >> 32);
#endif
}
// biased exponent of x (0 for zero/denormals, 2047 for inf/NaN)
PRT_HD unsigned exp_of(double x) { return (hi_word(x) >> 20) & 0x7ffu; }

constexpr unsigned kExpLo = 123;          // 2^-900
constexpr unsigned kExpSpan = 1923 - 123;  // up to 2^900

// The same window tested on several values at once: the exponent field in place (one mask per value), then
// integer minimum / maximum (three-input VIMNMX3 on sm_100) and two comparisons for the whole group,
// instead of a shift, a mask, a subtraction and a comparison per value.
PRT_HD unsigned mag_of(double x) { return hi_word(x) & 0x7ff00000u; }
PRT_HD unsigned umin2(unsigned a, unsigned b) { return a < b ? a : b; }
PRT_HD unsigned umax2(unsigned a, unsigned b) { return a > b ? a : b; }
PRT_HD bool mags_in_window(unsigned mn, unsigned mx) {
  return (mn >= (kExpLo << 20)) & (mx < ((kExpLo + kExpSpan) << 20));
}


// the reciprocal alone: the caller tests the denominator's window together with its numerators (mags_in_window)
PRT_HD Rcp make_rcp_unchecked(double b) {
  Rcp R;
  R.b = b;
  R.lim = kExpSpan;
#if defined(__CUDA_ARCH__)
  R.r = __drcp_rn(b);
#else
  R.r = 1.0 / b;
#endif
  return R;
}
PRT_HD unsigned lim_of(double b) { return (exp_of(b) - kExpLo < kExpSpan) ? kExpSpan : 0u; }

// RN(a / R.b) without guards; exact when a is 0 or has |a| in [2^-900, 2^900) and R.lim != 0
PRT_HD double div_fast(double a, const Rcp& R) {
  const double q = a * R.r;
  const double e = fma(-R.b, q, a);
  return fma(e, R.r, q);
}

// the rare operands div_by cannot handle in 3 instructions: one shared out-of-line IEEE division
// (inlining it at ~75 call sites made a quarter of the kernel's code)
PRT_HD_CALL double slow_div(double a, double b) { return a / b; }

PRT_HD double div_by(double a, const Rcp& R) {
  if (exp_of(a) - kExpLo < R.lim) return div_fast(a, R);
  // a zero numerator over a safe denominator (the zero components of an axis-aligned normal: every cuboid
  // face, every plane): the quotient is the correctly signed zero a * r
  if ((a == 0.0) & (R.lim != 0u)) return a * R.r;
  return slow_div(a, R.b);  // tiny, huge, inf, NaN or an unsafe denominator: one shared IEEE division
}

// Two quotients by a prepared reciprocal (a slab of the rare generic box test): the 3-instruction form is computed for both
// and one test decides whether any needs the guarded path (one branch instead of one per quotient).
PRT_HD void div_by2(double a0, double a1, const Rcp& R, double& q0, double& q1) {
  q0 = div_fast(a0, R);
  q1 = div_fast(a1, R);
  const unsigned m0 = mag_of(a0), m1 = mag_of(a1);
  const bool ok = (R.lim != 0u) & mags_in_window(umin2(m0, m1), umax2(m0, m1));
  if (!ok) {
    q0 = div_by(a0, R);
    q1 = div_by(a1, R);
  }
}

// a0/b, a1/b: one window test for the denominator and the two numerators together
PRT_HD void div2_by_value(double a0, double a1, double b, double& q0, double& q1) {
  Rcp R = make_rcp_unchecked(b);
  q0 = div_fast(a0, R);
  q1 = div_fast(a1, R);
  const unsigned m0 = mag_of(a0), m1 = mag_of(a1), mb = mag_of(b);
  const bool ok = mags_in_window(umin2(umin2(m0, m1), mb), umax2(umax2(m0, m1), mb));
  if (!ok) {
    R.lim = lim_of(b);
    q0 = div_by(a0, R);
    q1 = div_by(a1, R);
  }
}

// a0/b, a1/b, a2/b: one window test for the denominator and the three numerators together
PRT_HD void div3_by_value(double a0, double a1, double a2, double b, double& q0, double& q1, double& q2) {
  Rcp R = make_rcp_unchecked(b);
  q0 = div_fast(a0, R);
  q1 = div_fast(a1, R);
  q2 = div_fast(a2, R);
  const unsigned m0 = mag_of(a0), m1 = mag_of(a1), m2 = mag_of(a2), mb = mag_of(b);
  const bool ok = mags_in_window(umin2(umin2(m0, m1), umin2(m2, mb)), umax2(umax2(m0, m1), umax2(m2, mb)));
  if (!ok) {
    R.lim = lim_of(b);
    q0 = div_by(a0, R);
    q1 = div_by(a1, R);
    q2 = div_by(a2, R);
  }
}

// per-generation reciprocals of a ray direction, shared by every bounding-box test against it
struct RayInv {
  double r0, r1, r2;  // RN(1 / (d_k + z_k))
  unsigned bits;      // 0-2: z_k (|d_k| <= 1e-8), 3-5: s_k (1 if the ray runs towards -axis: near face = hi),
                      // 6: fast (guard-free slab arithmetic is provably exact for this ray, see cube_hits)
};

struct SceneView {
  const BlobHeader* h;
  const Comp* comps;
  const Op* ops;
  const double* aabb;
  const Leaf* leaves;
  const OrderEntry* order;   // [6][n_boxed] traversal tables
  const OrderEntry* bycomp;  // [6][n_components] the same entries in list order
  const int* unboxed;        // [n_unboxed]
};

PRT_HD SceneView make_view(const unsigned char* blob) {
  SceneView s;
  s.h = reinterpret_cast<const BlobHeader*>(blob);
  s.comps = reinterpret_cast<const Comp*>(blob + s.h->off_comps);
  s.ops = reinterpret_cast<const Op*>(blob + s.h->off_ops);
  s.aabb = reinterpret_cast<const double*>(blob + s.h->off_aabb);
  s.leaves = reinterpret_cast<const Leaf*>(blob + s.h->off_leaves);
  s.order = reinterpret_cast<const OrderEntry*>(blob + s.h->off_order);
  s.unboxed = reinterpret_cast<const int*>(blob + s.h->off_unboxed);
  s.bycomp = reinterpret_cast<const OrderEntry*>(blob + s.h->off_bycomp);
  return s;
}

// ---------------------------------------------------------------- primitives (object space)

// the z-slab clip shared by Cylinder (primitives.py:680-712) and Paraboloid (:369-399)
PRT_HD void clip_z(double s0, double s1, double zlo, double zhi, double oz, double dz,
                                       double b0_num, double& t0, double& t1) {
  const bool par = isz(dz);
  const double den_b = dz + (par ? 1.0 : 0.0);
  double b0, b1;
  div2_by_value(b0_num, zhi - oz, den_b, b0, b1);
  if (par) {
    b0 = ((oz >= zlo) && (oz <= zhi)) ? -PRT_INF : PRT_INF;
    b1 = PRT_INF;
  }
  sort2(b0, b1);
  const double lo = fmax(s0, b0);
  const double hi = fmin(s1, b1);
  if (lo <= hi) {
    t0 = lo;
    t1 = hi;
  } else {
    t0 = PRT_INF;
    t1 = PRT_INF;
  }
}

// one slab of Cube.intersect (primitives.py:531-565)
PRT_HD void cube_axis(double o, bool zf, const Rcp& den, double lo, double hi, double& mn, double& mx) {
  if (zf) {  // ray parallel to the slab: (-inf, +inf) inside, (+inf, +inf) outside
    mn = (o <= hi && o >= lo) ? -PRT_INF : PRT_INF;
    mx = PRT_INF;
    return;
  }
  double h0, h1;
  div_by2(-(o - lo), -(o - hi), den, h0, h1);
  sort2(h0, h1);
  mn = h0;
  mx = h1;
}

// a coordinate / span for which o - s is 0 or at least 2^-900 in magnitude
PRT_HD bool tame(double x) { return (x == 0.0) | (exp_of(x) - 200u < 1700u - 200u); }  // 0 or [2^-823, 2^677)

// reciprocals of the direction of a ray starting at (o0,o1,o2); `boxes_tame`: every box this
// ray will be tested against has spans that are 0 or in [2^-823, 2^677) (BlobHeader.flags & 1)
PRT_HD RayInv make_ray_inv(double o0, double o1, double o2, double d0, double d1, double d2, bool boxes_tame) {
  RayInv I;
  const bool z0 = isz(d0), z1 = isz(d1), z2 = isz(d2);
  const double e0 = d0 + (z0 ? 1.0 : 0.0), e1 = d1 + (z1 ? 1.0 : 0.0), e2 = d2 + (z2 ? 1.0 : 0.0);
  const Rcp R0 = make_rcp_unchecked(e0), R1 = make_rcp_unchecked(e1), R2 = make_rcp_unchecked(e2);
  I.r0 = R0.r;
  I.r1 = R1.r;
  I.r2 = R2.r;
  // Fast slab form: differences of tame numbers are 0 or >= 2^-876 and < 2^678, denominators are in
  // [2^-66, 2^66) -> every quotient and its FMA residual stay normal, so div_fast is exact.
  const bool den_ok = (exp_of(e0) - 957u < 132u) & (exp_of(e1) - 957u < 132u) & (exp_of(e2) - 957u < 132u);
  const bool fast = boxes_tame & !z0 & !z1 & !z2 & den_ok & tame(o0) & tame(o1) & tame(o2);
  I.bits = (z0 ? 1u : 0u) | (z1 ? 2u : 0u) | (z2 ? 4u : 0u) | (e0 < 0 ? 8u : 0u) | (e1 < 0 ? 16u : 0u) |
           (e2 < 0 ? 32u : 0u) | (fast ? 64u : 0u);
  return I;
}

// the literal three-slab form (parallel rays, guarded divisions): rare, kept out of line
PRT_HD void cube_hits_generic(const double* sp, double o0, double o1, double o2, double d0, double d1, double d2,
                                   const RayInv& I, double& lo, double& hi) {
  const bool z0 = I.bits & 1u, z1 = I.bits & 2u, z2 = I.bits & 4u;
  // (rare path: the windows of the three denominators are tested here, not once per generation)
  const double e0 = d0 + (z0 ? 1.0 : 0.0), e1 = d1 + (z1 ? 1.0 : 0.0), e2 = d2 + (z2 ? 1.0 : 0.0);
  const Rcp R0 = {e0, I.r0, lim_of(e0)};
  const Rcp R1 = {e1, I.r1, lim_of(e1)};
  const Rcp R2 = {e2, I.r2, lim_of(e2)};
  double mn0, mx0, mn1, mx1, mn2, mx2;
  cube_axis(o0, z0, R0, sp[0], sp[1], mn0, mx0);
  cube_axis(o1, z1, R1, sp[2], sp[3], mn1, mx1);
  cube_axis(o2, z2, R2, sp[4], sp[5], mn2, mx2);
  lo = fmax(fmax(mn0, mn1), mn2);
  hi = fmin(fmin(mx0, mx1), mx2);
}

// Cube.intersect (primitives.py:516-581); also every CSG node's world-space AABB (csg.py:126-128).
// (d0,d1,d2) is the direction the reciprocals in I were made from.
PRT_HD void cube_hits(const double* sp, double o0, double o1, double o2, double d0, double d1, double d2,
                      const RayInv& I, double& t0, double& t1) {
  double lo, hi;
  if (I.bits & 64u) {
    // the sign of the direction says which face is the near one: no sort, no parallel-ray cases;
    // values are bit-identical to the generic path (rounding is monotonic)
    const int s0 = (I.bits >> 3) & 1, s1 = (I.bits >> 4) & 1, s2 = (I.bits >> 5) & 1;
    const Rcp R0 = {d0, I.r0, 1u}, R1 = {d1, I.r1, 1u}, R2 = {d2, I.r2, 1u};
    const double mn0 = div_fast(-(o0 - sp[s0]), R0), mx0 = div_fast(-(o0 - sp[1 - s0]), R0);
    const double mn1 = div_fast(-(o1 - sp[2 + s1]), R1), mx1 = div_fast(-(o1 - sp[3 - s1]), R1);
    const double mn2 = div_fast(-(o2 - sp[4 + s2]), R2), mx2 = div_fast(-(o2 - sp[5 - s2]), R2);
    lo = mn0 > mn1 ? mn0 : mn1;
    lo = lo > mn2 ? lo : mn2;
    hi = mx0 < mx1 ? mx0 : mx1;
    hi = hi < mx2 ? hi : mx2;
  } else {
    cube_hits_generic(sp, o0, o1, o2, d0, d1, d2, I, lo, hi);
  }
  if (lo < hi) {
    t0 = lo;
    t1 = hi;
  } else {
    t0 = PRT_INF;
    t1 = PRT_INF;
  }
}

// one bounding axis of Plane.intersect (primitives.py:454-469)
PRT_HD void plane_axis(double o, double d, double dim, double& mn, double& mx) {
  const bool zf = isz(d);
  const double half = dim / 2;
  const double den_b = d + (zf ? 1.0 : 0.0);
  double v0, v1;
  div2_by_value(-(o - half), -(o + half), den_b, v0, v1);
  if (zf) {
    v0 = (fabs(o) <= half) ? -PRT_INF : PRT_INF;
    v1 = PRT_INF;
  }
  sort2(v0, v1);
  mn = v0;
  mx = v1;
}

// TracerSurface.intersect (world_objects.py:360-383): world->object transform, primitive
// intersect, sort.  Returns the sorted pair (t0 <= t1) or (+inf, +inf).
PRT_HD void leaf_hits(const Leaf& L, double p0, double p1, double p2, double v0, double v1,
                                          double v2, double& t0, double& t1) {
  const double o0 = L.m[0] * p0 + L.m[1] * p1 + L.m[2] * p2 + L.m[3];
  const double o1 = L.m[4] * p0 + L.m[5] * p1 + L.m[6] * p2 + L.m[7];
  const double o2 = L.m[8] * p0 + L.m[9] * p1 + L.m[10] * p2 + L.m[11];
  const double d0 = L.m[0] * v0 + L.m[1] * v1 + L.m[2] * v2;
  const double d1 = L.m[4] * v0 + L.m[5] * v1 + L.m[6] * v2;
  const double d2 = L.m[8] * v0 + L.m[9] * v1 + L.m[10] * v2;
  switch (L.type) {
    case PRT_SPHERE: {  // primitives.py:241-271
      const double r = L.prm[0];
      const double a = d0 * d0 + d1 * d1 + d2 * d2;
      const double b = 2 * (d0 * o0 + d1 * o1 + d2 * o2);
      const double c = (o0 * o0 + o1 * o1 + o2 * o2) - r * r;
      const double disc = b * b - 4 * a * c;
      const double root = sqrt(fmax(0.0, disc));
      const double den_b = 2 * a;
      div2_by_value(-b + root, -b - root, den_b, t0, t1);
      if (!(disc >= 0)) {
        t0 = PRT_INF;
        t1 = PRT_INF;
      }
      sort2(t0, t1);
    } break;
    case PRT_CYLINDER: {  // primitives.py:650-712 + operations.py:28-63
      const double r = L.prm[0];
      const double a = d0 * d0 + d1 * d1;
      const double b = 2 * (d0 * o0 + d1 * o1);
      const double c = (o0 * o0 + o1 * o1) - r * r;
      const double disc = b * b - 4 * a * c;
      const bool lin = isz(a);
      const double root = sqrt(fmax(0.0, disc));
      const double den_b = 2 * a + (lin ? 1.0 : 0.0);
      double s0, s1;
      div2_by_value(-b + root, -b - root, den_b, s0, s1);
      if (!(disc >= 0)) {
        s0 = PRT_INF;
        s1 = PRT_INF;
      }
      if (lin) {
        const double l = -c / (b + (b == 0 ? 1.0 : 0.0));
        s0 = l;
        s1 = l;
        if (isz(b)) {
          s0 = (c <= 0) ? -PRT_INF : PRT_INF;
          s1 = PRT_INF;
        }
      }
      sort2(s0, s1);
      clip_z(s0, s1, L.prm[1], L.prm[2], o2, d2, L.prm[1] - o2, t0, t1);
    } break;
    case PRT_PARABOLOID: {  // primitives.py:320-399
      const double f = L.prm[0];
      const double a = d0 * d0 + d1 * d1;
      const double b = 2 * (o0 * d0 + o1 * d1) - 4 * f * d2;
      const double c = (o0 * o0 + o1 * o1) - 4 * f * o2;
      const double disc = b * b - 4 * a * c;
      const bool lin = isz(a);
      const double root = sqrt(fmax(0.0, disc));
      const double den_b = 2 * a + (lin ? 1.0 : 0.0);
      double s0, s1;
      div2_by_value(-b + root, -b - root, den_b, s0, s1);
      if (!(disc >= 0)) {
        s0 = PRT_INF;
        s1 = PRT_INF;
      }
      if (lin) {
        s0 = -c / (b + (isz(b) ? 1.0 : 0.0));
        s1 = (d2 >= 0) ? PRT_INF : -PRT_INF;
      }
      sort2(s0, s1);
      clip_z(s0, s1, 0.0, L.prm[1], o2, d2, -o2, t0, t1);
    } break;
    case PRT_PLANE: {  // primitives.py:436-492
      double xmn, xmx, ymn, ymx;
      plane_axis(o0, d0, L.prm[0], xmn, xmx);
      plane_axis(o1, d1, L.prm[1], ymn, ymx);
      const double lo = fmax(xmn, ymn);
      const double hi = fmin(xmx, ymx);
      const bool skew = isz(d2);
      double t = -o2 / (d2 + (skew ? 1.0 : 0.0));
      if (skew) t = PRT_INF;
      if (!((t >= lo) & (t <= hi))) t = PRT_INF;
      t0 = t;
      t1 = t;
    } break;
    case PRT_CUBE:  // primitives.py:516-581
      // (the guard-free slab form of the world boxes was measured here for Cubes with tame spans: no gain --
      // config 5 K1 74.9 vs 72.8 ms inlined, 83.3 ms out of line -- and it costs the lens kernels registers)
      cube_hits(L.prm, o0, o1, o2, d0, d1, d2, make_ray_inv(o0, o1, o2, d0, d1, d2, false), t0, t1);
      break;
    default:
      t0 = PRT_INF;
      t1 = PRT_INF;
  }
}

// Two Sphere leaves at once (the two refracting surfaces of a lens): the same arithmetic as the
// PRT_SPHERE case of leaf_hits, written side by side so that the two dependency chains (transform,
// quadratic, square root, reciprocal, quotients) overlap in the pipeline instead of running one
// after the other.
PRT_HD void sphere_pair_hits(const Leaf& A, const Leaf& B, double p0, double p1, double p2, double v0, double v1,
                             double v2, double& a0, double& a1, double& b0, double& b1) {
  const double ao0 = A.m[0] * p0 + A.m[1] * p1 + A.m[2] * p2 + A.m[3];
  const double bo0 = B.m[0] * p0 + B.m[1] * p1 + B.m[2] * p2 + B.m[3];
  const double ao1 = A.m[4] * p0 + A.m[5] * p1 + A.m[6] * p2 + A.m[7];
  const double bo1 = B.m[4] * p0 + B.m[5] * p1 + B.m[6] * p2 + B.m[7];
  const double ao2 = A.m[8] * p0 + A.m[9] * p1 + A.m[10] * p2 + A.m[11];
  const double bo2 = B.m[8] * p0 + B.m[9] * p1 + B.m[10] * p2 + B.m[11];
  const double ad0 = A.m[0] * v0 + A.m[1] * v1 + A.m[2] * v2;
  const double bd0 = B.m[0] * v0 + B.m[1] * v1 + B.m[2] * v2;
  const double ad1 = A.m[4] * v0 + A.m[5] * v1 + A.m[6] * v2;
  const double bd1 = B.m[4] * v0 + B.m[5] * v1 + B.m[6] * v2;
  const double ad2 = A.m[8] * v0 + A.m[9] * v1 + A.m[10] * v2;
  const double bd2 = B.m[8] * v0 + B.m[9] * v1 + B.m[10] * v2;
  const double ra = A.prm[0], rb = B.prm[0];
  const double aa = ad0 * ad0 + ad1 * ad1 + ad2 * ad2;
  const double ba = bd0 * bd0 + bd1 * bd1 + bd2 * bd2;
  const double ab = 2 * (ad0 * ao0 + ad1 * ao1 + ad2 * ao2);
  const double bb = 2 * (bd0 * bo0 + bd1 * bo1 + bd2 * bo2);
  const double ac = (ao0 * ao0 + ao1 * ao1 + ao2 * ao2) - ra * ra;
  const double bc = (bo0 * bo0 + bo1 * bo1 + bo2 * bo2) - rb * rb;
  const double adisc = ab * ab - 4 * aa * ac;
  const double bdisc = bb * bb - 4 * ba * bc;
  const double aroot = sqrt(fmax(0.0, adisc));
  const double broot = sqrt(fmax(0.0, bdisc));
  div2_by_value(-ab + aroot, -ab - aroot, 2 * aa, a0, a1);
  div2_by_value(-bb + broot, -bb - broot, 2 * ba, b0, b1);
  if (!(adisc >= 0)) {
    a0 = PRT_INF;
    a1 = PRT_INF;
  }
  if (!(bdisc >= 0)) {
    b0 = PRT_INF;
    b1 = PRT_INF;
  }
  sort2(a0, a1);
  sort2(b0, b1);
}

// A whole lens at once: the aperture Cylinder and the two Sphere surfaces of thick_lens / biconvex_lens
// (pyrayt/components.py:73-198), i.e. three of leaf_hits' cases written as one straight-line block so that
// their dependency chains (transform, quadratic, square root, reciprocal, quotients) overlap in the pipeline
// -- the kernel is bound by dependent-issue latency, not by issue slots.  Every rare case of the literal
// code (near-axis / perpendicular rays: the isclose branches of binomial_root and the cap clip; operands
// the 3-instruction division cannot take) is folded into ONE predicate: when it fails nothing here is
// used and the caller evaluates the leaves one by one with leaf_hits.  When it holds, every value below
// is computed by the same expressions as in leaf_hits, so the result is bit-identical.
PRT_HD bool lens3_hits_fast(const Leaf& Y, const Leaf& A, const Leaf& B, double p0, double p1, double p2, double v0,
                            double v1, double v2, double& y0, double& y1, double& a0, double& a1, double& b0,
                            double& b1) {
  // world -> object (world_objects.py:367-369)
  const double yo0 = Y.m[0] * p0 + Y.m[1] * p1 + Y.m[2] * p2 + Y.m[3];
  const double ao0 = A.m[0] * p0 + A.m[1] * p1 + A.m[2] * p2 + A.m[3];
  const double bo0 = B.m[0] * p0 + B.m[1] * p1 + B.m[2] * p2 + B.m[3];
  const double yo1 = Y.m[4] * p0 + Y.m[5] * p1 + Y.m[6] * p2 + Y.m[7];
  const double ao1 = A.m[4] * p0 + A.m[5] * p1 + A.m[6] * p2 + A.m[7];
  const double bo1 = B.m[4] * p0 + B.m[5] * p1 + B.m[6] * p2 + B.m[7];
  const double yo2 = Y.m[8] * p0 + Y.m[9] * p1 + Y.m[10] * p2 + Y.m[11];
  const double ao2 = A.m[8] * p0 + A.m[9] * p1 + A.m[10] * p2 + A.m[11];
  const double bo2 = B.m[8] * p0 + B.m[9] * p1 + B.m[10] * p2 + B.m[11];
  const double yd0 = Y.m[0] * v0 + Y.m[1] * v1 + Y.m[2] * v2;
  const double ad0 = A.m[0] * v0 + A.m[1] * v1 + A.m[2] * v2;
  const double bd0 = B.m[0] * v0 + B.m[1] * v1 + B.m[2] * v2;
  const double yd1 = Y.m[4] * v0 + Y.m[5] * v1 + Y.m[6] * v2;
  const double ad1 = A.m[4] * v0 + A.m[5] * v1 + A.m[6] * v2;
  const double bd1 = B.m[4] * v0 + B.m[5] * v1 + B.m[6] * v2;
  const double yd2 = Y.m[8] * v0 + Y.m[9] * v1 + Y.m[10] * v2;
  const double ad2 = A.m[8] * v0 + A.m[9] * v1 + A.m[10] * v2;
  const double bd2 = B.m[8] * v0 + B.m[9] * v1 + B.m[10] * v2;
  // the three quadratics (primitives.py:252-262, :664-667 + operations.py:39-57)
  const double yr = Y.prm[0], ar = A.prm[0], br = B.prm[0];
  const double ya = yd0 * yd0 + yd1 * yd1;
  const double aa = ad0 * ad0 + ad1 * ad1 + ad2 * ad2;
  const double ba = bd0 * bd0 + bd1 * bd1 + bd2 * bd2;
  const double yb = 2 * (yd0 * yo0 + yd1 * yo1);
  const double ab = 2 * (ad0 * ao0 + ad1 * ao1 + ad2 * ao2);
  const double bb = 2 * (bd0 * bo0 + bd1 * bo1 + bd2 * bo2);
  const double yc = (yo0 * yo0 + yo1 * yo1) - yr * yr;
  const double ac = (ao0 * ao0 + ao1 * ao1 + ao2 * ao2) - ar * ar;
  const double bc = (bo0 * bo0 + bo1 * bo1 + bo2 * bo2) - br * br;
  const double ydisc = yb * yb - 4 * ya * yc;
  const double adisc = ab * ab - 4 * aa * ac;
  const double bdisc = bb * bb - 4 * ba * bc;
  const double yroot = sqrt(fmax(0.0, ydisc));
  const double aroot = sqrt(fmax(0.0, adisc));
  const double broot = sqrt(fmax(0.0, bdisc));
  // the literal code's rare branches: isclose(a, 0) in binomial_root, isclose(d_z, 0) in the cap clip
  const bool rare = isz(ya) | isz(yd2);
  const double yd = 2 * ya + 0.0, ad = 2 * aa, bd = 2 * ba, zd = yd2 + 0.0;
  const Rcp yden = make_rcp_unchecked(yd);
  const Rcp aden = make_rcp_unchecked(ad);
  const Rcp bden = make_rcp_unchecked(bd);
  const Rcp zden = make_rcp_unchecked(zd);
  const double yn0 = -yb + yroot, yn1 = -yb - yroot;
  const double an0 = -ab + aroot, an1 = -ab - aroot;
  const double bn0 = -bb + broot, bn1 = -bb - broot;
  const double zn0 = Y.prm[1] - yo2, zn1 = Y.prm[2] - yo2;
  double s0 = div_fast(yn0, yden), s1 = div_fast(yn1, yden);
  a0 = div_fast(an0, aden);
  a1 = div_fast(an1, aden);
  b0 = div_fast(bn0, bden);
  b1 = div_fast(bn1, bden);
  double c0 = div_fast(zn0, zden), c1 = div_fast(zn1, zden);
  const unsigned my0 = mag_of(yn0), my1 = mag_of(yn1), ma0 = mag_of(an0), ma1 = mag_of(an1);
  const unsigned mb0 = mag_of(bn0), mb1 = mag_of(bn1), mz0 = mag_of(zn0), mz1 = mag_of(zn1);
  const unsigned dy = mag_of(yd), da = mag_of(ad), db = mag_of(bd), dz = mag_of(zd);  // the four denominators
  const unsigned mn = umin2(umin2(umin2(umin2(my0, my1), umin2(ma0, ma1)), umin2(umin2(mb0, mb1), umin2(mz0, mz1))),
                            umin2(umin2(dy, da), umin2(db, dz)));
  const unsigned mx = umax2(umax2(umax2(umax2(my0, my1), umax2(ma0, ma1)), umax2(umax2(mb0, mb1), umax2(mz0, mz1))),
                            umax2(umax2(dy, da), umax2(db, dz)));
  const bool ok = mags_in_window(mn, mx);
  if (!(ydisc >= 0)) {
    s0 = PRT_INF;
    s1 = PRT_INF;
  }
  if (!(adisc >= 0)) {
    a0 = PRT_INF;
    a1 = PRT_INF;
  }
  if (!(bdisc >= 0)) {
    b0 = PRT_INF;
    b1 = PRT_INF;
  }
  sort2(s0, s1);
  sort2(a0, a1);
  sort2(b0, b1);
  sort2(c0, c1);
  // cap clip of the cylinder (primitives.py:680-712)
  const double lo = fmax(s0, c0);
  const double hi = fmin(s1, c1);
  const bool hit = lo <= hi;
  y0 = hit ? lo : PRT_INF;
  y1 = hit ? hi : PRT_INF;
  return ok & !rare;
}

// TracerSurface.get_world_normals (world_objects.py:401-418) with the primitives' normal()
// (Sphere :273-296, Paraboloid :401-419, Plane :494-498, Cube :583-602, Cylinder :714-741)
PRT_HD void world_normal(const Leaf& L, double p0, double p1, double p2, double& n0,
                                             double& n1, double& n2) {
  const double q0 = L.m[0] * p0 + L.m[1] * p1 + L.m[2] * p2 + L.m[3];
  const double q1 = L.m[4] * p0 + L.m[5] * p1 + L.m[6] * p2 + L.m[7];
  const double q2 = L.m[8] * p0 + L.m[9] * p1 + L.m[10] * p2 + L.m[11];
  double a0, a1, a2;
  bool unit = false;  // object normal already unit length: the reference still divides by its norm (== 1)
  switch (L.type) {
    case PRT_SPHERE:
      a0 = q0;
      a1 = q1;
      a2 = q2;
      break;
    case PRT_PARABOLOID:
      if (iscl(q2, L.prm[1])) {
        a0 = 0;
        a1 = 0;
        a2 = 1;
        unit = true;
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
      unit = true;
      break;
    case PRT_CUBE:
      a0 = iscl(q0, L.prm[1]) ? 1.0 : (iscl(q0, L.prm[0]) ? -1.0 : 0.0);
      a1 = iscl(q1, L.prm[3]) ? 1.0 : (iscl(q1, L.prm[2]) ? -1.0 : 0.0);
      a2 = iscl(q2, L.prm[5]) ? 1.0 : (iscl(q2, L.prm[4]) ? -1.0 : 0.0);
      unit = (fabs(a0) + fabs(a1) + fabs(a2)) == 1.0;  // one face: dividing by the norm (exactly 1) changes nothing
      break;
    default:  // PRT_CYLINDER
      a0 = q0;
      a1 = q1;
      a2 = 0;
      if (L.prm[3] != 0.0) {
        if (iscl(q2, L.prm[1])) {
          a0 = 0;
          a1 = 0;
          a2 = -1;
          unit = true;
        }
        if (iscl(q2, L.prm[2])) {
          a0 = 0;
          a1 = 0;
          a2 = 1;
          unit = true;
        }
      }
      break;
  }
  if (!unit) {
    div3_by_value(a0, a1, a2, sqrt(a0 * a0 + a1 * a1 + a2 * a2), a0, a1, a2);
  }
  // M_obj^T n_obj, w dropped, normalise, flip (world_objects.py:411-418)
  double w0 = L.m[0] * a0 + L.m[4] * a1 + L.m[8] * a2;
  double w1 = L.m[1] * a0 + L.m[5] * a1 + L.m[9] * a2;
  double w2 = L.m[2] * a0 + L.m[6] * a1 + L.m[10] * a2;
  div3_by_value(w0, w1, w2, sqrt(w0 * w0 + w1 * w1 + w2 * w2), n0, n1, n2);
  n0 *= L.nscale;
  n1 *= L.nscale;
  n2 *= L.nscale;
}

// ---------------------------------------------------------------- CSG hit lists
//
// The reference keeps fixed-length lists padded with +inf (csg.py:152-160).  Entries at
// +inf always sort behind every other entry and only influence other +inf entries
// (array_csg's counts are a forward cumsum and the roll() wrap-around reads the last
// count, which always equals the start value because enter/exit entries pair up), so a
// list is represented by its prefix of entries < +inf.  Parity is by *index in the child
// list* exactly as in csg.py:41,:46.

template <class T>  // T = double (FP64 path) or float (FP32 fast mode)
struct HitStackT {
  typedef T value_type;
  T t[kMaxDepth * 2][kMaxSlots];
  unsigned short leaf[kMaxDepth * 2][kMaxSlots];
  int len[kMaxDepth];
  unsigned flags;  // bit s: which of the two buffers of level s is live
};
typedef HitStackT<double> HitStack;

template <class Stack>
PRT_HD int buf_of(const Stack& S, int lvl) { return lvl * 2 + ((S.flags >> lvl) & 1); }

// streaming form of array_csg (csg.py:13-61) + the two argsorts of CSGSurface.intersect (:138-149):
// L = live buffer of level `lvl`; R = (r_t, r_leaf, nR) supplied by get_r; result replaces L.
template <class Stack, class GetRT, class GetRL>
PRT_HD void merge_lists(Stack& S, int lvl, int op, int nR, GetRT r_t, GetRL r_leaf, bool& tie) {
  const int src = buf_of(S, lvl);
  const int dst = src ^ 1;
  const int nL = S.len[lvl];
  int i = 0, j = 0, k = 0;
  int cnt = (op == PRT_DIFFERENCE) ? 1 : 0;
  const bool is_union = (op == PRT_UNION);
  const int rflip = (op == PRT_DIFFERENCE) ? 1 : 0;
  while (i < nL || j < nR) {
    bool takeL;
    typename Stack::value_type tl = 0, tr = 0;
    if (j >= nR) {
      takeL = true;
      tl = S.t[src][i];
    } else if (i >= nL) {
      takeL = false;
      tr = r_t(j);
    } else {
      tl = S.t[src][i];
      tr = r_t(j);
      takeL = !(tr < tl);  // stable: the left child wins ties (csg.py:36-38, stable argsort)
      if (tl == tr) tie = true;
    }
    const int exitf = takeL ? (i & 1) : ((j & 1) ^ rflip);
    const int prev = cnt;
    cnt += exitf ? -1 : 1;
    const bool keep = is_union ? ((cnt != 0) != (prev != 0)) : (cnt == 2 || prev == 2);
    if (keep) {
      S.t[dst][k] = takeL ? tl : tr;
      S.leaf[dst][k] = takeL ? S.leaf[src][i] : (unsigned short)r_leaf(j);
      ++k;
    }
    if (takeL)
      ++i;
    else
      ++j;
  }
  S.len[lvl] = k;
  S.flags ^= (1u << lvl);
}

// component.intersect for one ray: runs ops [begin, end); the result is level 0 of S.
//
// `best_t` enables the exact pruning of whole components: when the encoder has proven that the
// root bounding box contains the solid (Op.c & 1, see prt_encode.h) a component whose box lies
// entirely behind the ray, or entirely beyond the nearest hit found so far, cannot change the
// result of _st_propagate and is skipped (returns false).  prune = false evaluates everything
// (component.intersect must also report the hits behind the ray).
PRT_HD bool eval_component(const SceneView& sc, int begin, int end, double p0, double p1, double p2, double v0,
                           double v1, double v2, const RayInv& inv, bool prune, double best_t, HitStack& S,
                           bool& tie) {
  int sp = 0;
  int pc = begin;
  while (pc < end) {
    const Op op = sc.ops[pc];
    if (op.kind == OP_ENTER) {
      double b0, b1;
      cube_hits(sc.aabb + 6 * op.a, p0, p1, p2, v0, v1, v2, inv, b0, b1);
      if (!(b0 < PRT_INF)) {  // cube_hits returns finite values or (+inf,+inf): csg.py:126-128
        if (pc == begin) return false;  // root culled: no hits at all
        S.len[sp++] = 0;
        pc = op.b;
        continue;
      }
      if (prune && pc == begin && (op.c & 1) && (b1 < -kCullMargin || b0 > best_t + kCullMargin)) return false;
    } else if (op.kind == OP_LEAF) {
      double t0, t1;
      leaf_hits(sc.leaves[op.a], p0, p1, p2, v0, v1, v2, t0, t1);
      const int b = buf_of(S, sp);
      const int n = (t0 < PRT_INF) ? ((t1 < PRT_INF) ? 2 : 1) : 0;
      S.t[b][0] = t0;
      S.t[b][1] = t1;
      S.leaf[b][0] = (unsigned short)op.a;
      S.leaf[b][1] = (unsigned short)op.a;
      S.len[sp++] = n;
    } else if (op.kind == OP_MERGE_LEAF) {
      double t0, t1;
      leaf_hits(sc.leaves[op.b], p0, p1, p2, v0, v1, v2, t0, t1);
      const int n = (t0 < PRT_INF) ? ((t1 < PRT_INF) ? 2 : 1) : 0;
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

// ---------------------------------------------------------------- register-only left-deep components
//
// Every reference factory builds a left-deep tree of two or three leaves ((A op1 B) op2 C,
// SURVEY 8(a3)).  For those shapes the two stable merges of CSGSurface.intersect are evaluated in
// closed form, entirely in registers.
//
// array_csg (csg.py:13-61) walks the merged, stably sorted entries with a counter and keeps the entries
// at which the counter enters or leaves its "inside" value.  Entry k of a child list enters the child
// when k is even and leaves it when k is odd (csg.py:41,:46), so the counter is a function of "inside
// left child" and "inside right child", and the kept entries are exactly the entries at which
// op(inside L, inside R) changes: UNION keeps zero <-> non-zero (:53-54), INTERSECT / DIFFERENCE keep
// count == 2 and the entry after it (:57-59; DIFFERENCE starts at 1 and flips the right child's sign,
// :45-48).  The kept entries of the first merge alternate enter / leave again, so the second merge sees
// op1(inside A, inside B) as its left state.  The whole component therefore keeps an entry x iff
//      F(inA, inB, inC) = op2(op1(inA, inB), inC)
// changes at x, where the state of the *other* leaves at x follows from how many of their entries sort
// before x (one: inside; none or both: outside), with the reference's tie order: stable sorts put the left
// child first, i.e. A before B before C on equal keys.  F is an 8-bit truth table made by the encoder
// (Comp.tt).  The nearest hit is the first kept entry with t > 0 in that order.  (Cross-checked against
// the streaming merge_lists on every sorted pair over {-inf, -2, -1, 1, 2, 3, +inf} for A, B and C, all
// nine operation pairs, with and without the inner bounding box: tests/test_kernel_emul.py.)

// running best of one component: first positive kept entry in merged order
template <class T>
PRT_HD void take_hit(bool keep, T t, int leaf, T& ct, int& cl) {
  const bool take = keep & (t > 0) & (t < ct);
  ct = take ? t : ct;
  cl = take ? leaf : cl;
}

// does F change when leaf `own` (bit mask 1 / 2 / 4) toggles, the state just before being `before`?
PRT_HD bool tt_changes(unsigned tt, unsigned before, unsigned own) {
  return (((tt >> before) ^ (tt >> (before ^ own))) & 1u) != 0u;
}

// (a0,a1), (b0,b1), (c0,c1): sorted hit pairs of leaves A, B, C (+inf = missing entry; LEFT2: c = +inf and
// tt ignores C).  Returns the nearest positive kept entry (ct, cl) and whether equal keys were compared.
// (T = double on the FP64 path, float in the FP32 fast mode.)
// TIES = false compiles the equal-key detection (a diagnostic counter, prt_counters.tie_rays) out.
template <bool TIES = true, class T>
PRT_HD void left_deep_first_hit(unsigned tt, T a0, T a1, T b0, T b1, T c0, T c1, int la, int lb, int lc, T& ct,
                                int& cl, bool& tie) {
  const T kInf = (T)PRT_INF;
  const bool va0 = a0 < kInf, va1 = a1 < kInf, vb0 = b0 < kInf, vb1 = b1 < kInf;
  const bool vc0 = c0 < kInf, vc1 = c1 < kInf;
  // y < x for every pair of entries of different leaves (bitwise logic on purpose: no short-circuit branches)
  const bool b0a0 = b0 < a0, b1a0 = b1 < a0, b0a1 = b0 < a1, b1a1 = b1 < a1;
  const bool c0a0 = c0 < a0, c1a0 = c1 < a0, c0a1 = c0 < a1, c1a1 = c1 < a1;
  const bool c0b0 = c0 < b0, c1b0 = c1 < b0, c0b1 = c0 < b1, c1b1 = c1 < b1;
  if (TIES)
    tie |= (va0 & vb0 & (a0 == b0)) | (va0 & vb1 & (a0 == b1)) | (va1 & vb0 & (a1 == b0)) | (va1 & vb1 & (a1 == b1)) |
           (va0 & vc0 & (a0 == c0)) | (va0 & vc1 & (a0 == c1)) | (va1 & vc0 & (a1 == c0)) | (va1 & vc1 & (a1 == c1)) |
           (vb0 & vc0 & (b0 == c0)) | (vb0 & vc1 & (b0 == c1)) | (vb1 & vc0 & (b1 == c0)) | (vb1 & vc1 & (b1 == c1));
  // state of the other leaves just before each entry.  An entry y of a later leaf precedes x only when
  // y < x; an entry y of an earlier leaf precedes x when y <= x, i.e. when !(x < y): in the XOR of a
  // leaf's two entries the two negations cancel, so the same comparison bits serve both directions.
  const unsigned inB_a0 = (unsigned)(b0a0 ^ b1a0), inB_a1 = (unsigned)(b0a1 ^ b1a1);
  const unsigned inC_a0 = (unsigned)(c0a0 ^ c1a0), inC_a1 = (unsigned)(c0a1 ^ c1a1);
  const unsigned inA_b0 = (unsigned)(b0a0 ^ b0a1), inA_b1 = (unsigned)(b1a0 ^ b1a1);
  const unsigned inC_b0 = (unsigned)(c0b0 ^ c1b0), inC_b1 = (unsigned)(c0b1 ^ c1b1);
  const unsigned inA_c0 = (unsigned)(c0a0 ^ c0a1), inA_c1 = (unsigned)(c1a0 ^ c1a1);
  const unsigned inB_c0 = (unsigned)(c0b0 ^ c0b1), inB_c1 = (unsigned)(c1b0 ^ c1b1);
  // entry 0 of a leaf enters it (own bit 0 -> 1), entry 1 leaves it (1 -> 0)
  const bool ka0 = va0 & tt_changes(tt, (inB_a0 << 1) | (inC_a0 << 2), 1u);
  const bool ka1 = va1 & tt_changes(tt, 1u | (inB_a1 << 1) | (inC_a1 << 2), 1u);
  const bool kb0 = vb0 & tt_changes(tt, inA_b0 | (inC_b0 << 2), 2u);
  const bool kb1 = vb1 & tt_changes(tt, inA_b1 | 2u | (inC_b1 << 2), 2u);
  const bool kc0 = vc0 & tt_changes(tt, inA_c0 | (inB_c0 << 1), 4u);
  const bool kc1 = vc1 & tt_changes(tt, inA_c1 | (inB_c1 << 1) | 4u, 4u);
  ct = kInf;
  cl = -1;
  take_hit(ka0, a0, la, ct, cl);
  take_hit(ka1, a1, la, ct, cl);
  take_hit(kb0, b0, lb, ct, cl);
  take_hit(kb1, b1, lb, ct, cl);
  take_hit(kc0, c0, lc, ct, cl);
  take_hit(kc1, c1, lc, ct, cl);
}

// The ray's dominant axis, for the ordered traversal of the boxed components (nearest_hit).  Along one
// axis the far face of a box bounds its exit distance from above and the near face its entry distance from
// below, so with u = +-x_k (sign chosen so that the ray runs towards +u), uo the origin and a = |d_k|:
//   far_u  < uo - 3 m a          =>  the box ends at least 3 m behind the ray: no positive hit possible
//   near_u > uo + (best + 2 m) a =>  the box begins more than 2 m - rounding beyond the best hit so far
// (m = kCullMargin; |u| <= 1e6 and 1/4 <= a <= 4 keep the rounding of both right-hand sides below 1e-9,
// far inside the margins; hit parameters and box parameters of the same point differ by ~1e-13).  Both
// imply that the exact proven-box test of the component would prune it as well.
struct DomAxis {
  double thr_far, a;  // uo - 3 m a (boxes ending before it are behind the ray), |direction component|
  int table;          // 2 * axis + sign: which traversal table; < 0: quick tests not usable for this ray
};

PRT_HD DomAxis make_dom_axis(double p0, double p1, double p2, double v0, double v1, double v2, bool small_boxes) {
  const double a0 = fabs(v0), a1 = fabs(v1), a2 = fabs(v2);
  const int k = (a0 >= a1) ? ((a0 >= a2) ? 0 : 2) : ((a1 >= a2) ? 1 : 2);
  DomAxis d;
  d.a = (k == 0) ? a0 : ((k == 1) ? a1 : a2);
  const double o = (k == 0) ? p0 : ((k == 1) ? p1 : p2);
  const double v = (k == 0) ? v0 : ((k == 1) ? v1 : v2);
  const int s = v < 0 ? 1 : 0;
  d.thr_far = (s ? -o : o) - (3 * kCullMargin) * d.a;
  d.table = (small_boxes && fabs(o) <= 1e6 && d.a >= 0.25 && d.a <= 4.0) ? 2 * k + s : -1;
  return d;
}

// shapes 2/3: (A op1 B) [op2 C] with their bounding boxes (csg.py:118-160 for a left-deep tree)
template <bool TIES = true>
PRT_HD void eval_left_deep(const SceneView& sc, const Comp& C, double p0, double p1, double p2, double v0, double v1,
                           double v2, const RayInv& inv, double best_t, double& ct, int& cl, bool& tie) {
  ct = PRT_INF;
  cl = -1;
  const int shape = C.shape;
  double b0, b1;
  cube_hits(C.root_box, p0, p1, p2, v0, v1, v2, inv, b0, b1);
  if (!(b0 < PRT_INF)) return;  // csg.py:126-133
  // proven-box pruning: a box behind the ray or beyond the best hit so far cannot matter
  if (((C.flags & 1) != 0) & ((b1 < -kCullMargin) | (b0 > best_t + kCullMargin))) return;
  bool inner_hit = true;
  if (shape == SHAPE_LEFT3) {
    cube_hits(C.inner_box, p0, p1, p2, v0, v1, v2, inv, b0, b1);
    inner_hit = b0 < PRT_INF;
  }
  const int la = C.leaf_a, lb = C.leaf_b, lc = C.leaf_c;
  double a0 = PRT_INF, a1 = PRT_INF, q0 = PRT_INF, q1 = PRT_INF, c0 = PRT_INF, c1 = PRT_INF;
  // one (not unrolled) loop over the leaves keeps a single copy of leaf_hits in the hot loop
  // Lenses carry two spherical surfaces ((aperture, sphere, sphere) for thick_lens, (sphere, sphere,
  // aperture) for biconvex_lens): those two leaves are evaluated side by side (see sphere_pair_hits),
  // the remaining leaf by the loop below.  `todo` = bit k set: leaf k still to be evaluated.
  unsigned todo = inner_hit ? ((shape == SHAPE_LEFT3) ? 7u : 3u) : 4u;
#ifndef PRT_NO_LENS3
  if (inner_hit & ((C.flags & 24) != 0)) {
    // a lens: capped Cylinder + two Spheres, the cylinder first (thick_lens, bit 3) or last (biconvex_lens, bit 4)
    const bool cyl_first = (C.flags & 8) != 0;
    const Leaf& Y = sc.leaves[cyl_first ? la : lc];
    const Leaf& A = sc.leaves[cyl_first ? lb : la];
    const Leaf& B = sc.leaves[cyl_first ? lc : lb];
    double y0, y1, s0, s1, u0, u1;
    if (lens3_hits_fast(Y, A, B, p0, p1, p2, v0, v1, v2, y0, y1, s0, s1, u0, u1)) {
      a0 = cyl_first ? y0 : s0;
      a1 = cyl_first ? y1 : s1;
      q0 = cyl_first ? s0 : u0;
      q1 = cyl_first ? s1 : u1;
      c0 = cyl_first ? u0 : y0;
      c1 = cyl_first ? u1 : y1;
      todo = 0u;
    }
  }
#endif
  if (inner_hit & (todo != 0u)) {
    const int ta = sc.leaves[la].type, tb = sc.leaves[lb].type;
    const int tc = (shape == SHAPE_LEFT3) ? sc.leaves[lc].type : 0;
    if ((tb == PRT_SPHERE) & (tc == PRT_SPHERE)) {
      sphere_pair_hits(sc.leaves[lb], sc.leaves[lc], p0, p1, p2, v0, v1, v2, q0, q1, c0, c1);
      todo = 1u;
    } else if ((ta == PRT_SPHERE) & (tb == PRT_SPHERE)) {
      sphere_pair_hits(sc.leaves[la], sc.leaves[lb], p0, p1, p2, v0, v1, v2, a0, a1, q0, q1);
      todo &= 4u;
    }
  }
#pragma unroll 1
  for (int k = 0; k < 3; ++k) {
    if (!((todo >> k) & 1u)) continue;
    const int lf = (k == 0) ? la : ((k == 1) ? lb : lc);
    double t0, t1;
    leaf_hits(sc.leaves[lf], p0, p1, p2, v0, v1, v2, t0, t1);
    if (k == 0) {
      a0 = t0;
      a1 = t1;
    } else if (k == 1) {
      q0 = t0;
      q1 = t1;
    } else {
      c0 = t0;
      c1 = t1;
    }
  }
  // a missed inner box empties (A op1 B) whatever the leaves say (csg.py:126-133): a0..q1 are still +inf
  left_deep_first_hit<TIES>((unsigned)C.tt, a0, a1, q0, q1, c0, c1, la, lb, lc, ct, cl, tie);
}

// nearest-hit of _st_propagate over all components (pyrayt/_pyrayt.py:376-386).  The reference visits the
// components in list order and replaces the running best only by a strictly smaller distance (:384), i.e.
// the smallest distance wins and, among equal distances, the earliest component.  Stated that way the result
// does not depend on the visiting order, so the boxed components are visited in the order the ray meets
// their boxes (see OrderEntry / DomAxis): the entries behind the ray are skipped by bisection, and the walk
// stops at the first box that begins beyond the best hit found so far -- in a lens train that is one or two
// evaluated components per generation instead of a test of every component.  Components without a usable
// box are always evaluated; rays the quick tests cannot serve visit every component in list order.
// GENERIC = false compiles the interpreter for arbitrary trees out: scenes whose components are all
// bare leaves or left-deep (every reference factory) run a smaller kernel with no lists in memory.
// TIES = false: the left-deep components do not look for equal merge keys (tie_rays is then only fed by the
// generic interpreter); the trace kernel counts ties under PRT_FLAG_DIAGNOSE only.
template <bool GENERIC, bool TIES = true>
PRT_HD void nearest_hit(const SceneView& sc, double p0, double p1, double p2, double v0, double v1, double v2,
                        int skip, HitStack* S, double& best_t, int& best_leaf, bool& tie) {
  best_t = PRT_INF;
  best_leaf = -1;
  const RayInv inv = make_ray_inv(p0, p1, p2, v0, v1, v2, (sc.h->flags & 1) != 0);
  const DomAxis dom = make_dom_axis(p0, p1, p2, v0, v1, v2, (sc.h->flags & 4) != 0);
  const bool quick = dom.table >= 0;                      // the threshold tests are usable for this ray
  const bool ordered = quick & ((sc.h->flags & 8) != 0);  // walk the boxes in ray order (scenes with many components)
  // phase 0 walks a table: the ray-ordered one from the bisection point, or the list-order one from its start
  // (every lane of the warp then sees component c in iteration c); phase 1: the components without a box
  const int n0 = ordered ? sc.h->n_boxed : sc.h->n_components;
  const int n1 = ordered ? sc.h->n_unboxed : 0;
  const OrderEntry* tab = (ordered ? sc.order : sc.bycomp) + (quick ? dom.table : 0) * n0;
  const double thr_far = quick ? dom.thr_far : -PRT_INF;
  double thr_near = PRT_INF;
  int j = 0;
  if (ordered) {
    int hi = n0;  // first entry whose running-maximum far face is not provably behind the ray
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
      const OrderEntry e = tab[j];
      const bool beyond = e.near_u > thr_near;
      const bool stop = ordered & beyond;  // ray order: every later box begins even further away
      ++j;
      if (stop | (j >= n0)) {
        phase = 1;
        j = 0;
      }
      if (beyond | (e.far_u < thr_far)) continue;
      c = e.comp;
    } else {
      if (j >= n1) break;
      c = sc.unboxed[j];
      ++j;
    }
    if (c == skip) continue;  // convex solid the ray left in the previous generation
    const Comp& C = sc.comps[c];
    const int shape = C.shape;
    double ct = PRT_INF;
    int cl = -1;
    if (shape == SHAPE_LEAF) {  // bare TracerSurface component: no list needed
      double t0, t1;
      leaf_hits(sc.leaves[C.leaf_a], p0, p1, p2, v0, v1, v2, t0, t1);
      ct = (t0 > 0) ? t0 : ((t1 > 0) ? t1 : PRT_INF);
      cl = C.leaf_a;
    } else if (shape == SHAPE_LEFT2 || shape == SHAPE_LEFT3) {
      eval_left_deep<TIES>(sc, C, p0, p1, p2, v0, v1, v2, inv, best_t, ct, cl, tie);
    } else if (GENERIC) {
      S->flags = 0;
      if (eval_component(sc, C.begin, C.end, p0, p1, p2, v0, v1, v2, inv, true, best_t, *S, tie)) {
        const int b = buf_of(*S, 0);
        const int n = S->len[0];
        for (int q = 0; q < n; ++q) {  // sorted: the first positive entry is the argmin of where(hits>0)
          const double t = S->t[b][q];
          if (t > 0) {
            ct = t;
            cl = S->leaf[b][q];
            break;
          }
        }
      }
    }
    // smallest distance, then earliest component (== the reference's in-order strict `<`)
    bool better = ct < best_t;
    if ((ct == best_t) & (ct < PRT_INF)) better = c < sc.leaves[best_leaf].comp;  // rare: coincident surfaces
    if (better) {
      best_t = ct;
      best_leaf = cl;
      // uo + (ct + 2 m) a, written from thr_far = uo - 3 m a (the re-association moves it by < 1e-9)
      if (quick) thr_near = thr_far + (ct + 5 * kCullMargin) * dom.a;
    }
  }
}


// ---------------------------------------------------------------- one generation of one ray

struct RayState {
  double p0, p1, p2, v0, v1, v2, wl, nidx;  // generation / intensity / id are only copied to the rows
  int skip;  // component the ray has just left for good (convex solid, see Comp.flags bit 1), or -1
};

struct StepOut {
  bool row;                  // a frame row was produced this generation
  double sid;                // surface id
  double e0, e1, e2;         // hit point (x1,y1,z1)
  double t0n, t1n, t2n;      // unit tilt of the incoming direction (TILT = false: only set for glass hits)
  double nv0, nv1, nv2;      // direction after the interaction
  double n_next;             // refractive index after the interaction
  int skip;                  // component that cannot be hit in the next generation, or -1
};

// per-ray event counters packed into two words (they stay live for the whole kernel):
// w0 = generations entered | rows << 16 ; w1 = mirror rows | flag bits.  generation_limit <= 65535.
struct StepCounters {
  unsigned w0, w1;
};
constexpr unsigned kCtrTie = 1u << 16, kCtrUntr = 1u << 17, kCtrNan = 1u << 18, kCtrLim = 1u << 19,
                   kCtrAbs = 1u << 20, kCtrGraze = 1u << 21, kCtrSeam = 1u << 22;

// PRT_FLAG_DIAGNOSE (include/pyrayt_b200.h): does the nearest-hit search answer differently when the origin
// is displaced by 1e-9 x max(1, |p|_inf) perpendicular to the direction?  Returns kCtrGraze (hit <-> miss)
// and / or kCtrSeam (another surface).  Plain IEEE arithmetic in this order (the CPU checker restates it).
template <bool GENERIC>
PRT_HD unsigned diagnose_generation(const SceneView& sc, double p0, double p1, double p2, double v0, double v1,
                                    double v2, double vn, int skip, HitStack* S, int hit_leaf) {
  const double a0 = fabs(v0), a1 = fabs(v1), a2 = fabs(v2);
  const int k = (a0 <= a1) ? ((a0 <= a2) ? 0 : 2) : ((a1 <= a2) ? 1 : 2);
  // e1 = v x axis_k
  double e0 = (k == 0) ? 0.0 : ((k == 1) ? -v2 : v1);
  double e1 = (k == 0) ? v2 : ((k == 1) ? 0.0 : -v0);
  double e2 = (k == 0) ? -v1 : ((k == 1) ? v0 : 0.0);
  const double en = sqrt(e0 * e0 + e1 * e1 + e2 * e2);
  e0 = e0 / en;
  e1 = e1 / en;
  e2 = e2 / en;
  // e2 = (v x e1) / |v|
  const double f0 = (v1 * e2 - v2 * e1) / vn, f1 = (v2 * e0 - v0 * e2) / vn, f2 = (v0 * e1 - v1 * e0) / vn;
  const double delta = 1e-9 * fmax(1.0, fmax(fabs(p0), fmax(fabs(p1), fabs(p2))));
  unsigned out = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int q = 0; q < 4; ++q) {
    const double s = (q & 1) ? -delta : delta;
    const double d0 = (q < 2) ? e0 : f0, d1 = (q < 2) ? e1 : f1, d2 = (q < 2) ? e2 : f2;
    double t;
    int leaf;
    bool tie = false;
    nearest_hit<GENERIC>(sc, p0 + s * d0, p1 + s * d1, p2 + s * d2, v0, v1, v2, skip, S, t, leaf, tie);
    if (leaf != hit_leaf) out |= ((leaf < 0) | (hit_leaf < 0)) ? kCtrGraze : kCtrSeam;
  }
  return out;
}


// A ray that leaves a convex solid through one of its faces (its new direction has a clearly positive
// component along the outward normal there) cannot hit that solid again until it changes direction:
// the whole solid lies behind the tangent plane.  Returns that component for the next generation's
// nearest-hit search to skip, else -1.  (`out_dot` = new direction . outward unit normal; the
// threshold keeps grazing exits, where rounding could matter, on the ordinary path.)
PRT_HD int leaves_for_good(const SceneView& sc, const Leaf& L, double out_dot) {
  return ((L.comp >= 0) && (out_dot > 1e-3) && (sc.comps[L.comp].flags & 2)) ? L.comp : -1;
}

// first half of a generation: is the ray still travelling?  Returns its speed |v| (0: it is not)
PRT_HD double step_speed(const RayState& r, StepCounters& c) {
  const double vn = sqrt(r.v0 * r.v0 + r.v1 * r.v1 + r.v2 * r.v2);
  if (isz(vn)) return 0.0;  // absorbed / zero direction (_pyrayt.py:415)
  if (isnan(r.v0) | isnan(r.v1) | isnan(r.v2)) {
    c.w1 |= kCtrNan;  // every hit compares false in the reference -> miss
    return 0.0;
  }
  c.w0 += 1u;
  return vn;
}

// the row's tilt columns: the unit incoming direction (pyrayt/_pyrayt.py:177); `vn` = |v| = step_speed.
// One definition, used by the interaction (refract()'s normalised vector is the same quotient,
// operations.py:125) and by the ordering pass that rebuilds the columns from the staged direction.
PRT_HD void unit_tilt(double v0, double v1, double v2, double vn, double& t0, double& t1, double& t2) {
  div3_by_value(v0, v1, v2, vn, t0, t1, t2);
}

// second half: _st_interact for a ray whose nearest hit is (best_t, best_leaf >= 0); `vn` = step_speed.
// TILT = false leaves o.t0n..t2n unset unless the material needs them (the trace kernel stages the
// direction itself and the ordering pass calls unit_tilt).
template <bool TILT = true>
PRT_HD bool step_interact(const SceneView& sc, const RayState& r, int g, int generation_limit, double vn,
                          double best_t, int best_leaf, StepOut& o, StepCounters& c) {
  o.row = false;
  const Leaf& L = sc.leaves[best_leaf];
  o.e0 = r.p0 + r.v0 * best_t;  // :404-407
  o.e1 = r.p1 + r.v1 * best_t;
  o.e2 = r.p2 + r.v2 * best_t;
  o.n_next = r.nidx;
  o.skip = -1;
  const bool is_glass = (L.mat == PRT_MAT_GLASS_CONST) | (L.mat == PRT_MAT_GLASS_SELLMEIER);
  if (TILT || is_glass) unit_tilt(r.v0, r.v1, r.v2, vn, o.t0n, o.t1n, o.t2n);
  bool goes_on = true;
  // mirrors and glasses need the surface normal (one call site: the primitive switch is large)
  double n0 = 0, n1 = 0, n2 = 0;
  if (L.mat == PRT_MAT_MIRROR || L.mat == PRT_MAT_GLASS_CONST || L.mat == PRT_MAT_GLASS_SELLMEIER)
    world_normal(L, o.e0, o.e1, o.e2, n0, n1, n2);
  if (L.mat == PRT_MAT_ABSORBER) {  // materials.py:47-50
    o.nv0 = 0;
    o.nv1 = 0;
    o.nv2 = 0;
    goes_on = false;  // recorded now, dead next generation (zero direction)
    c.w1 |= kCtrAbs;
  } else if (L.mat == PRT_MAT_MIRROR) {  // materials.py:58-62, operations.py:105-107
    c.w1 += 1u;
    const double dots = r.v0 * n0 + r.v1 * n1 + r.v2 * n2;
    o.nv0 = r.v0 - 2 * n0 * dots;
    o.nv1 = r.v1 - 2 * n1 * dots;
    o.nv2 = r.v2 - 2 * n2 * dots;
    o.skip = leaves_for_good(sc, L, o.nv0 * n0 + o.nv1 * n1 + o.nv2 * n2);
  } else if (L.mat == PRT_MAT_GLASS_CONST || L.mat == PRT_MAT_GLASS_SELLMEIER) {
    // materials.py:70-75,:112-118,:136-145 ; operations.py:110-162
    const double u0 = o.t0n, u1 = o.t1n, u2 = o.t2n;
    const double cp = u0 * n0 + u1 * n1 + u2 * n2;
    const bool exiting = cp > 0;  // leaving the glass always enters n = 1 (operations.py:134)
    double n2l = 1.0;
    if (!exiting) {  // index_at() only matters when entering
      if (L.mat == PRT_MAT_GLASS_CONST) {
        n2l = L.matp[0];
      } else {
        const double w2 = r.wl * r.wl;
        n2l = sqrt(1 + (L.matp[0] * w2) / (w2 - L.matp[3]) + (L.matp[1] * w2) / (w2 - L.matp[4]) +
                   (L.matp[2] * w2) / (w2 - L.matp[5]));
      }
    } else {
      n0 = -n0;
      n1 = -n1;
      n2 = -n2;
    }
    const double q = r.nidx / n2l;
    const double c1 = exiting ? cp : -cp;
    const double rad = 1 - (q * q) * (1 - c1 * c1);
    if (rad > 0) {
      const double k = q * c1 - sqrt(rad);
      o.nv0 = q * u0 + k * n0;
      o.nv1 = q * u1 + k * n1;
      o.nv2 = q * u2 + k * n2;
      o.n_next = n2l;
    } else {  // total internal reflection keeps the index
      const double k = 2 * c1;
      o.nv0 = u0 + k * n0;
      o.nv1 = u1 + k * n1;
      o.nv2 = u2 + k * n2;
    }
    div3_by_value(o.nv0, o.nv1, o.nv2, sqrt(o.nv0 * o.nv0 + o.nv1 * o.nv1 + o.nv2 * o.nv2), o.nv0, o.nv1, o.nv2);
    // when exiting, n0..n2 were flipped above and now hold minus the outward normal: a refracted ray
    // has a positive outward component, a totally reflected one a negative one
    o.skip = leaves_for_good(sc, L, exiting ? -(o.nv0 * n0 + o.nv1 * n1 + o.nv2 * n2) : -1.0);
  } else {
    c.w1 |= kCtrUntr;  // the reference raises AttributeError here
    return false;
  }
  c.w0 += 1u << 16;
  o.row = true;
  o.sid = L.sid;
  if (g + 1 == generation_limit) {
    c.w1 |= kCtrLim;
    return false;
  }
  return goes_on;
}


// _st_propagate + _st_interact for one ray (pyrayt/_pyrayt.py:370-452).  Fills `o`;
// returns true when the ray goes on to generation g+1.
template <bool GENERIC>
PRT_HD bool trace_step(const SceneView& sc, const RayState& r, int g, int generation_limit, HitStack* S,
                       StepOut& o, StepCounters& c) {
  o.row = false;
  const double vn = step_speed(r, c);
  if (vn == 0.0) return false;
  double best_t;
  int best_leaf;
  bool tie = false;
  nearest_hit<GENERIC>(sc, r.p0, r.p1, r.p2, r.v0, r.v1, r.v2, r.skip, S, best_t, best_leaf, tie);
  if (tie) c.w1 |= kCtrTie;
  if (best_leaf < 0) return false;  // miss: dead, nothing recorded (:415-420)
  return step_interact<true>(sc, r, g, generation_limit, vn, best_t, best_leaf, o, c);
}

// storage for the generic interpreter's hit lists: only the GENERIC kernel variants carry it
template <bool GENERIC>
struct StackFor {
  typedef HitStack type;
  PRT_HD static HitStack* ptr(HitStack& s) { return &s; }
};
template <>
struct StackFor<false> {
  typedef char type;
  PRT_HD static HitStack* ptr(char&) { return nullptr; }
};

// generation g -> g+1 (pyrayt/_pyrayt.py:436-449)
PRT_HD void advance_ray(RayState& r, const StepOut& o, int g, double ray_offset) {
  (void)g;
  r.skip = o.skip;
  r.nidx = o.n_next;
  r.v0 = o.nv0;
  r.v1 = o.nv1;
  r.v2 = o.nv2;
  r.p0 = o.e0 + ray_offset * o.nv0;
  r.p1 = o.e1 + ray_offset * o.nv1;
  r.p2 = o.e2 + ray_offset * o.nv2;
}

}  // namespace prt
