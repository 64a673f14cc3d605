// Literal, fixed-length evaluation of component.intersect (tinygfx/g3d/world_objects.py:360-383,
// csg.py:118-160): every slot the reference returns -- the +inf ones and the surface ids they carry
// included -- for the plugin entry point prt_intersect.  The trace kernel never needs those slots
// (an infinite hit can not be the nearest one) and uses the streaming / closed-form merges of
// prt_device.cuh; this version keeps whole lists in local memory and is not on the hot path.
#pragma once
#include "prt_device.cuh"

namespace prt {

struct LitList {
  double t[kMaxSlots];
  short leaf[kMaxSlots];  // -1: the reference reports surface id -1 (a CSG node whose box was missed)
  int n;
};

// np.argsort(kind="stable") of v[0..n): insertion sort, NaN last
PRT_HD void lit_argsort(const double* v, int n, unsigned char* idx) {
  for (int i = 0; i < n; ++i) idx[i] = (unsigned char)i;
  for (int i = 1; i < n; ++i) {
    const unsigned char k = idx[i];
    const double x = v[k];
    int j = i - 1;
    while (j >= 0) {
      const double y = v[idx[j]];
      const bool lt = (y != y) ? (x == x) : (x < y);  // x sorts before y
      if (!lt) break;
      idx[j + 1] = idx[j];
      --j;
    }
    idx[j + 1] = k;
  }
}

// CSGSurface.intersect below the box test (csg.py:134-160): array_csg on the concatenated children
// (index parity, DIFFERENCE flips the right child, np.roll wrap-around), then the hit-order argsort
// with the surfaces carried along.
PRT_HD void lit_merge(const LitList& L, const LitList& R, int op, LitList& out) {
  const int n1 = L.n, m = L.n + R.n;
  double merged[kMaxSlots], hits[kMaxSlots];
  int cnt[kMaxSlots];
  unsigned char perm[kMaxSlots], idx[kMaxSlots];
  for (int i = 0; i < n1; ++i) merged[i] = L.t[i];
  for (int i = 0; i < R.n; ++i) merged[n1 + i] = R.t[i];
  lit_argsort(merged, m, perm);
  int run = 0;
  for (int k = 0; k < m; ++k) {
    const int odd = perm[k] & 1;
    const int flip = (op == PRT_DIFFERENCE) ? (odd ^ (perm[k] >= n1 ? 1 : 0)) : odd;
    run += flip ? -1 : 1;
    cnt[k] = run + (op == PRT_DIFFERENCE ? 1 : 0);
  }
  for (int k = 0; k < m; ++k) {
    const int prev = cnt[(k + m - 1) % m];
    const bool keep = (op == PRT_UNION) ? ((cnt[k] != 0) != (prev != 0)) : ((cnt[k] == 2) || (prev == 2));
    hits[k] = keep ? merged[perm[k]] : PRT_INF;
  }
  lit_argsort(hits, m, idx);
  out.n = m;
  for (int k = 0; k < m; ++k) {
    const int src = perm[idx[k]];
    out.t[k] = hits[idx[k]];
    out.leaf[k] = (src < n1) ? L.leaf[src] : R.leaf[src - n1];
  }
}

PRT_HD void lit_leaf(const SceneView& sc, int leaf, double p0, double p1, double p2, double v0, double v1, double v2,
                     LitList& out) {
  double t0, t1;
  leaf_hits(sc.leaves[leaf], p0, p1, p2, v0, v1, v2, t0, t1);
  out.n = 2;
  out.t[0] = t0;
  out.t[1] = t1;
  out.leaf[0] = out.leaf[1] = (short)leaf;
}

// `stack` needs kMaxDepth + 1 lists; the result is left in stack[0]
PRT_HD void eval_component_literal(const SceneView& sc, int begin, int end, double p0, double p1, double p2,
                                   double v0, double v1, double v2, const RayInv& inv, LitList* stack) {
  int sp = 0;
  int pc = begin;
  while (pc < end) {
    const Op op = sc.ops[pc];
    if (op.kind == OP_ENTER) {
      double b0, b1;
      cube_hits(sc.aabb + 6 * op.a, p0, p1, p2, v0, v1, v2, inv, b0, b1);
      if (!(b0 < PRT_INF)) {  // csg.py:126-133: the whole node reports +inf hits with surface id -1
        int leaves = 0;
        for (int q = pc + 1; q < op.b; ++q) leaves += (sc.ops[q].kind == OP_LEAF) | (sc.ops[q].kind == OP_MERGE_LEAF);
        LitList& o = stack[sp++];
        o.n = 2 * leaves;
        for (int k = 0; k < o.n; ++k) {
          o.t[k] = PRT_INF;
          o.leaf[k] = -1;
        }
        pc = op.b;
        continue;
      }
    } else if (op.kind == OP_LEAF) {
      lit_leaf(sc, op.a, p0, p1, p2, v0, v1, v2, stack[sp++]);
    } else if (op.kind == OP_MERGE_LEAF) {
      LitList& r = stack[sp];
      lit_leaf(sc, op.b, p0, p1, p2, v0, v1, v2, r);
      LitList& res = stack[sp + 1];
      lit_merge(stack[sp - 1], r, op.a, res);
      stack[sp - 1] = res;
    } else {  // OP_MERGE
      LitList& res = stack[sp];
      lit_merge(stack[sp - 2], stack[sp - 1], op.a, res);
      stack[sp - 2] = res;
      --sp;
    }
    ++pc;
  }
}

constexpr int kLitStack = kMaxDepth + 2;

}  // namespace prt
