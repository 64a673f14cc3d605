// Host-side scene encoder: caller's postfix CSG programs (prt_scene_desc) -> the blob the
// kernels stage in shared memory (prt_scene.h).  Shared by the ABI layer and tests/emul.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pyrayt_b200.h"
#include "prt_scene.h"

namespace prt {

struct TreeNode {
  int kind;  // prt_node_kind
  int leaf;
  int node;  // index into the caller's node arrays (for the aabb)
  int l = -1, r = -1;
  int slots = 2;
};

// axis-aligned box used to prove that a caller-supplied bounding box contains its solid
struct Box3 {
  double lo[3], hi[3];
  bool empty;
};

// world-space bounds of leaf l: the primitive's object-space bounding cube (primitives.py
// bounding_points :225,:312,:431,:512,:635) pushed through the inverse of the object matrix
inline bool leaf_world_box(const prt_scene_desc* d, int l, Box3& out) {
  const double* M = d->leaf_obj + 16 * l;
  const double* p = d->leaf_param + 6 * l;
  double lo[3], hi[3];
  switch (d->leaf_type[l]) {
    case PRT_SPHERE: for (int k = 0; k < 3; ++k) { lo[k] = -std::fabs(p[0]); hi[k] = std::fabs(p[0]); } break;
    case PRT_PARABOLOID: {
      const double r = std::sqrt(4 * p[0] * p[1]);
      lo[0] = lo[1] = -r; hi[0] = hi[1] = r; lo[2] = 0; hi[2] = p[1];
    } break;
    case PRT_PLANE: lo[0] = -p[0] / 2; hi[0] = p[0] / 2; lo[1] = -p[1] / 2; hi[1] = p[1] / 2; lo[2] = hi[2] = 0; break;
    case PRT_CUBE: for (int k = 0; k < 3; ++k) { lo[k] = p[2 * k]; hi[k] = p[2 * k + 1]; } break;
    case PRT_CYLINDER: lo[0] = lo[1] = -std::fabs(p[0]); hi[0] = hi[1] = std::fabs(p[0]); lo[2] = p[1]; hi[2] = p[2]; break;
    default: return false;
  }
  // invert the affine object matrix: W = A^-1, t_w = -A^-1 t
  const double a = M[0], b = M[1], c = M[2], e = M[4], f = M[5], g = M[6], h = M[8], i = M[9], j = M[10];
  const double det = a * (f * j - g * i) - b * (e * j - g * h) + c * (e * i - f * h);
  if (!(std::fabs(det) > 1e-300) || !std::isfinite(det)) return false;
  const double W[9] = {(f * j - g * i) / det, (c * i - b * j) / det, (b * g - c * f) / det,
                       (g * h - e * j) / det, (a * j - c * h) / det, (c * e - a * g) / det,
                       (e * i - f * h) / det, (b * h - a * i) / det, (a * f - b * e) / det};
  const double t[3] = {M[3], M[7], M[11]};
  out.empty = false;
  for (int k = 0; k < 3; ++k) { out.lo[k] = INFINITY; out.hi[k] = -INFINITY; }
  for (int corner = 0; corner < 8; ++corner) {
    const double q[3] = {((corner & 1) ? hi[0] : lo[0]) - t[0], ((corner & 2) ? hi[1] : lo[1]) - t[1],
                         ((corner & 4) ? hi[2] : lo[2]) - t[2]};
    for (int k = 0; k < 3; ++k) {
      const double w = W[3 * k] * q[0] + W[3 * k + 1] * q[1] + W[3 * k + 2] * q[2];
      if (!std::isfinite(w)) return false;
      if (w < out.lo[k]) out.lo[k] = w;
      if (w > out.hi[k]) out.hi[k] = w;
    }
  }
  return true;
}

struct Encoder {
  const prt_scene_desc* s;
  std::vector<TreeNode> tree;
  std::vector<prt::Op> ops;
  std::vector<double> aabb;
  int max_depth = 0;

  // emit the preorder program of subtree t; returns the live-list depth it needs
  int emit(int t, int depth_in) {
    const TreeNode& n = tree[t];
    if (n.kind == PRT_LEAF) {
      ops.push_back({prt::OP_LEAF, n.leaf, 0, 0});
      return depth_in + 1;
    }
    const int enter = (int)ops.size();
    const int box = (int)aabb.size() / 6;
    for (int k = 0; k < 6; ++k) aabb.push_back(s->node_aabb[6 * n.node + k]);
    ops.push_back({prt::OP_ENTER, box, 0, 0});
    int d = emit(n.l, depth_in);
    if (d > max_depth) max_depth = d;
    if (tree[n.r].kind == PRT_LEAF) {
      ops.push_back({prt::OP_MERGE_LEAF, n.kind, tree[n.r].leaf, 0});
    } else {
      d = emit(n.r, depth_in + 1);
      if (d > max_depth) max_depth = d;
      ops.push_back({prt::OP_MERGE, n.kind, 0, 0});
    }
    ops[enter].b = (int)ops.size();  // skip target: first op after this node
    if (depth_in + 1 > max_depth) max_depth = depth_in + 1;
    return depth_in + 1;
  }

  // bounds of everything the reference's CSG evaluation of subtree t can report as a hit:
  // INTERSECT -> overlap of the children, DIFFERENCE -> the left child, UNION -> hull
  bool solid_box(int t, Box3& out) const {
    const TreeNode& n = tree[t];
    if (n.kind == PRT_LEAF) return leaf_world_box(s, n.leaf, out);
    Box3 l, r;
    if (!solid_box(n.l, l) || !solid_box(n.r, r)) return false;
    if (n.kind == PRT_DIFFERENCE) {
      out = l;
    } else if (n.kind == PRT_INTERSECT) {
      out.empty = l.empty || r.empty;
      for (int k = 0; k < 3; ++k) {
        out.lo[k] = l.lo[k] > r.lo[k] ? l.lo[k] : r.lo[k];
        out.hi[k] = l.hi[k] < r.hi[k] ? l.hi[k] : r.hi[k];
        if (out.lo[k] > out.hi[k]) out.empty = true;
      }
    } else {
      if (l.empty) { out = r; return true; }
      if (r.empty) { out = l; return true; }
      out.empty = false;
      for (int k = 0; k < 3; ++k) {
        out.lo[k] = l.lo[k] < r.lo[k] ? l.lo[k] : r.lo[k];
        out.hi[k] = l.hi[k] > r.hi[k] ? l.hi[k] : r.hi[k];
      }
    }
    return true;
  }

  // true when the caller's bounding box of root node t contains solid_box(t) (to rounding):
  // only then may the kernel prune the component by its box (prt_device.cuh eval_component)
  bool root_box_is_bound(int t) const {
    Box3 b;
    if (tree[t].kind == PRT_LEAF || !solid_box(t, b)) return false;
    if (b.empty) return true;
    const double* a = s->node_aabb + 6 * tree[t].node;
    for (int k = 0; k < 3; ++k) {
      const double tol = 1e-12 * (1.0 + std::fabs(b.lo[k]) + std::fabs(b.hi[k]));
      if (!(a[2 * k] <= b.lo[k] + tol) || !(a[2 * k + 1] >= b.hi[k] - tol)) return false;
    }
    return true;
  }
};


// returns PRT_OK or a negative prt_status with `err` set
inline int encode_scene(const prt_scene_desc* d, std::vector<unsigned char>& blob, std::vector<int>& comp_slots,
                        std::string& err) {
  auto fail = [&](int code, const char* msg) {
    err = msg;
    return code;
  };
  if (d->n_components < 0 || d->n_nodes < 0 || d->n_leaves < 0) return fail(PRT_ERR_INVALID, "negative counts");
  if (d->n_leaves > PRT_MAX_LEAVES || d->n_nodes > PRT_MAX_NODES)
    return fail(PRT_ERR_LIMIT, "scene exceeds PRT_MAX_LEAVES / PRT_MAX_NODES");

  Encoder enc;
  enc.s = d;
  std::vector<int> comp_begin(d->n_components + 1, 0);
  std::vector<prt::Comp> comps;
  comp_slots.assign(d->n_components, 0);
  int max_slots = 2;
  for (int c = 0; c < d->n_components; ++c) {
    const int b = d->comp_node_begin[c], e = d->comp_node_begin[c + 1];
    if (b < 0 || e > d->n_nodes || b >= e) return fail(PRT_ERR_INVALID, "bad component node range");
    std::vector<int> stack;
    for (int nd = b; nd < e; ++nd) {
      TreeNode t;
      t.kind = d->node_kind[nd];
      t.node = nd;
      t.leaf = d->node_leaf[nd];
      if (t.kind == PRT_LEAF) {
        if (t.leaf < 0 || t.leaf >= d->n_leaves) return fail(PRT_ERR_INVALID, "leaf index out of range");
      } else if (t.kind == PRT_UNION || t.kind == PRT_INTERSECT || t.kind == PRT_DIFFERENCE) {
        if (stack.size() < 2) return fail(PRT_ERR_INVALID, "malformed postfix program");
        t.r = stack.back();
        stack.pop_back();
        t.l = stack.back();
        stack.pop_back();
        t.slots = enc.tree[t.l].slots + enc.tree[t.r].slots;
      } else {
        return fail(PRT_ERR_INVALID, "unknown node kind");
      }
      enc.tree.push_back(t);
      stack.push_back((int)enc.tree.size() - 1);
    }
    if (stack.size() != 1) return fail(PRT_ERR_INVALID, "malformed postfix program");
    const int root = stack[0];
    comp_slots[c] = enc.tree[root].slots;
    if (comp_slots[c] > PRT_MAX_SLOTS) return fail(PRT_ERR_LIMIT, "component exceeds PRT_MAX_SLOTS hit slots");
    if (comp_slots[c] > max_slots) max_slots = comp_slots[c];
    comp_begin[c] = (int)enc.ops.size();
    enc.emit(root, 0);
    {
      const prt::Op* o = enc.ops.data() + comp_begin[c];
      const int len = (int)enc.ops.size() - comp_begin[c];
      int shape = prt::SHAPE_GENERIC;
      if (len == 1 && o[0].kind == prt::OP_LEAF) shape = prt::SHAPE_LEAF;
      else if (len == 3 && o[0].kind == prt::OP_ENTER && o[1].kind == prt::OP_LEAF && o[2].kind == prt::OP_MERGE_LEAF)
        shape = prt::SHAPE_LEFT2;
      else if (len == 5 && o[0].kind == prt::OP_ENTER && o[1].kind == prt::OP_ENTER && o[2].kind == prt::OP_LEAF &&
               o[3].kind == prt::OP_MERGE_LEAF && o[4].kind == prt::OP_MERGE_LEAF)
        shape = prt::SHAPE_LEFT3;
      prt::Comp C;
      std::memset(&C, 0, sizeof C);
      C.shape = shape;
      C.begin = comp_begin[c];
      C.end = (int)enc.ops.size();
      C.leaf_a = C.leaf_b = C.leaf_c = -1;
      if (shape == prt::SHAPE_LEAF) {
        C.leaf_a = o[0].a;
        // a bare surface has no box in the reference; the encoder gives it one for the quick prune only:
        // the world box of the primitive, inflated by 1e-6 of its size (hit points carry rounding of their own)
        Box3 b;
        if (leaf_world_box(d, C.leaf_a, b)) {
          bool ok = true;
          for (int k = 0; k < 3; ++k) {
            const double pad = 1e-6 * (1.0 + std::fmax(std::fabs(b.lo[k]), std::fabs(b.hi[k])));
            C.root_box[2 * k] = b.lo[k] - pad;
            C.root_box[2 * k + 1] = b.hi[k] + pad;
            if (!std::isfinite(C.root_box[2 * k]) || !std::isfinite(C.root_box[2 * k + 1])) ok = false;
          }
          if (ok) C.flags |= 4;  // bit 2: root_box bounds this bare leaf
        }
      }
      if (shape == prt::SHAPE_GENERIC && o[0].kind == prt::OP_ENTER)  // the traversal tables need the root box
        for (int k = 0; k < 6; ++k) C.root_box[k] = enc.aabb[6 * o[0].a + k];
      if (shape == prt::SHAPE_LEFT2) {
        C.leaf_a = o[1].a;
        C.leaf_b = o[2].b;
        C.op1 = o[2].a;
        for (int k = 0; k < 6; ++k) C.root_box[k] = enc.aabb[6 * o[0].a + k];
      }
      if (shape == prt::SHAPE_LEFT3) {
        C.leaf_a = o[2].a;
        C.leaf_b = o[3].b;
        C.leaf_c = o[4].b;
        C.op1 = o[3].a;
        C.op2 = o[4].a;
        for (int k = 0; k < 6; ++k) C.root_box[k] = enc.aabb[6 * o[0].a + k];
        for (int k = 0; k < 6; ++k) C.inner_box[k] = enc.aabb[6 * o[1].a + k];
      }
      if (shape == prt::SHAPE_LEFT2 || shape == prt::SHAPE_LEFT3) {
        auto apply = [](int op, int x, int y) { return op == PRT_UNION ? (x | y) : (op == PRT_INTERSECT ? (x & y) : (x & (y ^ 1))); };
        for (int bits = 0; bits < 8; ++bits) {
          const int f1 = apply(C.op1, bits & 1, (bits >> 1) & 1);
          const int f = (shape == prt::SHAPE_LEFT3) ? apply(C.op2, f1, (bits >> 2) & 1) : f1;
          C.tt |= f << bits;
        }
      }
      if (shape == prt::SHAPE_LEFT3) {
        const int ta = d->leaf_type[C.leaf_a], tb = d->leaf_type[C.leaf_b], tc = d->leaf_type[C.leaf_c];
        if (ta == PRT_CYLINDER && tb == PRT_SPHERE && tc == PRT_SPHERE) C.flags |= 8;
        if (ta == PRT_SPHERE && tb == PRT_SPHERE && tc == PRT_CYLINDER) C.flags |= 16;
      }
      // convex solids: intersections of convex primitives (every leaf type but the flat Plane)
      if (shape == prt::SHAPE_LEFT2 || shape == prt::SHAPE_LEFT3) {
        bool convex = C.op1 == PRT_INTERSECT && (shape == prt::SHAPE_LEFT2 || C.op2 == PRT_INTERSECT);
        const int lv[3] = {C.leaf_a, C.leaf_b, C.leaf_c};
        for (int k = 0; k < (shape == prt::SHAPE_LEFT3 ? 3 : 2); ++k)
          if (d->leaf_type[lv[k]] == PRT_PLANE || d->leaf_nscale[lv[k]] != 1.0) convex = false;
        if (convex) C.flags |= 2;
      }
      comps.push_back(C);
    }
    if (enc.tree[root].kind != PRT_LEAF && enc.root_box_is_bound(root)) {
      enc.ops[comp_begin[c]].c |= 1;
      comps.back().flags |= 1;
    }
  }
  comp_begin[d->n_components] = (int)enc.ops.size();
  if (enc.max_depth > prt::kMaxDepth) return fail(PRT_ERR_LIMIT, "CSG tree nests too deeply on the right");

  for (int l = 0; l < d->n_leaves; ++l) {
    const int ty = d->leaf_type[l];
    if (ty < PRT_SPHERE || ty > PRT_CYLINDER) return fail(PRT_ERR_INVALID, "unknown primitive type");
    const int m = d->leaf_mat[l];
    if (m < PRT_MAT_ABSORBER || m > PRT_MAT_UNTRACEABLE) return fail(PRT_ERR_INVALID, "unknown material kind");
    const double* M = d->leaf_obj + 16 * l;
    if (M[12] != 0.0 || M[13] != 0.0 || M[14] != 0.0 || M[15] != 1.0)
      return fail(PRT_ERR_UNSUPPORTED, "object transform is not affine");
  }

  // pack the blob
  prt::BlobHeader h;
  std::memset(&h, 0, sizeof h);
  h.n_components = d->n_components;
  h.n_ops = (int)enc.ops.size();
  h.n_leaves = d->n_leaves;
  h.n_aabb = (int)enc.aabb.size() / 6;
  h.max_slots = max_slots;
  {
    bool tame = true;
    for (double v : enc.aabb) {
      const double av = std::fabs(v);
      if (!(v == 0.0 || (av >= 0x1p-823 && av < 0x1p677))) tame = false;
    }
    h.flags = tame ? 1 : 0;
    bool small = true;  // every root box within +-1e6: the dominant-axis quick prune's rounding stays below its margin
    for (const prt::Comp& C : comps) {
      if (C.shape == prt::SHAPE_GENERIC) h.flags |= 2;
      if (C.flags & 5)  // (only proven boxes are pruned)
        for (double v : C.root_box)
          if (!(std::fabs(v) <= 1e6)) small = false;
    }
    if (small) h.flags |= 4;

  }
  auto align8 = [](int x) { return (x + 7) & ~7; };
  int off = align8((int)sizeof(prt::BlobHeader));
  h.off_comps = off;
  off = align8(off + (int)sizeof(prt::Comp) * (d->n_components > 0 ? d->n_components : 1));
  h.off_ops = off;
  off = align8(off + (int)sizeof(prt::Op) * h.n_ops);
  h.off_aabb = off;
  off = align8(off + (int)sizeof(double) * 6 * h.n_aabb);
  h.off_leaves = off;
  off = align8(off + (int)sizeof(prt::Leaf) * d->n_leaves);
  // traversal tables (see OrderEntry): only meaningful when the quick prune is allowed at all (flags bit 2)
  std::vector<int> boxed, unboxed;
  for (int c = 0; c < d->n_components; ++c) (((comps[c].flags & 5) && (h.flags & 4)) ? boxed : unboxed).push_back(c);
  h.n_boxed = (int)boxed.size();
  h.n_unboxed = (int)unboxed.size();
  off = (off + 15) & ~15;
  h.off_order = off;
  off += (int)sizeof(prt::OrderEntry) * 6 * h.n_boxed;
  h.off_unboxed = off;
  off = (off + (int)sizeof(int) * h.n_unboxed + 15) & ~15;
  h.off_bycomp = off;
  off += (int)sizeof(prt::OrderEntry) * 6 * (d->n_components > 0 ? d->n_components : 1);
  if (h.n_boxed > prt::kOrderedMinBoxed) h.flags |= 8;
  // experiments only: PRT_FORCE_TRAVERSAL=ordered|list overrides the choice (both give the same frame)
  if (const char* force = std::getenv("PRT_FORCE_TRAVERSAL")) {
    if (force[0] == 'o' && h.n_boxed > 0) h.flags |= 8;
    if (force[0] == 'l') h.flags &= ~8;
  }
  h.total_bytes = off;
  blob.assign((size_t)off, 0);
  std::memcpy(blob.data(), &h, sizeof h);
  if (!comps.empty()) std::memcpy(blob.data() + h.off_comps, comps.data(), sizeof(prt::Comp) * comps.size());
  if (h.n_ops) std::memcpy(blob.data() + h.off_ops, enc.ops.data(), sizeof(prt::Op) * enc.ops.size());
  if (h.n_aabb) std::memcpy(blob.data() + h.off_aabb, enc.aabb.data(), sizeof(double) * enc.aabb.size());
  prt::Leaf* leaves = reinterpret_cast<prt::Leaf*>(blob.data() + h.off_leaves);
  for (int l = 0; l < d->n_leaves; ++l) {
    prt::Leaf& L = leaves[l];
    for (int k = 0; k < 12; ++k) L.m[k] = d->leaf_obj[16 * l + k];
    for (int k = 0; k < 6; ++k) L.prm[k] = d->leaf_param[6 * l + k];
    for (int k = 0; k < 6; ++k) L.matp[k] = d->leaf_matp[6 * l + k];
    L.nscale = d->leaf_nscale[l];
    L.sid = (double)d->leaf_sid[l];
    L.type = d->leaf_type[l];
    L.mat = d->leaf_mat[l];
    L.comp = -1;
    L.pad = 0;
  }

  for (int c = 0; c < d->n_components; ++c)
    for (int nd = d->comp_node_begin[c]; nd < d->comp_node_begin[c + 1]; ++nd)
      if (d->node_kind[nd] == PRT_LEAF) leaves[d->node_leaf[nd]].comp = c;

  prt::OrderEntry* order = reinterpret_cast<prt::OrderEntry*>(blob.data() + h.off_order);
  for (int k = 0; k < 3; ++k)
    for (int sgn = 0; sgn < 2; ++sgn) {
      prt::OrderEntry* tab = order + (size_t)(2 * k + sgn) * h.n_boxed;
      for (int j = 0; j < h.n_boxed; ++j) {
        const prt::Comp& C = comps[boxed[j]];
        tab[j].near_u = sgn ? -C.root_box[2 * k + 1] : C.root_box[2 * k];
        tab[j].far_u = sgn ? -C.root_box[2 * k] : C.root_box[2 * k + 1];
        tab[j].comp = boxed[j];
        tab[j].pad = 0;
      }
      // ties in near_u keep component order (any order is correct: nearest_hit breaks ties by index)
      std::stable_sort(tab, tab + h.n_boxed,
                       [](const prt::OrderEntry& a, const prt::OrderEntry& b) { return a.near_u < b.near_u; });
      double run = -INFINITY;
      for (int j = 0; j < h.n_boxed; ++j) {
        run = tab[j].far_u > run ? tab[j].far_u : run;
        tab[j].pmfar_u = run;
      }
    }
  if (h.n_unboxed) std::memcpy(blob.data() + h.off_unboxed, unboxed.data(), sizeof(int) * unboxed.size());
  prt::OrderEntry* bycomp = reinterpret_cast<prt::OrderEntry*>(blob.data() + h.off_bycomp);
  for (int t = 0; t < 6; ++t)
    for (int c = 0; c < d->n_components; ++c) {
      prt::OrderEntry& e = bycomp[(size_t)t * d->n_components + c];
      const prt::Comp& C = comps[c];
      const bool has_box = (C.flags & 5) && (h.flags & 4);
      const int k = t >> 1, sgn = t & 1;
      e.near_u = has_box ? (sgn ? -C.root_box[2 * k + 1] : C.root_box[2 * k]) : -INFINITY;
      e.far_u = has_box ? (sgn ? -C.root_box[2 * k] : C.root_box[2 * k + 1]) : INFINITY;
      e.pmfar_u = INFINITY;
      e.comp = c;
      e.pad = 0;
    }
  return PRT_OK;
}

}  // namespace prt
