// Host-side scene encoder: caller's postfix CSG programs (prt_scene_desc) -> the blob the
// kernels stage in shared memory (prt_scene.h).  Shared by the ABI layer and tests/emul.
#pragma once
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
  if (d->n_leaves > 255) return fail(PRT_ERR_LIMIT, "leaf index must fit one byte");

  Encoder enc;
  enc.s = d;
  std::vector<int> comp_begin(d->n_components + 1, 0);
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
  auto align8 = [](int x) { return (x + 7) & ~7; };
  int off = align8((int)sizeof(prt::BlobHeader));
  h.off_comp = off;
  off = align8(off + (int)sizeof(int) * (d->n_components + 1));
  h.off_ops = off;
  off = align8(off + (int)sizeof(prt::Op) * h.n_ops);
  h.off_aabb = off;
  off = align8(off + (int)sizeof(double) * 6 * h.n_aabb);
  h.off_leaves = off;
  off = align8(off + (int)sizeof(prt::Leaf) * d->n_leaves);
  h.total_bytes = off;
  blob.assign((size_t)off, 0);
  std::memcpy(blob.data(), &h, sizeof h);
  std::memcpy(blob.data() + h.off_comp, comp_begin.data(), sizeof(int) * comp_begin.size());
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
  }

  return PRT_OK;
}

}  // namespace prt
