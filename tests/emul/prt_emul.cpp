// Host emulation of the trace kernel's per-ray code (TEST/DEBUG ONLY, never shipped).
//
// Includes pyrayt_b200/csrc/prt_device.cuh (the PRT_HD functions the CUDA kernels
// call) and runs them on the CPU so kernel logic can be checked against the oracle
// in the build container, which has no GPU.  The product library does not contain
// or call this; `pytest -m "not gpu"` builds it into tests/emul/libprt_emul.so.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

using std::isinf;
using std::isnan;

#include "../../pyrayt_b200/csrc/prt_device.cuh"
#include "../../pyrayt_b200/csrc/prt_literal.cuh"
#include "../../pyrayt_b200/csrc/prt_encode.h"
#include "../../pyrayt_b200/csrc/prt_device_f32.cuh"

static int g_diagnose = 0;

extern "C" {

// PRT_FLAG_DIAGNOSE for the following prt_emul_trace calls (counters[7] = grazing rays, [8] = seam rays)
void prt_emul_set_diagnose(int on) { g_diagnose = on; }

// frame: row-major scratch (rows ray-major, 15 per row) up to cap rows; nrows[i] rows per ray.
// returns total rows or <0.
long long prt_emul_trace(const prt_scene_desc* d, const double* rays, long long n, long long stride,
                         int generation_limit, double ray_offset, double* rows_out, long long cap,
                         int* nrows, unsigned long long* counters) {
  std::vector<unsigned char> blob;
  std::vector<int> slots;
  std::string err;
  if (prt::encode_scene(d, blob, slots, err) != PRT_OK) return -1;
  // PRT_EMUL_TRAVERSAL=ordered|list forces one of nearest_hit's two traversals (the encoder picks by scene size)
  if (const char* force = std::getenv("PRT_EMUL_TRAVERSAL")) {
    prt::BlobHeader* h = reinterpret_cast<prt::BlobHeader*>(blob.data());
    if (force[0] == 'o') h->flags |= 8;
    if (force[0] == 'l') h->flags &= ~8;
  }
  const prt::SceneView sc = prt::make_view(blob.data());
  long long total = 0;
  prt::StepCounters c = {0, 0};
  unsigned long long tie_rays = 0, gens = 0, segs = 0, untr = 0, nans = 0, lims = 0, graze = 0, seam = 0;
  for (long long i = 0; i < n; ++i) {
    prt::RayState r;
    r.skip = -1;
    r.p0 = rays[0 * stride + i]; r.p1 = rays[1 * stride + i]; r.p2 = rays[2 * stride + i];
    r.v0 = rays[4 * stride + i]; r.v1 = rays[5 * stride + i]; r.v2 = rays[6 * stride + i];
    const double gen0 = rays[8 * stride + i], inten = rays[9 * stride + i], id = rays[12 * stride + i];
    r.wl = rays[10 * stride + i];
    r.nidx = rays[11 * stride + i];
    prt::HitStack S;
    int k = 0;
    c.w1 &= ~prt::kCtrTie;
    for (int g = 0; g < generation_limit; ++g) {
      prt::StepOut o;
      if (g_diagnose) {  // as trace_kernel<.., DIAG = true>: the nearest hit, then four displaced searches
        const double vn = std::sqrt(r.v0 * r.v0 + r.v1 * r.v1 + r.v2 * r.v2);
        if (!prt::isz(vn) && !(std::isnan(r.v0) || std::isnan(r.v1) || std::isnan(r.v2))) {
          double bt;
          int bl;
          bool tie = false;
          prt::nearest_hit<true>(sc, r.p0, r.p1, r.p2, r.v0, r.v1, r.v2, r.skip, &S, bt, bl, tie);
          c.w1 |= prt::diagnose_generation<true>(sc, r.p0, r.p1, r.p2, r.v0, r.v1, r.v2, vn, r.skip, &S, bl);
        }
      }
      const bool on = prt::trace_step<true>(sc, r, g, generation_limit, &S, o, c);
      if (o.row) {
        if (total < cap) {
          double* w = rows_out + total * 15;
          w[0] = (g == 0) ? gen0 : (double)g; w[1] = inten; w[2] = r.wl; w[3] = r.nidx; w[4] = id; w[5] = o.sid;
          w[6] = r.p0; w[7] = r.p1; w[8] = r.p2; w[9] = o.e0; w[10] = o.e1; w[11] = o.e2;
          w[12] = o.t0n; w[13] = o.t1n; w[14] = o.t2n;
        }
        ++total;
        ++k;
      }
      if (!on) break;
      prt::advance_ray(r, o, g, ray_offset);
    }
    tie_rays += (c.w1 & prt::kCtrTie) ? 1 : 0;
    gens += c.w0 & 0xffffu;
    segs += c.w0 >> 16;
    untr += (c.w1 & prt::kCtrUntr) ? 1 : 0;
    nans += (c.w1 & prt::kCtrNan) ? 1 : 0;
    lims += (c.w1 & prt::kCtrLim) ? 1 : 0;
    graze += (c.w1 & prt::kCtrGraze) ? 1 : 0;
    seam += (c.w1 & prt::kCtrSeam) ? 1 : 0;
    c.w0 = 0;
    c.w1 = 0;
    nrows[i] = k;
  }
  counters[0] = (unsigned long long)n;
  counters[1] = gens;
  counters[2] = segs;
  counters[3] = tie_rays;
  counters[4] = untr;
  counters[5] = nans;
  counters[6] = lims;
  counters[7] = graze;
  counters[8] = seam;
  return total;
}

// The FP32 fast mode's per-ray code (prt_device_f32.cuh) on the host: same driver loop as trace_kernel_f32,
// rows expanded as gather_kernel_f32 does.  (The host compiler does not contract to FMA, so values differ
// from the device in the last bits; the mode's contract is a tolerance.)  Returns rows, -1 bad scene.
long long prt_emul_trace_f32(const prt_scene_desc* d, const double* rays, long long n, long long stride,
                             int generation_limit, double ray_offset, double* rows_out, long long cap, int* nrows) {
  std::vector<unsigned char> blob;
  std::vector<int> slots;
  std::string err;
  if (prt::encode_scene(d, blob, slots, err) != PRT_OK) return -1;
  const prt::SceneView sv = prt::make_view(blob.data());
  std::vector<prt::f32::LeafF> lf(sv.h->n_leaves);
  std::vector<prt::f32::CompF> cf(sv.h->n_components);
  for (int l = 0; l < sv.h->n_leaves; ++l) prt::f32::convert_leaf(sv.leaves[l], lf[l]);
  for (int c = 0; c < sv.h->n_components; ++c) prt::f32::convert_comp(sv.comps[c], cf[c]);
  std::vector<prt::f32::OrderEntryF> ordf(6 * (size_t)sv.h->n_boxed + 1);
  for (int e = 0; e < 6 * sv.h->n_boxed; ++e) prt::f32::convert_order(sv.order[e], ordf[e]);
  // PRT_EMUL_F32_LIST / PRT_EMUL_F32_ORDERED force one of the two walks (the kernel follows the encoder: flags bit 3)
  const bool walk = sv.h->n_boxed > 0 && (sv.h->flags & 4) && !std::getenv("PRT_EMUL_F32_LIST") &&
                    ((sv.h->flags & 8) || std::getenv("PRT_EMUL_F32_ORDERED"));
  std::vector<float> aabbf(6 * (size_t)sv.h->n_aabb + 1);
  for (int e = 0; e < 6 * sv.h->n_aabb; ++e)
    aabbf[e] = (e % 6 & 1) ? prt::f32::to_float_up(sv.aabb[e]) : prt::f32::to_float_dn(sv.aabb[e]);
  prt::f32::SceneViewF sc = {sv.h, sv.comps, sv.leaves, lf.data(), cf.data(), walk ? ordf.data() : nullptr,
                             sv.unboxed, sv.ops, aabbf.data()};
  prt::f32::HitStackF S;
  long long total = 0;
  for (long long i = 0; i < n; ++i) {
    prt::f32::RayStateF r = {(float)rays[0 * stride + i], (float)rays[1 * stride + i], (float)rays[2 * stride + i],
                             (float)rays[4 * stride + i], (float)rays[5 * stride + i], (float)rays[6 * stride + i],
                             (float)rays[10 * stride + i], (float)rays[11 * stride + i], -1, -1};
    const double gen0 = rays[8 * stride + i], inten = rays[9 * stride + i], id = rays[12 * stride + i];
    const double wl = rays[10 * stride + i];
    prt::StepCounters c = {0, 0};
    int k = 0;
    for (int g = 0; g < generation_limit; ++g) {
      const float vn = prt::f32::step_speed(r, c);
      if (vn == 0.0f) break;
      float t;
      int leaf;
      bool tie = false;
      if (walk) prt::f32::nearest_hit<true, true>(sc, r, prt::f32::ray_scale(r), &S, t, leaf, tie);
      else prt::f32::nearest_hit<false, true>(sc, r, prt::f32::ray_scale(r), &S, t, leaf, tie);
      if (leaf < 0) break;
      if (lf[leaf].mat == PRT_MAT_UNTRACEABLE) break;
      if (total < cap) {
        double* w = rows_out + total * 15;
        const float rv = 1.0f / std::sqrt(r.v0 * r.v0 + r.v1 * r.v1 + r.v2 * r.v2);
        w[0] = (g == 0) ? gen0 : (double)g; w[1] = inten; w[2] = wl; w[3] = (double)r.nidx; w[4] = id;
        w[5] = sv.leaves[leaf].sid;
        w[6] = r.p0; w[7] = r.p1; w[8] = r.p2;
        w[9] = r.p0 + r.v0 * t; w[10] = r.p1 + r.v1 * t; w[11] = r.p2 + r.v2 * t;
        w[12] = r.v0 * rv; w[13] = r.v1 * rv; w[14] = r.v2 * rv;
      }
      ++total;
      ++k;
      prt::f32::StepOutF o;
      if (!prt::f32::step_interact(sc, r, g, generation_limit, vn, t, leaf, o, c)) break;
      prt::f32::advance_ray(r, o, leaf, (float)ray_offset);
    }
    nrows[i] = k;
  }
  return total;
}

int prt_emul_intersect(const prt_scene_desc* d, int comp, const double* rays, long long n, double* hits,
                       long long* sids) {
  std::vector<unsigned char> blob;
  std::vector<int> slots;
  std::string err;
  if (prt::encode_scene(d, blob, slots, err) != PRT_OK) return -1;
  const prt::SceneView sc = prt::make_view(blob.data());
  const int m = slots[comp];
  for (long long i = 0; i < n; ++i) {
    std::vector<prt::LitList> stack(prt::kLitStack);
    const double p0 = rays[0 * n + i], p1 = rays[1 * n + i], p2 = rays[2 * n + i];
    const double v0 = rays[4 * n + i], v1 = rays[5 * n + i], v2 = rays[6 * n + i];
    prt::eval_component_literal(sc, sc.comps[comp].begin, sc.comps[comp].end, p0, p1, p2, v0, v1, v2,
                                prt::make_ray_inv(p0, p1, p2, v0, v1, v2, (sc.h->flags & 1) != 0), stack.data());
    const prt::LitList& r = stack[0];
    // the streaming evaluation the trace kernel uses must agree on every finite slot
    prt::HitStack S;
    S.flags = 0;
    bool tie = false;
    const bool any = prt::eval_component(sc, sc.comps[comp].begin, sc.comps[comp].end, p0, p1, p2, v0, v1, v2,
                                         prt::make_ray_inv(p0, p1, p2, v0, v1, v2, (sc.h->flags & 1) != 0), false,
                                         INFINITY, S, tie);
    const int b = prt::buf_of(S, 0);
    const int len = any ? S.len[0] : 0;
    if (r.n != m) return -2;
    for (int k = 0; k < m; ++k) {
      const bool fin = r.t[k] < INFINITY;
      if (fin != (k < len)) return -3;
      if (fin && (S.t[b][k] != r.t[k] || (int)S.leaf[b][k] != (int)r.leaf[k])) return -4;
      hits[k * n + i] = r.t[k];
      sids[k * n + i] = r.leaf[k] >= 0 ? (long long)sc.leaves[r.leaf[k]].sid : -1;
    }
  }
  return 0;
}

// per component: 1 when the encoder proved the root box bounds the solid (pruning enabled)
int prt_emul_prune_flags(const prt_scene_desc* d, int* flags) {
  std::vector<unsigned char> blob;
  std::vector<int> slots;
  std::string err;
  if (prt::encode_scene(d, blob, slots, err) != PRT_OK) return -1;
  const prt::SceneView sc = prt::make_view(blob.data());
  for (int c = 0; c < sc.h->n_components; ++c) {
    flags[c] = (sc.comps[c].shape == prt::SHAPE_LEAF) ? -1 : (sc.comps[c].flags & 1);
  }
  return 0;
}

// Exhaustive cross-check of the closed-form left-deep evaluation (left_deep_first_hit with the encoder's
// truth tables) against the streaming merge_lists on small sorted pairs with ties, -inf and +inf
// entries: returns the number of disagreements in the selected first-positive hit (or a tie the
// streaming merge saw and the closed form did not report).
long long prt_emul_selfcheck_left_deep(long long* cases_out) {
  const double V[] = {-INFINITY, -2.0, -1.0, 1.0, 2.0, 3.0, INFINITY};
  const int nv = 7;
  long long bad = 0, cases = 0;
  auto apply = [](int op, int x, int y) { return op == PRT_UNION ? (x | y) : (op == PRT_INTERSECT ? (x & y) : (x & (y ^ 1))); };
  for (int op1 = 1; op1 <= 3; ++op1)
    for (int op2 = 1; op2 <= 3; ++op2) {
      unsigned tt2 = 0, tt3 = 0;  // as prt_encode.h builds Comp.tt for SHAPE_LEFT2 / SHAPE_LEFT3
      for (int bits = 0; bits < 8; ++bits) {
        const int f1 = apply(op1, bits & 1, (bits >> 1) & 1);
        tt2 |= (unsigned)f1 << bits;
        tt3 |= (unsigned)apply(op2, f1, (bits >> 2) & 1) << bits;
      }
      for (int ia0 = 0; ia0 < nv; ++ia0) for (int ia1 = ia0; ia1 < nv; ++ia1)
      for (int ib0 = 0; ib0 < nv; ++ib0) for (int ib1 = ib0; ib1 < nv; ++ib1)
      for (int ic0 = 0; ic0 < nv; ++ic0) for (int ic1 = ic0; ic1 < nv; ++ic1)
      for (int inner = 0; inner < 2; ++inner) {  // inner = 0: the box of (A op1 B) was missed
        const double a[2] = {inner ? V[ia0] : INFINITY, inner ? V[ia1] : INFINITY};
        const double b[2] = {inner ? V[ib0] : INFINITY, inner ? V[ib1] : INFINITY};
        const double c[2] = {V[ic0], V[ic1]};
        if (!inner && (ia0 | ia1 | ib0 | ib1)) continue;  // one representative of the missed-box case
        // streaming reference
        prt::HitStack S;
        S.flags = 0;
        bool tie = false;
        const int bf = prt::buf_of(S, 0);
        S.t[bf][0] = a[0]; S.t[bf][1] = a[1]; S.leaf[bf][0] = S.leaf[bf][1] = 10;
        S.len[0] = (a[0] < INFINITY) ? ((a[1] < INFINITY) ? 2 : 1) : 0;
        const int nb = (b[0] < INFINITY) ? ((b[1] < INFINITY) ? 2 : 1) : 0;
        const int ncn = (c[0] < INFINITY) ? ((c[1] < INFINITY) ? 2 : 1) : 0;
        prt::merge_lists(S, 0, op1, nb, [&](int j) { return b[j]; }, [&](int) { return 11; }, tie);
        double want2 = INFINITY; int wl2 = -1;
        { const int q = prt::buf_of(S, 0);
          for (int k = 0; k < S.len[0]; ++k) if (S.t[q][k] > 0) { want2 = S.t[q][k]; wl2 = S.leaf[q][k]; break; } }
        const bool tie2 = tie;
        prt::merge_lists(S, 0, op2, ncn, [&](int j) { return c[j]; }, [&](int) { return 12; }, tie);
        double want3 = INFINITY; int wl3 = -1;
        { const int q = prt::buf_of(S, 0);
          for (int k = 0; k < S.len[0]; ++k) if (S.t[q][k] > 0) { want3 = S.t[q][k]; wl3 = S.leaf[q][k]; break; } }
        ++cases;
        // closed form, shape 2: (A op1 B), C absent
        double ct; int cl; bool t2 = false;
        if (op2 == 1 && ic0 == 0 && ic1 == 0) {
          prt::left_deep_first_hit(tt2, a[0], a[1], b[0], b[1], (double)INFINITY, (double)INFINITY, 10, 11, -1, ct, cl, t2);
          if (!(ct == want2 && cl == wl2) || (tie2 && !t2)) ++bad;
        }
        // closed form, shape 3
        bool t3 = false;
        prt::left_deep_first_hit(tt3, a[0], a[1], b[0], b[1], c[0], c[1], 10, 11, 12, ct, cl, t3);
        if (!(ct == want3 && cl == wl3) || (tie && !t3)) ++bad;
      }
    }
  *cases_out = cases;
  return bad;
}
}
