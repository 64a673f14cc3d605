// Host emulation of the trace kernel's per-ray code (TEST/DEBUG ONLY, never shipped).
//
// Includes pyrayt_b200/csrc/prt_device.cuh (the PRT_HD functions the CUDA kernels
// call) and runs them on the CPU so kernel logic can be checked against the oracle
// in the build container, which has no GPU.  The product library does not contain
// or call this; `pytest -m "not gpu"` builds it into tests/emul/libprt_emul.so.
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

using std::isinf;
using std::isnan;

#include "../../pyrayt_b200/csrc/prt_device.cuh"
#include "../../pyrayt_b200/csrc/prt_encode.h"

extern "C" {

// frame: row-major scratch (rows ray-major, 15 per row) up to cap rows; nrows[i] rows per ray.
// returns total rows or <0.
long long prt_emul_trace(const prt_scene_desc* d, const double* rays, long long n, long long stride,
                         int generation_limit, double ray_offset, double* rows_out, long long cap,
                         int* nrows, unsigned long long* counters) {
  std::vector<unsigned char> blob;
  std::vector<int> slots;
  std::string err;
  if (prt::encode_scene(d, blob, slots, err) != PRT_OK) return -1;
  const prt::SceneView sc = prt::make_view(blob.data());
  long long total = 0;
  prt::StepCounters c = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long tie_rays = 0;
  for (long long i = 0; i < n; ++i) {
    prt::RayState r;
    r.p0 = rays[0 * stride + i]; r.p1 = rays[1 * stride + i]; r.p2 = rays[2 * stride + i];
    r.v0 = rays[4 * stride + i]; r.v1 = rays[5 * stride + i]; r.v2 = rays[6 * stride + i];
    r.gen = rays[8 * stride + i]; r.inten = rays[9 * stride + i]; r.wl = rays[10 * stride + i];
    r.nidx = rays[11 * stride + i]; r.id = rays[12 * stride + i];
    prt::HitStack S;
    int k = 0;
    c.tie = 0;
    for (int g = 0; g < generation_limit; ++g) {
      prt::StepOut o;
      const bool on = prt::trace_step(sc, r, g, generation_limit, S, o, c);
      if (o.row) {
        if (total < cap) {
          double* w = rows_out + total * 15;
          w[0] = r.gen; w[1] = r.inten; w[2] = r.wl; w[3] = r.nidx; w[4] = r.id; w[5] = o.sid;
          w[6] = r.p0; w[7] = r.p1; w[8] = r.p2; w[9] = o.e0; w[10] = o.e1; w[11] = o.e2;
          w[12] = o.t0n; w[13] = o.t1n; w[14] = o.t2n;
        }
        ++total;
        ++k;
      }
      if (!on) break;
      prt::advance_ray(r, o, g, ray_offset);
    }
    tie_rays += c.tie;
    nrows[i] = k;
  }
  counters[0] = (unsigned long long)n;
  counters[1] = c.gen;
  counters[2] = c.seg;
  counters[3] = tie_rays;
  counters[4] = c.untr;
  counters[5] = c.nan;
  counters[6] = c.lim;
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
    prt::HitStack S;
    S.flags = 0;
    bool tie = false;
    const bool any = prt::eval_component(
        sc, sc.comp[comp], sc.comp[comp + 1], rays[0 * n + i], rays[1 * n + i], rays[2 * n + i], rays[4 * n + i],
        rays[5 * n + i], rays[6 * n + i],
        prt::make_ray_inv(rays[0 * n + i], rays[1 * n + i], rays[2 * n + i], rays[4 * n + i], rays[5 * n + i],
                          rays[6 * n + i], (sc.h->flags & 1) != 0),
        false, INFINITY, S, tie);
    const int b = prt::buf_of(S, 0);
    const int len = any ? S.len[0] : 0;
    for (int k = 0; k < m; ++k) {
      hits[k * n + i] = (k < len) ? S.t[b][k] : INFINITY;
      sids[k * n + i] = (k < len) ? (long long)sc.leaves[S.leaf[b][k]].sid : -1;
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
    const prt::Op first = sc.ops[sc.comp[c]];
    flags[c] = (first.kind == prt::OP_ENTER) ? (first.c & 1) : -1;
  }
  return 0;
}
}
