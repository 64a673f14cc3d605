// pyrayt_b200 trace kernels for sm_100a (B200).
//
// K1 trace_kernel<RECORD, GENERIC>
//                      one thread = one ray, every generation in a persistent loop
//                      (replaces RayTracer._st_propagate/_st_interact and everything under
//                      them: pyrayt/_pyrayt.py:370-452).  GENERIC = false (all components bare
//                      leaves or left-deep trees) compiles the CSG interpreter out.
// K2 scan_runs_kernel / gen_offsets_kernel / gather_kernel
//                      put the staged rows in the reference's (generation, id) order
//                      (pyrayt/_pyrayt.py:186,:428-435).
// K3 source kernels    seeded synthetic sources.
// intersect_kernel     component.intersect(rays) (world_objects.py:360-383, csg.py:118-160).
//
// Arithmetic is IEEE float64 with no FMA contraction (the file is compiled with
// -fmad=false): NumPy never fuses, and the order of operations below follows the
// reference expressions so that results agree to rounding.  No tensor cores: the
// work is per-ray scalar math, not a contraction.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/pyrayt_b200.h"
#include "prt_scene.h"
#include "prt_device.cuh"
#include "prt_literal.cuh"

namespace prt {

// ---------------------------------------------------------------- K1: the trace kernel

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

#ifndef PRT_MIN_BLOCKS
#define PRT_MIN_BLOCKS 2
#endif
#ifdef PRT_TIES_ALWAYS  // experiments: count equal merge keys in every trace, as before round 2's last step
constexpr bool kTiesAlways = true;
#else
constexpr bool kTiesAlways = false;
#endif

template <bool RECORD, bool GENERIC, bool DIAG = false, bool GLOBAL = false>
__global__ void __launch_bounds__(kTileRays, PRT_MIN_BLOCKS) trace_kernel(const TraceArgs a) {
  extern __shared__ __align__(16) unsigned char s_blob[];
  __shared__ int s_wcount[kTileRays / 32];
  __shared__ long long s_base;
  // per-ray values that the nearest-hit search does not need live in shared memory, not in registers:
  // the wavelength / refractive index (interaction only) and the event counters.  (generation, intensity
  // and id are only ever copied into rows: the ordering pass reads them from the RaySet itself.)
  __shared__ double s_wl[kTileRays], s_nidx[kTileRays];
  __shared__ unsigned s_ctr0[kTileRays], s_ctr1[kTileRays];

  // stage the scene in shared memory once per block (GLOBAL: a scene too large for that is read in place,
  // through L1 / L2 -- the ordered traversal touches a few components per ray and generation)
  if (!GLOBAL) {
    const int words = a.blob_bytes / 8;
    const double* src = reinterpret_cast<const double*>(a.blob);
    double* dst = reinterpret_cast<double*>(s_blob);
    for (int w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
  }
  __syncthreads();
  const SceneView sc = make_view(GLOBAL ? a.blob : s_blob);

  const long long tile = blockIdx.x;
  const long long i = tile * kTileRays + threadIdx.x;
  const bool valid = i < a.n_rays;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;

  RayState rs = {0, 0, 0, 0, 0, 0, 0, 1, -1};
  unsigned c_drop = 0, c_badw = 0;
  s_ctr0[threadIdx.x] = 0;
  s_ctr1[threadIdx.x] = 0;
  if (valid) {
    const double* r = a.rays + i;
    rs.p0 = r[0 * a.stride];
    rs.p1 = r[1 * a.stride];
    rs.p2 = r[2 * a.stride];
    const double pw = r[3 * a.stride];
    rs.v0 = r[4 * a.stride];
    rs.v1 = r[5 * a.stride];
    rs.v2 = r[6 * a.stride];
    const double vw = r[7 * a.stride];
    s_wl[threadIdx.x] = r[10 * a.stride];
    s_nidx[threadIdx.x] = r[11 * a.stride];
    if (pw != 1.0 || vw != 0.0) c_badw = 1;
  }
  bool alive = valid;
  // hit lists for the generic interpreter only; the left-deep variant keeps everything in registers
  typename StackFor<GENERIC>::type stack_storage;
  HitStack* S = StackFor<GENERIC>::ptr(stack_storage);

  for (int g = 0; g < a.generation_limit; ++g) {
    // _st_propagate: nearest hit of every live ray
    double vn = 0.0, hit_t = 0.0;
    int hit_leaf = -1;
    if (alive) {
      StepCounters ctr = {s_ctr0[threadIdx.x], s_ctr1[threadIdx.x]};
      vn = step_speed(rs, ctr);
      if (vn != 0.0) {
        bool tie = false;
        // (equal-key detection in the closed-form merges is a diagnostic: compiled in under PRT_FLAG_DIAGNOSE)
        nearest_hit<GENERIC, DIAG || kTiesAlways>(sc, rs.p0, rs.p1, rs.p2, rs.v0, rs.v1, rs.v2, rs.skip, S, hit_t, hit_leaf, tie);
        if (tie) ctr.w1 |= kCtrTie;
        if (DIAG)  // PRT_FLAG_DIAGNOSE: four more searches from origins displaced by 1e-9
          ctr.w1 |= diagnose_generation<GENERIC>(sc, rs.p0, rs.p1, rs.p2, rs.v0, rs.v1, rs.v2, vn, rs.skip, S,
                                                 hit_leaf);
      }
      s_ctr0[threadIdx.x] = ctr.w0;
      s_ctr1[threadIdx.x] = ctr.w1;
    }
    // What the interaction will decide is already known from the leaf that was hit: a row is produced unless
    // the material cannot be traced, and the ray goes on unless it is absorbed or at the generation limit
    // (step_interact's return value, restated).  So the row -- which holds pre-interaction state only, see
    // kStageCols -- is reserved and written first, and the interaction then updates the ray in place.
    const int mat = (hit_leaf >= 0) ? sc.leaves[hit_leaf].mat : PRT_MAT_UNTRACEABLE;
    const bool has_row = (hit_leaf >= 0) & (mat != PRT_MAT_UNTRACEABLE);
    const bool next_alive = has_row & (mat != PRT_MAT_ABSORBER) & (g + 1 != a.generation_limit);

    int any_alive = 1;
    if (RECORD) {
      const bool write =
          has_row && (a.record_mode == PRT_RECORD_ALL || sc.leaves[hit_leaf].sid == a.detector_sid);
      // block-aggregated append: one reservation per tile and generation, rows in ray order
      const unsigned m = __ballot_sync(0xffffffffu, write);
      if (lane == 0) s_wcount[warp] = __popc(m);
      any_alive = __syncthreads_or(next_alive);
      int before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < kTileRays / 32; ++w) {
        const int cw = s_wcount[w];
        if (w < warp) before += cw;
        total += cw;
      }
      if (threadIdx.x == 0 && total > 0) {
        const long long base =
            (long long)atomicAdd(reinterpret_cast<unsigned long long*>(&a.ctr->rows_reserved),
                                 (unsigned long long)total);
        a.run_start[(long long)g * a.n_tiles + tile] = base;
        a.run_count[(long long)g * a.n_tiles + tile] = (base + total <= a.capacity) ? total : 0;
        s_base = base;
      }
      __syncthreads();
      if (write) {
        const long long run = s_base;
        if (run + total <= a.capacity) {
          const long long row = run + before + __popc(m & ((1u << lane) - 1u));
          // the staged record (kStageCols): pre-interaction state + what was hit; gather_kernel expands it
          double* o = a.stage + row;
          const long long cs = a.capacity;
          o[0 * cs] = rs.p0;
          o[1 * cs] = rs.p1;
          o[2 * cs] = rs.p2;
          o[3 * cs] = rs.v0;
          o[4 * cs] = rs.v1;
          o[5 * cs] = rs.v2;
          o[6 * cs] = hit_t;
          o[7 * cs] = s_nidx[threadIdx.x];
          o[8 * cs] = __longlong_as_double(((long long)threadIdx.x << 32) | (long long)hit_leaf);
        } else {
          c_drop++;
        }
      }
    }

    // _st_interact: material, new direction, offset start of the next generation
    alive = false;
    if (hit_leaf >= 0) {
      StepCounters ctr = {s_ctr0[threadIdx.x], s_ctr1[threadIdx.x]};
      StepOut so;
      rs.wl = s_wl[threadIdx.x];
      rs.nidx = s_nidx[threadIdx.x];
      alive = step_interact<false>(sc, rs, g, a.generation_limit, vn, hit_t, hit_leaf, so, ctr);
      s_ctr0[threadIdx.x] = ctr.w0;
      s_ctr1[threadIdx.x] = ctr.w1;
      if (alive) {
        advance_ray(rs, so, g, a.ray_offset);
        s_nidx[threadIdx.x] = rs.nidx;
      }
    }
    // a recording tile leaves together (its barriers), after the last interaction has fed the counters
    if (RECORD ? !any_alive : !alive) break;
  }

  // counters: warp-reduce, one atomic per warp and counter
  const StepCounters sc_ctr = {s_ctr0[threadIdx.x], s_ctr1[threadIdx.x]};
  unsigned long long vals[13] = {valid ? 1ull : 0ull,
                                 sc_ctr.w0 & 0xffffu,
                                 sc_ctr.w0 >> 16,
                                 c_drop,
                                 (sc_ctr.w1 & kCtrTie) ? 1ull : 0ull,
                                 (sc_ctr.w1 & kCtrUntr) ? 1ull : 0ull,
                                 c_badw,
                                 (sc_ctr.w1 & kCtrNan) ? 1ull : 0ull,
                                 (sc_ctr.w1 & kCtrLim) ? 1ull : 0ull,
                                 (sc_ctr.w1 & kCtrAbs) ? 1ull : 0ull,
                                 sc_ctr.w1 & 0xffffu,
                                 (sc_ctr.w1 & kCtrGraze) ? 1ull : 0ull,
                                 (sc_ctr.w1 & kCtrSeam) ? 1ull : 0ull};
  unsigned long long* dst[13] = {
      reinterpret_cast<unsigned long long*>(&a.ctr->rays),
      reinterpret_cast<unsigned long long*>(&a.ctr->generations),
      reinterpret_cast<unsigned long long*>(&a.ctr->segments),
      reinterpret_cast<unsigned long long*>(&a.ctr->rows_dropped),
      reinterpret_cast<unsigned long long*>(&a.ctr->tie_rays),
      reinterpret_cast<unsigned long long*>(&a.ctr->untraceable_hits),
      reinterpret_cast<unsigned long long*>(&a.ctr->bad_w),
      reinterpret_cast<unsigned long long*>(&a.ctr->nan_rays),
      reinterpret_cast<unsigned long long*>(&a.ctr->limit_rays),
      reinterpret_cast<unsigned long long*>(&a.ctr->absorber_segments),
      reinterpret_cast<unsigned long long*>(&a.ctr->mirror_segments),
      reinterpret_cast<unsigned long long*>(&a.ctr->grazing_rays),
      reinterpret_cast<unsigned long long*>(&a.ctr->seam_rays)};
#pragma unroll
  for (int q = 0; q < (DIAG ? 13 : 11); ++q) {
    const unsigned long long s = warp_sum(vals[q]);
    if (lane == 0 && s) atomicAdd(dst[q], s);
  }
}

// ---------------------------------------------------------------- K2: (generation, id) ordering

// one block per generation: exclusive scan of run_count over tiles -> run_base (relative), total
__global__ void __launch_bounds__(1024) scan_runs_kernel(const int* run_count, long long* run_base,
                                                         long long n_tiles, long long* gen_total) {
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  const int g = blockIdx.x;
  const int* cnt = run_count + (long long)g * n_tiles;
  long long* base = run_base + (long long)g * n_tiles;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long t0 = 0; t0 < n_tiles; t0 += blockDim.x) {
    const long long t = t0 + threadIdx.x;
    const long long c = (t < n_tiles) ? cnt[t] : 0;
    long long x = c;
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      long long w = s_warp[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const long long carry = s_carry;
    const long long excl = carry + (warp ? s_warp[warp - 1] : 0) + x - c;
    if (t < n_tiles) base[t] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = carry + s_warp[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) gen_total[g] = s_carry;
}

// exclusive scan over generations (tiny): gen_offsets[g] = first row of generation g, [G] = total
__global__ void gen_offsets_kernel(long long* gen_offsets, int generation_limit) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long acc = 0;
    for (int g = 0; g < generation_limit; ++g) {
      const long long c = gen_offsets[g];
      gen_offsets[g] = acc;
      acc += c;
    }
    gen_offsets[generation_limit] = acc;
  }
}

// one block per tile: expand each of the tile's runs of staged records into its final frame rows
// (_RayTraceDataframe.insert, pyrayt/_pyrayt.py:168-186: current metadata, surface id, start point, end
// point = start + direction * distance (:404-407), unit tilt (:177)).  Same functions and the same
// -fmad=false arithmetic as the trace kernel, so the columns are bit-identical to computing them there.
// Four blocks per SM (64 registers; 28 bytes of spills): at the 72 registers the compiler takes unasked the
// pass loses a quarter of its warps and 0.9 ms; five blocks (48 registers) and plain IEEE divisions in place of
// unit_tilt were measured too (config 4 step 45.9 / 45.2 ms against 44.9; profiles/r2/kbench_r2o.txt).
#ifndef PRT_GATHER_BLOCKS
#define PRT_GATHER_BLOCKS 4
#endif
template <int LAYOUT>
__global__ void __launch_bounds__(kTileRays, PRT_GATHER_BLOCKS) gather_kernel(const GatherArgs a) {
  const long long tile = blockIdx.x;
  const Leaf* leaves =
      reinterpret_cast<const Leaf*>(a.blob + reinterpret_cast<const BlobHeader*>(a.blob)->off_leaves);
  for (int g = 0; g < a.generation_limit; ++g) {
    const long long idx = (long long)g * a.n_tiles + tile;
    const int c = a.run_count[idx];
    if (c == 0) continue;
    if ((int)threadIdx.x < c) {
      const long long src = a.run_start[idx] + threadIdx.x;
      const long long dst = a.gen_offsets[g] + a.run_base[idx] + threadIdx.x;
      if (dst >= a.frame_capacity) continue;
      // all loads first (memory-level parallelism: a version that stored each column as soon as it was known
      // was 1.5 ms slower), then the fifteen stores
      double s[kStageCols];
#pragma unroll
      for (int k = 0; k < kStageCols; ++k) s[k] = __ldcs(a.stage + k * a.capacity + src);
      const long long meta = __double_as_longlong(s[8]);
      const long long ray = tile * kTileRays + (meta >> 32);
      const Leaf& L = leaves[(int)(meta & 0xffffffffll)];
      double v[kFrameCols];
      v[0] = (g == 0) ? a.rays[8 * a.ray_stride + ray] : (double)g;  // :440-441
      v[1] = a.rays[9 * a.ray_stride + ray];
      v[2] = a.rays[10 * a.ray_stride + ray];
      v[3] = s[7];
      v[4] = a.rays[12 * a.ray_stride + ray];
      v[5] = L.sid;
      v[6] = s[0];
      v[7] = s[1];
      v[8] = s[2];
      v[9] = s[0] + s[3] * s[6];
      v[10] = s[1] + s[4] * s[6];
      v[11] = s[2] + s[5] * s[6];
      const double vn = sqrt(s[3] * s[3] + s[4] * s[4] + s[5] * s[5]);  // as step_speed
      unit_tilt(s[3], s[4], s[5], vn, v[12], v[13], v[14]);
      if (LAYOUT == 0) {
#pragma unroll
        for (int k = 0; k < kFrameCols; ++k) __stcs(a.frame + k * a.frame_stride + dst, v[k]);
      } else {
#pragma unroll
        for (int k = 0; k < kFrameCols; ++k) a.frame[dst * kFrameCols + k] = v[k];
      }
    }
  }
}

// ---------------------------------------------------------------- component.intersect

template <bool GLOBAL>
__global__ void __launch_bounds__(kTileRays) intersect_kernel(const unsigned char* blob, int blob_bytes,
                                                              int component, const double* rays, long long n,
                                                              double* hits, long long* sids, int slots) {
  extern __shared__ __align__(16) unsigned char s_blob[];
  if (!GLOBAL) {  // (GLOBAL: a scene too large for shared memory is read in place, through L1 / L2)
    const int words = blob_bytes / 8;
    const double* src = reinterpret_cast<const double*>(blob);
    double* dst = reinterpret_cast<double*>(s_blob);
    for (int w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
  }
  __syncthreads();
  const SceneView sc = make_view(GLOBAL ? blob : s_blob);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double p0 = rays[0 * n + i], p1 = rays[1 * n + i], p2 = rays[2 * n + i];
  const double v0 = rays[4 * n + i], v1 = rays[5 * n + i], v2 = rays[6 * n + i];
  // the literal fixed-length lists: +inf slots and the surface ids they carry are part of what
  // component.intersect returns (prt_literal.cuh)
  LitList stack[kLitStack];
  eval_component_literal(sc, sc.comps[component].begin, sc.comps[component].end, p0, p1, p2, v0, v1, v2,
                         make_ray_inv(p0, p1, p2, v0, v1, v2, (sc.h->flags & 1) != 0), stack);
  const LitList& r = stack[0];
  for (int k = 0; k < slots; ++k) {
    hits[k * n + i] = (k < r.n) ? r.t[k] : PRT_INF;
    sids[k * n + i] = (k < r.n && r.leaf[k] >= 0) ? (long long)sc.leaves[r.leaf[k]].sid : -1;
  }
}

// ---------------------------------------------------------------- _st_propagate alone (renderers, probes)

template <bool GLOBAL>
__global__ void __launch_bounds__(kTileRays, PRT_MIN_BLOCKS) nearest_kernel(const unsigned char* blob, int blob_bytes,
                                                                           const double* rays, long long n, double* t_out,
                                                                           long long* sid_out, double* normals) {
  extern __shared__ __align__(16) unsigned char s_blob[];
  if (!GLOBAL) {  // (GLOBAL: a scene too large for shared memory is read in place, through L1 / L2)
    const int words = blob_bytes / 8;
    const double* src = reinterpret_cast<const double*>(blob);
    double* dst = reinterpret_cast<double*>(s_blob);
    for (int w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
  }
  __syncthreads();
  const SceneView sc = make_view(GLOBAL ? blob : s_blob);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double p0 = rays[0 * n + i], p1 = rays[1 * n + i], p2 = rays[2 * n + i];
  const double v0 = rays[4 * n + i], v1 = rays[5 * n + i], v2 = rays[6 * n + i];
  HitStack S;
  double best_t;
  int best_leaf;
  bool tie = false;
  nearest_hit<true>(sc, p0, p1, p2, v0, v1, v2, -1, &S, best_t, best_leaf, tie);
  t_out[i] = best_t;
  sid_out[i] = best_leaf >= 0 ? (long long)sc.leaves[best_leaf].sid : -1;
  if (normals) {
    double n0 = CUDART_NAN, n1 = CUDART_NAN, n2 = CUDART_NAN;
    if (best_leaf >= 0)
      world_normal(sc.leaves[best_leaf], p0 + v0 * best_t, p1 + v1 * best_t, p2 + v2 * best_t, n0, n1, n2);
    normals[0 * n + i] = n0;
    normals[1 * n + i] = n1;
    normals[2 * n + i] = n2;
  }
}

// EdgeRender / ShadedRenderer._st_propagate (tinygfx/g3d/renderers.py:72-94,:188-210): the tracer's loop,
// except that the distance and surface are read from the *unfiltered* hit array at the argmin of the
// filtered one -- a pixel whose component hits are all behind the camera reports slot 0, a negative
// distance.  Needs every hit of every component, so it runs the interpreter without pruning.
template <bool GLOBAL>
__global__ void __launch_bounds__(kTileRays) render_hit_kernel(const unsigned char* blob, int blob_bytes,
                                                               const double* rays, long long n, double* t_out,
                                                               long long* sid_out, double* normals) {
  extern __shared__ __align__(16) unsigned char s_blob[];
  if (!GLOBAL) {  // (GLOBAL: a scene too large for shared memory is read in place, through L1 / L2)
    const int words = blob_bytes / 8;
    const double* src = reinterpret_cast<const double*>(blob);
    double* dst = reinterpret_cast<double*>(s_blob);
    for (int w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
  }
  __syncthreads();
  const SceneView sc = make_view(GLOBAL ? blob : s_blob);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double p0 = rays[0 * n + i], p1 = rays[1 * n + i], p2 = rays[2 * n + i];
  const double v0 = rays[4 * n + i], v1 = rays[5 * n + i], v2 = rays[6 * n + i];
  const RayInv inv = make_ray_inv(p0, p1, p2, v0, v1, v2, (sc.h->flags & 1) != 0);
  double best_t = PRT_INF;
  int best_leaf = -1;
  for (int c = 0; c < sc.h->n_components; ++c) {
    HitStack S;
    S.flags = 0;
    bool tie = false;
    const bool any = eval_component(sc, sc.comps[c].begin, sc.comps[c].end, p0, p1, p2, v0, v1, v2, inv, false,
                                    PRT_INF, S, tie);
    const int len = any ? S.len[0] : 0;
    if (len == 0) continue;  // culled or empty: every slot is +inf / -1
    const int b = buf_of(S, 0);
    int arg = 0;  // ascending list: the first positive entry is the argmin of where(hits > 0, hits, inf)
    while (arg < len && !(S.t[b][arg] > 0.0)) ++arg;
    if (arg == len) arg = 0;
    const double t = S.t[b][arg];
    if (t < best_t) {
      best_t = t;
      best_leaf = S.leaf[b][arg];
    }
  }
  t_out[i] = best_t;
  sid_out[i] = best_leaf >= 0 ? (long long)sc.leaves[best_leaf].sid : -1;
  if (normals) {
    double n0 = CUDART_NAN, n1 = CUDART_NAN, n2 = CUDART_NAN;
    if (best_leaf >= 0)
      world_normal(sc.leaves[best_leaf], p0 + v0 * best_t, p1 + v1 * best_t, p2 + v2 * best_t, n0, n1, n2);
    normals[0 * n + i] = n0;
    normals[1 * n + i] = n1;
    normals[2 * n + i] = n2;
  }
}

// ---------------------------------------------------------------- K3: seeded synthetic sources
//
// Counter-based uniforms u(i,k) = mix64(seed ^ (i*C1 + k*C2)) >> 11 * 2^-53; only + - * / sqrt
// follow, so the NumPy restatement (oracle/sources_np.py) is bit-identical.

__device__ __forceinline__ double u01(unsigned long long seed, unsigned long long i, unsigned long long k) {
  unsigned long long z = seed ^ (i * 0x9E3779B97F4A7C15ull + k * 0xD1B54A32D192ED03ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// rejection-sample a point of the unit disk (radius^2 in (1e-12, 1])
__device__ __forceinline__ void unit_disk(unsigned long long seed, unsigned long long i, unsigned long long k0,
                                          double& a, double& b) {
  a = 0;
  b = 0;
  for (unsigned long long k = 0; k < 64; k += 2) {
    const double x = 2 * u01(seed, i, k0 + k) - 1;
    const double y = 2 * u01(seed, i, k0 + k + 1) - 1;
    const double r2 = x * x + y * y;
    if (r2 <= 1.0 && r2 > 1e-12) {
      a = x;
      b = y;
      return;
    }
  }
}

__global__ void source_kernel(const prt_source_desc src, double* rays, long long n, long long stride,
                              long long first) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const unsigned long long i = (unsigned long long)(first + j);
  double px = src.origin[0], py = src.origin[1], pz = src.origin[2];
  double dx = 1, dy = 0, dz = 0, wl = 0.633, inten = 100.0;
  if (src.kind == 1) {
    // collimated field fan: disk of radius p[0] in the plane x = origin.x; field i%3 rotates +x about z
    // by (cos,sin) = p[2+2f], p[3+2f]; wavelength p[8 + (i/3)%3]; intensity p[11]
    double a, b;
    unit_disk(src.seed, i, 0, a, b);
    py = src.origin[1] + src.p[0] * a;
    pz = src.origin[2] + src.p[0] * b;
    const int f = (int)(i % 3ull);
    dx = src.p[2 + 2 * f];
    dy = src.p[3 + 2 * f];
    dz = 0;
    wl = src.p[8 + (int)((i / 3ull) % 3ull)];
    inten = src.p[11];
  } else if (src.kind == 2) {
    // point source, directions uniform in solid angle inside a cone about +x: p[0] = cos(theta_max)
    const double ct = 1 - u01(src.seed, i, 0) * (1 - src.p[0]);
    const double st = sqrt(1 - ct * ct);
    double a, b;
    unit_disk(src.seed, i, 1, a, b);
    const double r = sqrt(a * a + b * b);
    dx = ct;
    dy = st * (a / r);
    dz = st * (b / r);
    wl = src.p[1];
    inten = src.p[2];
  } else if (src.kind == 3) {
    // point source, Lambertian within a cone about -x: p[0] = sin(theta_max) (Malley's method)
    double a, b;
    unit_disk(src.seed, i, 0, a, b);
    const double u = src.p[0] * a, v = src.p[0] * b;
    dx = -sqrt(1 - (u * u + v * v));
    dy = u;
    dz = v;
    wl = src.p[1];
    inten = src.p[2];
  }
  double* r = rays + j;
  r[0 * stride] = px;
  r[1 * stride] = py;
  r[2 * stride] = pz;
  r[3 * stride] = 1.0;
  r[4 * stride] = dx;
  r[5 * stride] = dy;
  r[6 * stride] = dz;
  r[7 * stride] = 0.0;
  r[8 * stride] = 0.0;
  r[9 * stride] = inten;
  r[10 * stride] = wl;
  r[11 * stride] = 1.0;
  r[12 * stride] = (double)i;
}

// The reference's deterministic Source classes generated on the device:
// _local_ray_generation of LineOfRays / CircleOfRays / ConeOfRays / WedgeOfRays
// (pyrayt/components.py:511-613), then Source.generate_rays' world transform and direction
// normalisation (:481-496).  p[0] = spacing | diameter | cone angle [rad] | wedge angle [rad],
// p[1] = wavelength, p[2] = rays of this source, p[3] = id of its first ray, p[4..15] = rows 0..2
// of the source's world matrix.  linspace / arange arithmetic follows NumPy's formulas exactly;
// sin/cos are the CUDA double-precision functions (within 1-2 ulp of NumPy's).  A launch writes the
// window [first, first + count) of the source's n rays (a rank's share of a sharded source).
__global__ void reference_source_kernel(const prt_source_desc src, double* rays, long long stride, long long first,
                                        long long count) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)src.p[2];
  if (w >= count) return;
  const long long j = first + w;
  const double jd = (double)j, nd = (double)n;
  double lx = 0, ly = 0, lz = 0, ux = 0, uy = 0, uz = 0, inten = 100.0;
  const double two_pi = 2 * 3.141592653589793;
  if (src.kind == 10) {  // LineOfRays: y = linspace(-s/2, s/2, n), direction +x
    if (n > 1) {
      const double start = -src.p[0] / 2, stop = src.p[0] / 2;
      const double step = (stop - start) / (nd - 1);
      ly = (j == n - 1) ? stop : (step != 0 ? jd * step + start : jd / (nd - 1) * (stop - start) + start);
    }
    ux = 1;
  } else if (src.kind == 11) {  // CircleOfRays: theta = linspace(0, 2 pi, n)
    double th = 0;
    if (n > 1) {
      const double step = two_pi / (nd - 1);
      th = (j == n - 1) ? two_pi : jd * step;
    }
    ly = src.p[0] / 2 * sin(th);
    lz = src.p[0] / 2 * cos(th);
    ux = 1;
  } else if (src.kind == 12) {  // ConeOfRays: angles = 2 pi arange(n) / n
    if (n > 1) {
      const double ang = two_pi * jd / nd;
      uy = sin(src.p[0]) * sin(ang);
      uz = sin(src.p[0]) * cos(ang);
    }
    ux = cos(src.p[0]);
  } else if (src.kind == 14) {
    // Lamp / StaticLamp (pyrayt/components.py:616-662, _sphere_sample :56-70): same law -- theta =
    // arccos(1 - u (1 - cos max_angle)), phi = 2 pi u', start point uniform on the width x length rectangle,
    // intensity 100 cos(theta) -- with the counter-based uniforms u01(seed, ray id, k) in place of NumPy's
    // global Mersenne-Twister stream (opt-in: RayTracer.lamp_seed; origin[0] = width, origin[1] = length)
    const unsigned long long rid = (unsigned long long)(src.p[3] + jd);
    const double theta = acos(1 - u01(src.seed, rid, 0) * (1 - cos(src.p[0])));
    const double phi = u01(src.seed, rid, 1) * two_pi;
    ly = src.origin[0] * (u01(src.seed, rid, 2) - 0.5);
    lz = src.origin[1] * (u01(src.seed, rid, 3) - 0.5);
    ux = cos(theta);
    uy = sin(theta) * cos(phi);
    uz = sin(theta) * sin(phi);
    inten = 100.0 * cos(theta);
  } else {  // 13 WedgeOfRays: angles = linspace(-a/2, a/2, n)
    double ang;
    if (n > 1) {
      const double start = -src.p[0] / 2, stop = src.p[0] / 2;
      const double step = (stop - start) / (nd - 1);
      ang = (j == n - 1) ? stop : (step != 0 ? jd * step + start : jd / (nd - 1) * (stop - start) + start);
    } else {
      ang = -src.p[0] / 2;  // linspace(start, stop, 1) = [start]
    }
    ux = cos(ang);
    uy = sin(ang);
  }
  const double* M = src.p + 4;
  const double px = M[0] * lx + M[1] * ly + M[2] * lz + M[3];
  const double py = M[4] * lx + M[5] * ly + M[6] * lz + M[7];
  const double pz = M[8] * lx + M[9] * ly + M[10] * lz + M[11];
  double dx = M[0] * ux + M[1] * uy + M[2] * uz;
  double dy = M[4] * ux + M[5] * uy + M[6] * uz;
  double dz = M[8] * ux + M[9] * uy + M[10] * uz;
  const double nrm = sqrt(dx * dx + dy * dy + dz * dz);
  dx /= nrm;
  dy /= nrm;
  dz /= nrm;
  double* r = rays + w;
  r[0 * stride] = px;
  r[1 * stride] = py;
  r[2 * stride] = pz;
  r[3 * stride] = 1.0;
  r[4 * stride] = dx;
  r[5 * stride] = dy;
  r[6 * stride] = dz;
  r[7 * stride] = 0.0;
  r[8 * stride] = 0.0;
  r[9 * stride] = inten;
  r[10 * stride] = src.p[1];
  r[11 * stride] = 1.0;
  r[12 * stride] = src.p[3] + jd;
}

// ---------------------------------------------------------------- FP64 pipe probe (roofline denominator)
//
// 8 independent DFMA chains per thread; bench.py times it with CUDA events to get the
// measured FP64 FMA rate of this GPU (flops = threads * iters * 8 * 2).
__global__ void __launch_bounds__(256) fp64_probe_kernel(double* out, int iters, double seed) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6,
         a7 = seed + 7;
  const double m = 1.0000001, c = 1e-9;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c);
    a1 = fma(a1, m, c);
    a2 = fma(a2, m, c);
    a3 = fma(a3, m, c);
    a4 = fma(a4, m, c);
    a5 = fma(a5, m, c);
    a6 = fma(a6, m, c);
    a7 = fma(a7, m, c);
  }
  const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (r == 0.123456789) out[0] = r;  // keeps the chains alive, never true in practice
}

}  // namespace prt

// ---------------------------------------------------------------- launchers used by prt_abi.cpp

template <bool RECORD, bool GENERIC, bool DIAG = false>
static cudaError_t launch_trace_variant(const prt::TraceArgs* a, unsigned tiles, size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(prt::trace_kernel<RECORD, GENERIC, DIAG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem);
  prt::trace_kernel<RECORD, GENERIC, DIAG><<<tiles, prt::kTileRays, smem, st>>>(*a);
  return cudaGetLastError();
}

extern "C" {

// generic != 0: some component needs the interpreter for arbitrary CSG trees; diagnose != 0: PRT_FLAG_DIAGNOSE
// (one variant, the one with the interpreter, serves every scene)
cudaError_t prt_launch_trace(const prt::TraceArgs* a, int record, int generic, int diagnose, cudaStream_t st) {
  const long long tiles = (a->n_rays + prt::kTileRays - 1) / prt::kTileRays;
  if (tiles == 0) return cudaSuccess;
  if (a->blob_bytes > prt::kMaxSharedBlob) {  // large scene: read from global memory, interpreter variant
    if (diagnose) {
      if (record) prt::trace_kernel<true, true, true, true><<<(unsigned)tiles, prt::kTileRays, 0, st>>>(*a);
      else prt::trace_kernel<false, true, true, true><<<(unsigned)tiles, prt::kTileRays, 0, st>>>(*a);
    } else {
      if (record) prt::trace_kernel<true, true, false, true><<<(unsigned)tiles, prt::kTileRays, 0, st>>>(*a);
      else prt::trace_kernel<false, true, false, true><<<(unsigned)tiles, prt::kTileRays, 0, st>>>(*a);
    }
    return cudaGetLastError();
  }
  const size_t smem = (size_t)a->blob_bytes;
  if (diagnose)
    return record ? launch_trace_variant<true, true, true>(a, (unsigned)tiles, smem, st)
                  : launch_trace_variant<false, true, true>(a, (unsigned)tiles, smem, st);
  if (record) {
    return generic ? launch_trace_variant<true, true>(a, (unsigned)tiles, smem, st)
                   : launch_trace_variant<true, false>(a, (unsigned)tiles, smem, st);
  }
  return generic ? launch_trace_variant<false, true>(a, (unsigned)tiles, smem, st)
                 : launch_trace_variant<false, false>(a, (unsigned)tiles, smem, st);
}

cudaError_t prt_launch_scan(const int* run_count, long long* run_base, long long n_tiles, int generation_limit,
                            long long* gen_offsets, cudaStream_t st) {
  if (generation_limit <= 0) return cudaSuccess;
  prt::scan_runs_kernel<<<generation_limit, 1024, 0, st>>>(run_count, run_base, n_tiles, gen_offsets);
  prt::gen_offsets_kernel<<<1, 32, 0, st>>>(gen_offsets, generation_limit);
  return cudaGetLastError();
}

cudaError_t prt_launch_gather(const prt::GatherArgs* a, int layout, cudaStream_t st) {
  if (a->n_tiles == 0) return cudaSuccess;
  if (layout == 0)
    prt::gather_kernel<0><<<(unsigned)a->n_tiles, prt::kTileRays, 0, st>>>(*a);
  else
    prt::gather_kernel<1><<<(unsigned)a->n_tiles, prt::kTileRays, 0, st>>>(*a);
  return cudaGetLastError();
}

cudaError_t prt_launch_intersect(const unsigned char* blob, int blob_bytes, int component, const double* rays,
                                 long long n, double* hits, long long* sids, int slots, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((n + prt::kTileRays - 1) / prt::kTileRays);
  if (blob_bytes > prt::kMaxSharedBlob) {
    prt::intersect_kernel<true><<<blocks, prt::kTileRays, 0, st>>>(blob, blob_bytes, component, rays, n, hits, sids,
                                                                   slots);
    return cudaGetLastError();
  }
  if (blob_bytes > 48 * 1024)
    cudaFuncSetAttribute(prt::intersect_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, blob_bytes);
  prt::intersect_kernel<false><<<blocks, prt::kTileRays, (size_t)blob_bytes, st>>>(blob, blob_bytes, component, rays,
                                                                                   n, hits, sids, slots);
  return cudaGetLastError();
}

cudaError_t prt_launch_nearest(const unsigned char* blob, int blob_bytes, const double* rays, long long n, double* t_out,
                               long long* sid_out, double* normals, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((n + prt::kTileRays - 1) / prt::kTileRays);
  if (blob_bytes > prt::kMaxSharedBlob) {
    prt::nearest_kernel<true><<<blocks, prt::kTileRays, 0, st>>>(blob, blob_bytes, rays, n, t_out, sid_out, normals);
    return cudaGetLastError();
  }
  if (blob_bytes > 48 * 1024)
    cudaFuncSetAttribute(prt::nearest_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, blob_bytes);
  prt::nearest_kernel<false><<<blocks, prt::kTileRays, (size_t)blob_bytes, st>>>(blob, blob_bytes, rays, n, t_out,
                                                                                 sid_out, normals);
  return cudaGetLastError();
}

cudaError_t prt_launch_render_hit(const unsigned char* blob, int blob_bytes, const double* rays, long long n,
                                  double* t_out, long long* sid_out, double* normals, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((n + prt::kTileRays - 1) / prt::kTileRays);
  if (blob_bytes > prt::kMaxSharedBlob) {
    prt::render_hit_kernel<true><<<blocks, prt::kTileRays, 0, st>>>(blob, blob_bytes, rays, n, t_out, sid_out,
                                                                    normals);
    return cudaGetLastError();
  }
  if (blob_bytes > 48 * 1024)
    cudaFuncSetAttribute(prt::render_hit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, blob_bytes);
  prt::render_hit_kernel<false><<<blocks, prt::kTileRays, (size_t)blob_bytes, st>>>(blob, blob_bytes, rays, n, t_out,
                                                                                    sid_out, normals);
  return cudaGetLastError();
}

cudaError_t prt_launch_source(const prt_source_desc* src, double* rays, long long n, long long stride,
                              long long first, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (src->kind >= 10)
    prt::reference_source_kernel<<<blocks, 256, 0, st>>>(*src, rays, stride, first, n);
  else
    prt::source_kernel<<<blocks, 256, 0, st>>>(*src, rays, n, stride, first);
  return cudaGetLastError();
}

cudaError_t prt_launch_fp64_probe(double* out, int blocks, int iters, cudaStream_t st) {
  prt::fp64_probe_kernel<<<blocks, 256, 0, st>>>(out, iters, 1.0);
  return cudaGetLastError();
}

}  // extern "C"
