// The optional FP32 fast mode of the trace (PRT_FLAG_FP32): trace_kernel_f32 + gather_kernel_f32.
//
// Same structure as prt_kernels.cu -- one thread = one ray, every generation in a persistent loop, scene in
// shared memory, block-aggregated append of staged records, ordering pass -- with the per-ray arithmetic of
// prt_device_f32.cuh.  This file is compiled WITH FMA contraction (the Makefile drops -fmad=false for it):
// the mode's contract is a tolerance, not the reference's roundings.
// A staged record is 40 bytes (five 8-byte columns of the same staging buffer): start position, direction,
// hit distance and refractive index as floats, packed two to a word, plus the (tile slot, leaf) word.
// The frame keeps the reference's fifteen float64 columns.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pyrayt_b200.h"
#include "prt_scene.h"
#include "prt_device_f32.cuh"

namespace prt {
namespace f32 {

constexpr int kStageColsF = 5;

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double pack2(float a, float b) {
  return __longlong_as_double(((long long)__float_as_uint(b) << 32) | (long long)__float_as_uint(a));
}
__device__ __forceinline__ void unpack2(double w, float& a, float& b) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(w);
  a = __uint_as_float((unsigned)(u & 0xffffffffull));
  b = __uint_as_float((unsigned)(u >> 32));
}

// bytes of dynamic shared memory: the FP64 blob, then LeafF[n_leaves], CompF[n_components],
// OrderEntryF[6][n_components] and float[6][n_aabb]
__host__ __device__ inline int blob_aligned(int blob_bytes) { return (blob_bytes + 15) & ~15; }

template <bool GENERIC>
struct StackForF {
  typedef HitStackF type;
  __device__ static HitStackF* ptr(HitStackF& s) { return &s; }
};
template <>
struct StackForF<false> {
  typedef char type;
  __device__ static HitStackF* ptr(char&) { return nullptr; }
};

#ifndef PRT_F32_MIN_BLOCKS
#define PRT_F32_MIN_BLOCKS 4
#endif

template <bool RECORD, bool ORDERED, bool GENERIC>
__global__ void __launch_bounds__(kTileRays, PRT_F32_MIN_BLOCKS) trace_kernel_f32(const TraceArgs a, int n_leaves, int n_components) {
  extern __shared__ __align__(16) unsigned char s_mem[];
  __shared__ int s_wcount[kTileRays / 32];
  __shared__ long long s_base;
  __shared__ unsigned s_ctr0[kTileRays], s_ctr1[kTileRays];

  // stage the scene: the blob as it is, then single-precision copies of what the per-ray code reads
  {
    const int words = a.blob_bytes / 8;
    const double* src = reinterpret_cast<const double*>(a.blob);
    double* dst = reinterpret_cast<double*>(s_mem);
    for (int w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
  }
  __syncthreads();
  LeafF* lf = reinterpret_cast<LeafF*>(s_mem + blob_aligned(a.blob_bytes));
  CompF* cf = reinterpret_cast<CompF*>(lf + n_leaves);
  OrderEntryF* ordf = reinterpret_cast<OrderEntryF*>(cf + n_components);
  float* aabbf = reinterpret_cast<float*>(ordf + 6 * n_components);  // GENERIC: node boxes of the interpreter
  {
    const BlobHeader* h = reinterpret_cast<const BlobHeader*>(s_mem);
    const OrderEntry* ord = reinterpret_cast<const OrderEntry*>(s_mem + h->off_order);
    for (int e = threadIdx.x; e < 6 * h->n_boxed; e += blockDim.x) convert_order(ord[e], ordf[e]);
    if (GENERIC) {
      const double* box = reinterpret_cast<const double*>(s_mem + h->off_aabb);
      for (int e = threadIdx.x; e < 6 * h->n_aabb; e += blockDim.x)
        aabbf[e] = (e % 6 & 1) ? to_float_up(box[e]) : to_float_dn(box[e]);
    }
    const Leaf* leaves = reinterpret_cast<const Leaf*>(s_mem + h->off_leaves);
    const Comp* comps = reinterpret_cast<const Comp*>(s_mem + h->off_comps);
    for (int l = threadIdx.x; l < n_leaves; l += blockDim.x) convert_leaf(leaves[l], lf[l]);
    for (int c = threadIdx.x; c < n_components; c += blockDim.x) convert_comp(comps[c], cf[c]);
  }
  __syncthreads();
  SceneViewF sc;
  sc.h = reinterpret_cast<const BlobHeader*>(s_mem);
  sc.comps = reinterpret_cast<const Comp*>(s_mem + sc.h->off_comps);
  sc.leaves = reinterpret_cast<const Leaf*>(s_mem + sc.h->off_leaves);
  sc.lf = lf;
  sc.cf = cf;
  // ray-ordered traversal for scenes with many boxed components (the encoder's choice, flags bit 3, as on the
  // FP64 path: config 4 K1 19.5 -> 14.3 ms; small scenes are faster in list order and run the ORDERED = false
  // variant, which carries none of the table walk)
  sc.order = ORDERED ? ordf : nullptr;
  sc.unboxed = reinterpret_cast<const int*>(s_mem + sc.h->off_unboxed);
  sc.ops = reinterpret_cast<const Op*>(s_mem + sc.h->off_ops);
  sc.aabb = GENERIC ? aabbf : nullptr;
  // hit lists of the interpreter (local memory): only the GENERIC variants carry them
  typename StackForF<GENERIC>::type stack_storage;
  HitStackF* S = StackForF<GENERIC>::ptr(stack_storage);

  const long long tile = blockIdx.x;
  const long long i = tile * kTileRays + threadIdx.x;
  const bool valid = i < a.n_rays;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;

  RayStateF rs = {0, 0, 0, 0, 0, 0, 0, 1, -1, -1};
  unsigned c_drop = 0, c_badw = 0;
  s_ctr0[threadIdx.x] = 0;
  s_ctr1[threadIdx.x] = 0;
  if (valid) {
    const double* r = a.rays + i;
    rs.p0 = (float)r[0 * a.stride];
    rs.p1 = (float)r[1 * a.stride];
    rs.p2 = (float)r[2 * a.stride];
    const double pw = r[3 * a.stride];
    rs.v0 = (float)r[4 * a.stride];
    rs.v1 = (float)r[5 * a.stride];
    rs.v2 = (float)r[6 * a.stride];
    const double vw = r[7 * a.stride];
    rs.wl = (float)r[10 * a.stride];
    rs.nidx = (float)r[11 * a.stride];
    if (pw != 1.0 || vw != 0.0) c_badw = 1;
  }
  bool alive = valid;
  const float ray_offset = (float)a.ray_offset;

  for (int g = 0; g < a.generation_limit; ++g) {
    float vn = 0.0f, hit_t = 0.0f;
    int hit_leaf = -1;
    StepCounters ctr = {s_ctr0[threadIdx.x], s_ctr1[threadIdx.x]};
    if (alive) {
      vn = step_speed(rs, ctr);
      if (vn != 0.0f) {
        bool tie = false;
        nearest_hit<ORDERED, GENERIC>(sc, rs, ray_scale(rs), S, hit_t, hit_leaf, tie);
        if (tie) ctr.w1 |= kCtrTie;
      }
    }
    const int mat = (hit_leaf >= 0) ? sc.lf[hit_leaf].mat : PRT_MAT_UNTRACEABLE;
    const bool has_row = (hit_leaf >= 0) & (mat != PRT_MAT_UNTRACEABLE);
    const bool next_alive = has_row & (mat != PRT_MAT_ABSORBER) & (g + 1 != a.generation_limit);

    int any_alive = 1;
    if (RECORD) {
      const bool write =
          has_row && (a.record_mode == PRT_RECORD_ALL || sc.leaves[hit_leaf].sid == a.detector_sid);
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
          double* o = a.stage + row;
          const long long cs = a.capacity;
          o[0 * cs] = pack2(rs.p0, rs.p1);
          o[1 * cs] = pack2(rs.p2, rs.v0);
          o[2 * cs] = pack2(rs.v1, rs.v2);
          o[3 * cs] = pack2(hit_t, rs.nidx);
          o[4 * cs] = __longlong_as_double(((long long)threadIdx.x << 32) | (long long)hit_leaf);
        } else {
          c_drop++;
        }
      }
    }

    alive = false;
    if (hit_leaf >= 0) {
      StepOutF so;
      alive = step_interact(sc, rs, g, a.generation_limit, vn, hit_t, hit_leaf, so, ctr);
      if (alive) advance_ray(rs, so, hit_leaf, ray_offset);
    }
    s_ctr0[threadIdx.x] = ctr.w0;
    s_ctr1[threadIdx.x] = ctr.w1;
    if (RECORD ? !any_alive : !alive) break;
  }

  const StepCounters sc_ctr = {s_ctr0[threadIdx.x], s_ctr1[threadIdx.x]};
  unsigned long long vals[11] = {valid ? 1ull : 0ull,
                                 sc_ctr.w0 & 0xffffu,
                                 sc_ctr.w0 >> 16,
                                 c_drop,
                                 (sc_ctr.w1 & kCtrTie) ? 1ull : 0ull,
                                 (sc_ctr.w1 & kCtrUntr) ? 1ull : 0ull,
                                 c_badw,
                                 (sc_ctr.w1 & kCtrNan) ? 1ull : 0ull,
                                 (sc_ctr.w1 & kCtrLim) ? 1ull : 0ull,
                                 (sc_ctr.w1 & kCtrAbs) ? 1ull : 0ull,
                                 sc_ctr.w1 & 0xffffu};
  unsigned long long* dst[11] = {
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
      reinterpret_cast<unsigned long long*>(&a.ctr->mirror_segments)};
#pragma unroll
  for (int q = 0; q < 11; ++q) {
    const unsigned long long s = warp_sum(vals[q]);
    if (lane == 0 && s) atomicAdd(dst[q], s);
  }
}

// ordering pass of the fast mode: 40-byte staged records -> the fifteen float64 frame columns in
// (generation, id) order (_RayTraceDataframe.insert, pyrayt/_pyrayt.py:168-186)
__global__ void __launch_bounds__(kTileRays) gather_kernel_f32(const GatherArgs a) {
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
      double s[kStageColsF];
#pragma unroll
      for (int k = 0; k < kStageColsF; ++k) s[k] = __ldcs(a.stage + k * a.capacity + src);
      float p0, p1, p2, v0, v1, v2, t, nidx;
      unpack2(s[0], p0, p1);
      unpack2(s[1], p2, v0);
      unpack2(s[2], v1, v2);
      unpack2(s[3], t, nidx);
      const long long meta = __double_as_longlong(s[4]);
      const long long ray = tile * kTileRays + (meta >> 32);
      const Leaf& L = leaves[(int)(meta & 0xffffffffll)];
      const float rv = frcp(fsqrt(v0 * v0 + v1 * v1 + v2 * v2));
      double v[kFrameCols];
      v[0] = (g == 0) ? a.rays[8 * a.ray_stride + ray] : (double)g;
      v[1] = a.rays[9 * a.ray_stride + ray];
      v[2] = a.rays[10 * a.ray_stride + ray];
      v[3] = (double)nidx;
      v[4] = a.rays[12 * a.ray_stride + ray];
      v[5] = L.sid;
      v[6] = (double)p0;
      v[7] = (double)p1;
      v[8] = (double)p2;
      v[9] = (double)(p0 + v0 * t);  // the same expression (and contraction) as step_interact's hit point
      v[10] = (double)(p1 + v1 * t);
      v[11] = (double)(p2 + v2 * t);
      v[12] = (double)(v0 * rv);
      v[13] = (double)(v1 * rv);
      v[14] = (double)(v2 * rv);
#pragma unroll
      for (int k = 0; k < kFrameCols; ++k) __stcs(a.frame + k * a.frame_stride + dst, v[k]);
    }
  }
}

}  // namespace f32
}  // namespace prt

extern "C" {

// dynamic shared memory of trace_kernel_f32 for a scene
size_t prt_f32_smem_bytes(int blob_bytes, int n_leaves, int n_components, int n_aabb) {
  return (size_t)prt::f32::blob_aligned(blob_bytes) + (size_t)n_leaves * sizeof(prt::f32::LeafF) +
         (size_t)n_components * sizeof(prt::f32::CompF) + (size_t)6 * n_components * sizeof(prt::f32::OrderEntryF) +
         (size_t)6 * n_aabb * sizeof(float);
}

// ordered != 0: the scene's header asks for the ray-ordered walk (BlobHeader.flags bits 2 and 3, n_boxed > 0);
// generic != 0: some component needs the interpreter for arbitrary CSG trees (n_aabb = its node boxes)
cudaError_t prt_launch_trace_f32(const prt::TraceArgs* a, int record, int ordered, int generic, int n_leaves,
                                 int n_components, int n_aabb, cudaStream_t st) {
  const long long tiles = (a->n_rays + prt::kTileRays - 1) / prt::kTileRays;
  if (tiles == 0) return cudaSuccess;
  const size_t smem = prt_f32_smem_bytes(a->blob_bytes, n_leaves, n_components, generic ? n_aabb : 0);
  auto launch = [&](auto kernel) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kernel<<<(unsigned)tiles, prt::kTileRays, smem, st>>>(*a, n_leaves, n_components);
  };
  const int variant = (record ? 4 : 0) | (ordered ? 2 : 0) | (generic ? 1 : 0);
  switch (variant) {
    case 0: launch(prt::f32::trace_kernel_f32<false, false, false>); break;
    case 1: launch(prt::f32::trace_kernel_f32<false, false, true>); break;
    case 2: launch(prt::f32::trace_kernel_f32<false, true, false>); break;
    case 3: launch(prt::f32::trace_kernel_f32<false, true, true>); break;
    case 4: launch(prt::f32::trace_kernel_f32<true, false, false>); break;
    case 5: launch(prt::f32::trace_kernel_f32<true, false, true>); break;
    case 6: launch(prt::f32::trace_kernel_f32<true, true, false>); break;
    default: launch(prt::f32::trace_kernel_f32<true, true, true>); break;
  }
  return cudaGetLastError();
}

cudaError_t prt_launch_gather_f32(const prt::GatherArgs* a, cudaStream_t st) {
  if (a->n_tiles == 0) return cudaSuccess;
  prt::f32::gather_kernel_f32<<<(unsigned)a->n_tiles, prt::kTileRays, 0, st>>>(*a);
  return cudaGetLastError();
}

}  // extern "C"
