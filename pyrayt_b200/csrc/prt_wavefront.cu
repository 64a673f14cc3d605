// Wavefront form of the generation loop: one launch per generation, rows written in place.
//
// prt_trace (prt_kernels.cu) runs every generation of a ray inside one thread; it cannot know where a
// row belongs in the (generation, id)-ordered frame, so rows go to a staging buffer and a second pass
// moves them: 240 B of device memory and of HBM traffic per row.  Here the ray state lives in HBM and
// the generations are separate launches of one kernel:
//
//   step g   finishes generation g-1 for every ray that hit something (_st_interact: material, new
//            direction, the row, the new state) and starts generation g for every ray that goes on
//            (_st_propagate: nearest hit), counting per tile the rows generation g will produce;
//   scan g   turns those counts into frame positions (exclusive scan over the tiles, first row of
//            generation g+1).
//
// Because step g+1 knows the exact position of every row of generation g, rows are written straight
// to their final place: no staging buffer, no ordering pass.  All launches are enqueued at once; a
// launch for a generation no ray reaches returns immediately (device-side count of survivors), so the
// host never synchronises inside a trace.  The per-ray arithmetic is the same PRT_HD code as the
// single-kernel path (prt_device.cuh): frames are bit-identical.
//
// Measured on config 4 (2^24 rays): 73.3 ms per trace against 70.9 ms for prt_trace + scan + gather:
// the state round trip through HBM costs what the missing gather saves.  It is the path for ray sets
// whose staging buffer + frame would not fit the device (Engine.trace(method="auto")).
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/pyrayt_b200.h"
#include "prt_scene.h"
#include "prt_device.cuh"

namespace prt {

#ifndef PRT_WAVE_TILE
#define PRT_WAVE_TILE 256
#endif
constexpr int kWaveTile = PRT_WAVE_TILE;

// per-ray flag word: bits 0-7 skip component + 1, bit 8 a tie was already counted, bit 9 dead
constexpr int kFlagTie = 1 << 8, kFlagDead = 1 << 9;
constexpr int kLeafDead = -2;  // hit_leaf: the ray took no step this generation

__device__ __forceinline__ void stage_blob(unsigned char* s_blob, const unsigned char* blob, int blob_bytes) {
  const int words = blob_bytes / 8;
  const double* src = reinterpret_cast<const double*>(blob);
  double* dst = reinterpret_cast<double*>(s_blob);
  for (int w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
  __syncthreads();
}

__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void count(uint64_t* dst, unsigned long long v, int lane) {
  const unsigned long long s = warp_sum64(v);
  if (lane == 0 && s) atomicAdd(reinterpret_cast<unsigned long long*>(dst), s);
}

// ---------------------------------------------------------------- asynchronous tile prefetch
//
// What a thread needs at the top of a tile (its ray's previous hit, state, wavelength and the columns
// a row copies from the input) sits in 14 different arrays; loaded on demand, three dependent round
// trips to HBM open every tile while the tile's barriers keep the block's warps in step.  Instead each
// thread copies the values of its ray in the *next* tile into shared memory with cp.async while the
// current tile is computed (two buffers).  A thread only ever reads the slots it copied itself, so
// waiting for its own copy group is all the synchronisation the buffers need.
constexpr int kPreDoubles = 12;  // hit_t, state[7], wavelength, first generation, intensity, id
constexpr int kPreInts = 2;      // hit_leaf, flag
constexpr int kPreBytes = kWaveTile * (kPreDoubles * 8 + kPreInts * 4);

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct PreBuf {
  double* d;  // [kPreDoubles][kWaveTile]
  int* k;     // [kPreInts][kWaveTile]
};

__device__ __forceinline__ PreBuf pre_buf(unsigned char* base, int which) {
  unsigned char* b = base + (size_t)which * kPreBytes;
  PreBuf p;
  p.d = reinterpret_cast<double*>(b);
  p.k = reinterpret_cast<int*>(b + (size_t)kWaveTile * kPreDoubles * 8);
  return p;
}

// start the copies for ray i of a tile (generation g >= 1 data); always commits a group
__device__ __forceinline__ void pre_issue(const WaveArgs& a, const PreBuf& p, long long i) {
  const int t = threadIdx.x;
  if (i < a.n_rays) {
    cp_async4(p.k + 0 * kWaveTile + t, a.hit_leaf + i);
    cp_async4(p.k + 1 * kWaveTile + t, a.flag + i);
    cp_async8(p.d + 0 * kWaveTile + t, a.hit_t + i);
#pragma unroll
    for (int q = 0; q < 7; ++q) cp_async8(p.d + (1 + q) * kWaveTile + t, a.st + q * a.n_rays + i);
    cp_async8(p.d + 8 * kWaveTile + t, a.rays + 10 * a.stride + i);
    cp_async8(p.d + 9 * kWaveTile + t, a.rays + 8 * a.stride + i);
    cp_async8(p.d + 10 * kWaveTile + t, a.rays + 9 * a.stride + i);
    cp_async8(p.d + 11 * kWaveTile + t, a.rays + 12 * a.stride + i);
  }
  cp_async_commit();
}

// ---------------------------------------------------------------- one launch per generation
//
// Launch g finishes generation g-1 (interaction, row, new state) and starts generation g (nearest
// hit, per-tile row count) for every ray, so a ray's state is read and written once per generation
// and the arithmetic of the single-kernel loop body is kept together.  The rows of generation g-1
// can be placed exactly because launch g-1 counted them per tile and the scan in between turned the
// counts into positions.  Tile counts / bases are double-buffered by generation parity.
#ifndef PRT_WAVE_MIN_BLOCKS
#define PRT_WAVE_MIN_BLOCKS 2
#endif
template <bool GENERIC>
__global__ void __launch_bounds__(kWaveTile, PRT_WAVE_MIN_BLOCKS) wave_step_kernel(const WaveArgs a) {
  extern __shared__ __align__(16) unsigned char s_blob[];
  __shared__ int s_wcount[kWaveTile / 32];
  const int g = a.g;
  if (g >= 2 && a.alive[g - 1] == 0) return;  // no ray entered generation g-1: nothing to finish or start
  stage_blob(s_blob, a.blob, a.blob_bytes);    // once per block: the blocks are persistent and walk the tiles
  const SceneView sc = make_view(s_blob);
  unsigned char* pre_base = s_blob + ((a.blob_bytes + 15) & ~15);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long first_row = g > 0 ? a.gen_off[g - 1] : 0;
  const long long* base_prev = a.blk_base + (long long)((g + 1) & 1) * a.n_tiles;  // of generation g-1
  int* count_cur = a.blk_count + (long long)(g & 1) * a.n_tiles;                    // of generation g
  unsigned n_gen = 0, n_tie = 0, n_nan = 0, n_rays = 0, n_badw = 0;
  unsigned n_alive = 0, n_seg = 0, n_rows = 0, n_drop = 0, n_untr = 0, n_lim = 0, n_abs = 0, n_mir = 0;
  int which = 0;
  if (g > 0 && (long long)blockIdx.x < a.n_tiles)
    pre_issue(a, pre_buf(pre_base, 0), (long long)blockIdx.x * kWaveTile + threadIdx.x);
  for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, which ^= 1) {
    const long long i = tile * kWaveTile + threadIdx.x;
    const bool valid = i < a.n_rays;
    RayState rs = {0, 0, 0, 0, 0, 0, 0, 1, -1};
    int flag = kFlagDead;
    bool alive = false;
    double row_gen0 = 0, row_inten = 0, row_id = 0;
    if (g > 0) {  // copies for the next tile go out before this one is touched
      const long long nt = tile + gridDim.x;
      if (nt < a.n_tiles) {
        pre_issue(a, pre_buf(pre_base, which ^ 1), nt * kWaveTile + threadIdx.x);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
    }
    const PreBuf pb = pre_buf(pre_base, which);
    if (g == 0) {
      if (valid) {
        const double* r = a.rays + i;
        rs.p0 = r[0 * a.stride];
        rs.p1 = r[1 * a.stride];
        rs.p2 = r[2 * a.stride];
        rs.v0 = r[4 * a.stride];
        rs.v1 = r[5 * a.stride];
        rs.v2 = r[6 * a.stride];
        rs.nidx = r[11 * a.stride];
        if (r[3 * a.stride] != 1.0 || r[7 * a.stride] != 0.0) n_badw += 1;
        n_rays += 1;
        flag = 0;
        alive = true;
      }
    } else {
      // ---- finish generation g-1: _st_interact for the rays that hit something
      const int t = threadIdx.x;
      const int leaf = valid ? pb.k[0 * kWaveTile + t] : kLeafDead;
      StepCounters c = {0, 0};
      StepOut so;
      so.row = false;
      if (leaf != kLeafDead) flag = pb.k[1 * kWaveTile + t];
      if (leaf >= 0) {
        rs.p0 = pb.d[1 * kWaveTile + t];
        rs.p1 = pb.d[2 * kWaveTile + t];
        rs.p2 = pb.d[3 * kWaveTile + t];
        rs.v0 = pb.d[4 * kWaveTile + t];
        rs.v1 = pb.d[5 * kWaveTile + t];
        rs.v2 = pb.d[6 * kWaveTile + t];
        rs.nidx = pb.d[7 * kWaveTile + t];
        rs.wl = pb.d[8 * kWaveTile + t];
        row_gen0 = pb.d[9 * kWaveTile + t];
        row_inten = pb.d[10 * kWaveTile + t];
        row_id = pb.d[11 * kWaveTile + t];
        const double vn = sqrt(rs.v0 * rs.v0 + rs.v1 * rs.v1 + rs.v2 * rs.v2);  // as step_speed
        alive = step_interact(sc, rs, g - 1, a.generation_limit, vn, pb.d[0 * kWaveTile + t], leaf, so, c);
      }
      const bool write = so.row && (a.record_mode == PRT_RECORD_ALL || so.sid == a.detector_sid);
      const unsigned m = __ballot_sync(0xffffffffu, write);
      if (lane == 0) s_wcount[warp] = __popc(m);
      __syncthreads();
      if (write) {
        int before = 0;
        for (int w = 0; w < warp; ++w) before += s_wcount[w];
        const long long row = first_row + base_prev[tile] + before + __popc(m & ((1u << lane) - 1u));
        if (row < a.capacity) {
          double* o = a.frame + row;
          const long long cs = a.frame_stride;
          __stcs(o + 0 * cs, (g == 1) ? row_gen0 : (double)(g - 1));  // :440-441
          __stcs(o + 1 * cs, row_inten);
          __stcs(o + 2 * cs, rs.wl);
          __stcs(o + 3 * cs, rs.nidx);
          __stcs(o + 4 * cs, row_id);
          __stcs(o + 5 * cs, so.sid);
          __stcs(o + 6 * cs, rs.p0);
          __stcs(o + 7 * cs, rs.p1);
          __stcs(o + 8 * cs, rs.p2);
          __stcs(o + 9 * cs, so.e0);
          __stcs(o + 10 * cs, so.e1);
          __stcs(o + 11 * cs, so.e2);
          __stcs(o + 12 * cs, so.t0n);
          __stcs(o + 13 * cs, so.t1n);
          __stcs(o + 14 * cs, so.t2n);
        } else {
          n_drop += 1;
        }
        n_rows += 1;
      }
      n_seg += c.w0 >> 16;
      n_untr += (c.w1 & kCtrUntr) ? 1u : 0u;
      n_lim += (c.w1 & kCtrLim) ? 1u : 0u;
      n_abs += (c.w1 & kCtrAbs) ? 1u : 0u;
      n_mir += c.w1 & 0xffffu;
      if (alive) {
        advance_ray(rs, so, g - 1, a.ray_offset);
        flag = (flag & kFlagTie) | ((rs.skip + 1) & 0xff);
        n_alive += 1;
      } else if (leaf != kLeafDead) {
        flag = kFlagDead;  // miss, absorbed, generation limit, untraceable surface
      }
    }
    // ---- start generation g: _st_propagate for the rays that go on
    int leaf = kLeafDead;
    double best_t = PRT_INF;
    bool row = false;
    if (alive && g < a.generation_limit) {
      StepCounters c = {0, 0};
      rs.skip = (flag & 0xff) - 1;
      const double vn = step_speed(rs, c);
      if (vn != 0.0) {
        typename StackFor<GENERIC>::type stack_storage;
        HitStack* S = StackFor<GENERIC>::ptr(stack_storage);
        bool tie = false;
        nearest_hit<GENERIC, false>(sc, rs.p0, rs.p1, rs.p2, rs.v0, rs.v1, rs.v2, rs.skip, S, best_t, leaf, tie);
        if (tie && !(flag & kFlagTie)) {
          flag |= kFlagTie;
          n_tie += 1;
        }
        if (leaf >= 0) {
          const Leaf& L = sc.leaves[leaf];
          const bool traceable = L.mat == PRT_MAT_ABSORBER || L.mat == PRT_MAT_MIRROR ||
                                 L.mat == PRT_MAT_GLASS_CONST || L.mat == PRT_MAT_GLASS_SELLMEIER;
          row = traceable && (a.record_mode == PRT_RECORD_ALL || L.sid == a.detector_sid);
        }
      } else {
        flag = kFlagDead;  // zero / NaN direction: no step, no row (_pyrayt.py:415)
      }
      n_gen += c.w0 & 0xffffu;
      n_nan += (c.w1 & kCtrNan) ? 1u : 0u;
      a.st[0 * a.n_rays + i] = rs.p0;
      a.st[1 * a.n_rays + i] = rs.p1;
      a.st[2 * a.n_rays + i] = rs.p2;
      a.st[3 * a.n_rays + i] = rs.v0;
      a.st[4 * a.n_rays + i] = rs.v1;
      a.st[5 * a.n_rays + i] = rs.v2;
      a.st[6 * a.n_rays + i] = rs.nidx;
    }
    if (valid && g < a.generation_limit) {
      a.hit_t[i] = best_t;
      a.hit_leaf[i] = leaf;
      a.flag[i] = flag;
    }
    const int rows = __syncthreads_count(row);  // (also fences s_wcount for the next tile)
    if (threadIdx.x == 0) count_cur[tile] = rows;
  }
  if (g > 0) count(reinterpret_cast<uint64_t*>(&a.alive[g]), n_alive, lane);
  count(&a.ctr->generations, n_gen, lane);
  count(&a.ctr->tie_rays, n_tie, lane);
  count(&a.ctr->nan_rays, n_nan, lane);
  count(&a.ctr->rays, n_rays, lane);
  count(&a.ctr->bad_w, n_badw, lane);
  count(&a.ctr->segments, n_seg, lane);
  count(&a.ctr->rows_reserved, n_rows, lane);
  count(&a.ctr->rows_dropped, n_drop, lane);
  count(&a.ctr->untraceable_hits, n_untr, lane);
  count(&a.ctr->limit_rays, n_lim, lane);
  count(&a.ctr->absorber_segments, n_abs, lane);
  count(&a.ctr->mirror_segments, n_mir, lane);
}

// ---------------------------------------------------------------- tile counts of generation g -> frame positions
__global__ void __launch_bounds__(1024) wave_scan_kernel(const WaveArgs a) {
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  const int g = a.g;
  if (g >= 2 && a.alive[g - 1] == 0) {  // launch g returned at once: no rows in generation g
    if (threadIdx.x == 0) a.gen_off[g + 1] = a.gen_off[g];
    return;
  }
  const int* cnt = a.blk_count + (long long)(g & 1) * a.n_tiles;
  long long* base = a.blk_base + (long long)(g & 1) * a.n_tiles;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long t0 = 0; t0 < a.n_tiles; t0 += blockDim.x) {
    const long long t = t0 + threadIdx.x;
    const long long c = (t < a.n_tiles) ? cnt[t] : 0;
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
    if (t < a.n_tiles) base[t] = carry + (warp ? s_warp[warp - 1] : 0) + x - c;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = carry + s_warp[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (g == 0) a.gen_off[0] = 0;
    a.gen_off[g + 1] = (g == 0 ? 0 : a.gen_off[g]) + s_carry;
  }
}

}  // namespace prt

extern "C" {

int prt_wave_tile(void) { return prt::kWaveTile; }

// enqueue every generation; events (optional): 2 per launch of the step kernel (generation_limit + 1 launches)
cudaError_t prt_launch_wavefront(prt::WaveArgs* a, int generic, cudaEvent_t* events, cudaStream_t st) {
  const long long tiles = (a->n_rays + prt::kWaveTile - 1) / prt::kWaveTile;
  if (tiles == 0) return cudaMemsetAsync(a->gen_off, 0, sizeof(long long) * (a->generation_limit + 1), st);
  const size_t smem = (size_t)((a->blob_bytes + 15) & ~15) + 2 * (size_t)prt::kPreBytes;  // blob + two prefetch buffers
  cudaFuncSetAttribute(prt::wave_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(prt::wave_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaError_t e = cudaMemsetAsync(a->alive, 0, sizeof(long long) * (a->generation_limit + 1), st);
  if (e != cudaSuccess) return e;
  a->n_tiles = tiles;
  // persistent blocks (the scene blob is staged once per block and launch): resident blocks x SMs
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long want = (long long)sms * PRT_WAVE_MIN_BLOCKS;
  const unsigned grid = (unsigned)(tiles < want ? tiles : want);
  for (int g = 0; g <= a->generation_limit; ++g) {
    a->g = g;
    if (events) cudaEventRecord(events[2 * g], st);
    if (generic)
      prt::wave_step_kernel<true><<<grid, prt::kWaveTile, smem, st>>>(*a);
    else
      prt::wave_step_kernel<false><<<grid, prt::kWaveTile, smem, st>>>(*a);
    if (events) cudaEventRecord(events[2 * g + 1], st);
    if (g < a->generation_limit) prt::wave_scan_kernel<<<1, 1024, 0, st>>>(*a);
  }
  return cudaGetLastError();
}

}  // extern "C"
