// On-device read-out of a results frame: the reductions the reference's users
// run over the 15-column frame in pandas (examples/lens_design.ipynb cells 11-20) -- spot centroid /
// RMS per source on the imager, the focus (x-axis intercept) of every imager ray against its launch
// radius -- computed where the frame already is instead of copying 120 B per segment to the host.
//
// The frame is the column-major (generation, id)-ordered frame prt_gather_frame wrote:
// column c of row r at frame[c*stride + r], columns as pyrayt/_pyrayt.py:15.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

#include "../../include/pyrayt_b200.h"

namespace prt {

enum FrameCol {
  kColGeneration = 0, kColIntensity, kColWavelength, kColIndex, kColId, kColSurface,
  kColX0, kColY0, kColZ0, kColX1, kColY1, kColZ1, kColXt, kColYt, kColZt
};

__device__ __forceinline__ bool row_selected(const double* frame, long long stride, long long r, int select,
                                             double value) {
  if (select == PRT_SELECT_SURFACE) return frame[kColSurface * stride + r] == value;
  if (select == PRT_SELECT_GENERATION) return frame[kColGeneration * stride + r] == value;
  return true;
}

// results.loc[...]: -x_tilt * y0 / y_tilt + x0 (lens_design.ipynb cell 12), evaluated left to right like pandas
__device__ __forceinline__ double axis_focus(const double* frame, long long stride, long long r) {
  const double xt = frame[kColXt * stride + r], yt = frame[kColYt * stride + r];
  const double x0 = frame[kColX0 * stride + r], y0 = frame[kColY0 * stride + r];
  return __dadd_rn(__ddiv_rn(__dmul_rn(-xt, y0), yt), x0);
}

// order-preserving key so that min / max of doubles can use the integer atomics
__device__ __forceinline__ long long ordered_key(double v) {
  const long long b = __double_as_longlong(v);
  return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double from_ordered_key(long long k) {
  return __longlong_as_double(k >= 0 ? k : (k ^ 0x7fffffffffffffffLL));
}

constexpr int kSpotUnroll = 4;
constexpr int kSpotSliceRows = 8192;  // rows one block reduces by default
constexpr int kSelectRowsPerBlock = 1024;  // 256 threads x 4 rows

struct SpotAcc {
  double n, sy, sz, syy, szz, syz, nf, sf, sff, st, stt, ss2;
  double ymin, ymax, zmin, zmax;
  __device__ void clear() {
    n = sy = sz = syy = szz = syz = nf = sf = sff = st = stt = ss2 = 0.0;
    ymin = zmin = INFINITY;
    ymax = zmax = -INFINITY;
  }
};

__device__ __forceinline__ void spot_flush(double* s_acc, int g, const SpotAcc& a) {
  if (g < 0 || a.n == 0.0) return;
  double* o = s_acc + g * PRT_SPOT_COLS;
  atomicAdd(o + 0, a.n);
  atomicAdd(o + 1, a.sy);
  atomicAdd(o + 2, a.sz);
  atomicAdd(o + 3, a.syy);
  atomicAdd(o + 4, a.szz);
  atomicAdd(o + 5, a.syz);
  atomicMin(reinterpret_cast<long long*>(o + 6), ordered_key(a.ymin));
  atomicMax(reinterpret_cast<long long*>(o + 7), ordered_key(a.ymax));
  atomicMin(reinterpret_cast<long long*>(o + 8), ordered_key(a.zmin));
  atomicMax(reinterpret_cast<long long*>(o + 9), ordered_key(a.zmax));
  atomicAdd(o + 10, a.nf);
  atomicAdd(o + 11, a.sf);
  atomicAdd(o + 12, a.sff);
  atomicAdd(o + 13, a.st);
  atomicAdd(o + 14, a.stt);
  atomicAdd(o + 15, a.ss2);
}

// End-of-slice flush of a whole warp (all 32 lanes call it).  Lanes that hold the same group -- the usual
// case, groups being long id ranges -- are summed with shuffles first, so that a block issues 8 x 16
// shared-memory atomics instead of 256 x 16 on the same 16 words.
__device__ __forceinline__ void spot_flush_warp(double* s_acc, int g, SpotAcc a) {
  const unsigned full = 0xffffffffu;
  const int mine = a.n > 0.0 ? g : -1;
  const unsigned have = __ballot_sync(full, mine >= 0);
  if (have == 0u) return;
  const int lead = __shfl_sync(full, mine, __ffs(have) - 1);
  if (!__all_sync(full, mine < 0 || mine == lead)) {
    spot_flush(s_acc, mine, a);
    return;
  }
  if (mine < 0) a.clear();
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    a.n += __shfl_xor_sync(full, a.n, o);
    a.sy += __shfl_xor_sync(full, a.sy, o);
    a.sz += __shfl_xor_sync(full, a.sz, o);
    a.syy += __shfl_xor_sync(full, a.syy, o);
    a.szz += __shfl_xor_sync(full, a.szz, o);
    a.syz += __shfl_xor_sync(full, a.syz, o);
    a.nf += __shfl_xor_sync(full, a.nf, o);
    a.sf += __shfl_xor_sync(full, a.sf, o);
    a.sff += __shfl_xor_sync(full, a.sff, o);
    a.st += __shfl_xor_sync(full, a.st, o);
    a.stt += __shfl_xor_sync(full, a.stt, o);
    a.ss2 += __shfl_xor_sync(full, a.ss2, o);
    a.ymin = fmin(a.ymin, __shfl_xor_sync(full, a.ymin, o));
    a.ymax = fmax(a.ymax, __shfl_xor_sync(full, a.ymax, o));
    a.zmin = fmin(a.zmin, __shfl_xor_sync(full, a.zmin, o));
    a.zmax = fmax(a.zmax, __shfl_xor_sync(full, a.zmax, o));
  }
  if ((threadIdx.x & 31) == 0) spot_flush(s_acc, lead, a);
}

__device__ __forceinline__ bool is_key_col(int k) { return k >= 6 && k <= 9; }

// d_out must have been initialised by spot_init_kernel.  Each block reduces one contiguous slice of
// the rows (ids, hence groups, are contiguous inside a generation, so a thread rarely changes group)
// into shared memory and then adds its partial sums to d_out.
__global__ void __launch_bounds__(256, 3) spot_moments_kernel(const double* __restrict__ frame, long long rows,
                                                           long long stride, int select, double value,
                                                           long long rays_per_group, int n_groups,
                                                           const double* __restrict__ center,
                                                           double* __restrict__ out) {
  extern __shared__ double s_acc[];
  const int n_acc = n_groups * PRT_SPOT_COLS;
  for (int i = threadIdx.x; i < n_acc; i += blockDim.x) {
    const int k = i % PRT_SPOT_COLS;
    double v = 0.0;
    if (k == 6 || k == 8) v = __longlong_as_double(ordered_key(INFINITY));
    if (k == 7 || k == 9) v = __longlong_as_double(ordered_key(-INFINITY));
    s_acc[i] = v;
  }
  __syncthreads();

  const long long per_block = (rows + gridDim.x - 1) / gridDim.x;
  const long long r_begin = per_block * blockIdx.x;
  const long long r_end = min(rows, r_begin + per_block);
  SpotAcc a;
  a.clear();
  int cur = -1;
  double cy = 0.0, cz = 0.0, cf = 0.0, ct = 0.0;
  // every row reads the selection column only (kSpotUnroll independent loads per step, at four
  // blocks per SM, to keep HBM busy); the selected ones then read seven more.
  for (long long r0 = r_begin + threadIdx.x; r0 < r_end; r0 += (long long)kSpotUnroll * blockDim.x) {
    unsigned sel = 0u;
#pragma unroll
    for (int u = 0; u < kSpotUnroll; ++u) {
      const long long r = r0 + (long long)u * blockDim.x;
      sel |= (r < r_end && row_selected(frame, stride, r, select, value)) ? (1u << u) : 0u;
    }
#pragma unroll 1
    for (int u = 0; u < kSpotUnroll; ++u) {
      if (!((sel >> u) & 1u)) continue;
      const long long r = r0 + (long long)u * blockDim.x;
      const double v_id = frame[kColId * stride + r];
      const double y = frame[kColY1 * stride + r], z = frame[kColZ1 * stride + r];
      const double yt = frame[kColYt * stride + r], xt = frame[kColXt * stride + r];
      const double x0 = frame[kColX0 * stride + r], y0 = frame[kColY0 * stride + r];
      const long long id = (long long)v_id;
      const long long g64 = id / rays_per_group;  // calculate_source_ids (pyrayt/_pyrayt.py:316-327)
      if (id < 0 || g64 >= n_groups) continue;
      const int g = (int)g64;
      if (g != cur) {
        spot_flush(s_acc, cur, a);
        a.clear();
        cur = g;
        if (center) {
          cy = center[g * PRT_SPOT_CENTER_COLS + 0];
          cz = center[g * PRT_SPOT_CENTER_COLS + 1];
          cf = center[g * PRT_SPOT_CENTER_COLS + 2];
          ct = center[g * PRT_SPOT_CENTER_COLS + 3];
        }
      }
      // -x_tilt * y0 / y_tilt + x0 (lens_design.ipynb cell 12), evaluated left to right like pandas
      const double f = __dadd_rn(__ddiv_rn(__dmul_rn(-xt, y0), yt), x0);
      const double dy = y - cy, dz = z - cz;
      a.n += 1.0;
      a.sy += dy;
      a.sz += dz;
      a.syy += dy * dy;
      a.szz += dz * dz;
      a.syz += dy * dz;
      a.ymin = fmin(a.ymin, y);
      a.ymax = fmax(a.ymax, y);
      a.zmin = fmin(a.zmin, z);
      a.zmax = fmax(a.zmax, z);
      if (isfinite(f)) {
        const double df = f - cf;
        a.nf += 1.0;
        a.sf += df;
        a.sff += df * df;
      }
      const double dt = yt - ct, ds = sin(yt) - ct;  // cell 20: mean((sin(y_tilt) - sin(angle))^2)
      a.st += dt;
      a.stt += dt * dt;
      a.ss2 += ds * ds;
    }
  }
  spot_flush_warp(s_acc, cur, a);
  __syncthreads();
  for (int i = threadIdx.x; i < n_acc; i += blockDim.x) {
    if (s_acc[(i / PRT_SPOT_COLS) * PRT_SPOT_COLS] == 0.0) continue;  // no row of this group in the slice
    const int k = i % PRT_SPOT_COLS;
    if (is_key_col(k)) {
      const long long key = __double_as_longlong(s_acc[i]);
      if (k == 6 || k == 8)
        atomicMin(reinterpret_cast<long long*>(out + i), key);
      else
        atomicMax(reinterpret_cast<long long*>(out + i), key);
    } else if (s_acc[i] != 0.0) {
      atomicAdd(out + i, s_acc[i]);
    }
  }
}

__global__ void spot_init_kernel(double* out, int n_acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_acc) return;
  const int k = i % PRT_SPOT_COLS;
  double v = 0.0;
  if (k == 6 || k == 8) v = __longlong_as_double(ordered_key(INFINITY));
  if (k == 7 || k == 9) v = __longlong_as_double(ordered_key(-INFINITY));
  out[i] = v;
}

// min / max columns back from ordered keys to doubles
__global__ void spot_decode_kernel(double* out, int n_acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_acc) return;
  if (is_key_col(i % PRT_SPOT_COLS)) out[i] = from_ordered_key(__double_as_longlong(out[i]));
}

// centres for the second (centred) pass: previous centre + mean residual
__global__ void spot_centers_kernel(const double* sums, const double* center_in, int n_groups, double* center_out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  const double* s = sums + g * PRT_SPOT_COLS;
  double c[PRT_SPOT_CENTER_COLS] = {0.0, 0.0, 0.0, 0.0};
  if (center_in)
    for (int k = 0; k < PRT_SPOT_CENTER_COLS; ++k) c[k] = center_in[g * PRT_SPOT_CENTER_COLS + k];
  const double n = s[0], nf = s[10];
  if (n > 0.0) {
    c[0] += s[1] / n;
    c[1] += s[2] / n;
    c[3] += s[13] / n;
  }
  if (nf > 0.0) c[2] += s[11] / nf;
  for (int k = 0; k < PRT_SPOT_CENTER_COLS; ++k) center_out[g * PRT_SPOT_CENTER_COLS + k] = c[k];
}

// ---- focus table: one output row per selected frame row, in frame order (three passes: count, scan, write)

__global__ void __launch_bounds__(256) select_count_kernel(const double* __restrict__ frame, long long rows,
                                                           long long stride, int select, double value,
                                                           int* __restrict__ block_count) {
  const long long base = (long long)blockIdx.x * kSelectRowsPerBlock + threadIdx.x;
  int mine = 0;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long r = base + u * 256;
    mine += (r < rows && row_selected(frame, stride, r, select, value)) ? 1 : 0;
  }
  for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  __shared__ int s_warp[8];
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0;
    for (int w = 0; w < 8; ++w) c += s_warp[w];
    block_count[blockIdx.x] = c;
  }
}

__global__ void __launch_bounds__(256) axis_table_kernel(const double* __restrict__ frame, long long rows,
                                                         long long stride, int select, double value,
                                                         long long first_id, long long gen0_rows,
                                                         const long long* __restrict__ block_base,
                                                         double* __restrict__ table, long long table_stride,
                                                         long long table_capacity) {
  __shared__ int s_cnt[4 * 8];  // selected rows per (sub-chunk u, warp): frame order is u-major
  const long long base = (long long)blockIdx.x * kSelectRowsPerBlock + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool sel[4];
  unsigned mask[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long r = base + u * 256;
    sel[u] = r < rows && row_selected(frame, stride, r, select, value);
    mask[u] = __ballot_sync(0xffffffffu, sel[u]);
    if (lane == 0) s_cnt[u * 8 + warp] = __popc(mask[u]);
  }
  __syncthreads();
  const long long out0 = block_base[blockIdx.x];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (!sel[u]) continue;
    int before = 0;
    for (int k = 0; k < u * 8 + warp; ++k) before += s_cnt[k];
    const long long dst = out0 + before + __popc(mask[u] & ((1u << lane) - 1u));
    if (dst >= table_capacity) continue;
    const long long r = base + u * 256;
    const double id = frame[kColId * stride + r];
    // results.loc[(generation == 0) & id.isin(...)]['y0']: generation-0 rows are the first rows of the
    // frame in id order, but not one per ray -- a ray that misses everything in generation 0 has no row --
    // so the row is found by bisection on the id column (first_id only seeds the search: row id - first_id
    // when no earlier ray missed)
    double radius = NAN;
    {
      const double* ids = frame + kColId * stride;
      long long lo = 0, hi = gen0_rows;  // first row in [0, gen0_rows) with ids[row] >= id
      const long long guess = (long long)id - first_id;
      if (guess >= 0 && guess < gen0_rows && ids[guess] == id) {
        lo = hi = guess;
      }
      while (lo < hi) {
        const long long mid = lo + ((hi - lo) >> 1);
        if (ids[mid] < id) lo = mid + 1; else hi = mid;
      }
      if (lo < gen0_rows && ids[lo] == id && frame[kColGeneration * stride + lo] == 0.0)
        radius = frame[kColY0 * stride + lo];
    }
    table[0 * table_stride + dst] = id;
    table[1 * table_stride + dst] = radius;
    table[2 * table_stride + dst] = axis_focus(frame, stride, r);
    table[3 * table_stride + dst] = frame[kColWavelength * stride + r];
  }
}

}  // namespace prt

extern "C" {

cudaError_t prt_launch_scan(const int* run_count, long long* run_base, long long n_tiles, int generation_limit,
                            long long* gen_offsets, cudaStream_t st);

cudaError_t prt_launch_spot_moments(const double* frame, long long rows, long long stride, int select, double value,
                                    long long rays_per_group, int n_groups, const double* center, double* out,
                                    int blocks, cudaStream_t st) {
  const int n_acc = n_groups * PRT_SPOT_COLS;
  prt::spot_init_kernel<<<(n_acc + 255) / 256, 256, 0, st>>>(out, n_acc);
  if (rows > 0) {
    long long grid = (rows + prt::kSpotSliceRows - 1) / prt::kSpotSliceRows;  // default: short slices, because the
    if (blocks > 0 && blocks < grid) grid = blocks;                           // selected rows cluster in late generations
    if (grid > 0x7fffffffLL) grid = 0x7fffffffLL;
    prt::spot_moments_kernel<<<(unsigned)grid, 256, (size_t)n_acc * sizeof(double), st>>>(frame, rows, stride, select, value,
                                                                               rays_per_group, n_groups, center, out);
  }
  prt::spot_decode_kernel<<<(n_acc + 255) / 256, 256, 0, st>>>(out, n_acc);
  return cudaGetLastError();
}

cudaError_t prt_launch_spot_centers(const double* sums, const double* center_in, int n_groups, double* center_out,
                                    cudaStream_t st) {
  prt::spot_centers_kernel<<<(n_groups + 127) / 128, 128, 0, st>>>(sums, center_in, n_groups, center_out);
  return cudaGetLastError();
}

cudaError_t prt_launch_axis_table(const double* frame, long long rows, long long stride, int select, double value,
                                  long long first_id, long long gen0_rows, int* block_count, long long* block_base,
                                  long long* total, double* table, long long table_stride, long long table_capacity,
                                  cudaStream_t st) {
  const long long blocks = (rows + prt::kSelectRowsPerBlock - 1) / prt::kSelectRowsPerBlock;
  if (blocks == 0) return cudaMemsetAsync(total, 0, 2 * sizeof(long long), st);
  prt::select_count_kernel<<<(unsigned)blocks, 256, 0, st>>>(frame, rows, stride, select, value, block_count);
  // one "generation" of `blocks` runs: exclusive scan into block_base, total into total[1] (total[0] = 0)
  cudaError_t e = prt_launch_scan(block_count, block_base, blocks, 1, total, st);
  if (e != cudaSuccess) return e;
  prt::axis_table_kernel<<<(unsigned)blocks, 256, 0, st>>>(frame, rows, stride, select, value, first_id, gen0_rows,
                                                          block_base, table, table_stride, table_capacity);
  return cudaGetLastError();
}

}  // extern "C"
