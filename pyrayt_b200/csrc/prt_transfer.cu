// Lean device -> host transfer of a results frame.
//
// Copying the (15, rows) frame to the host is PCIe-bound (120 B per row), and five of its columns
// carry no information of their own: `generation` follows from the row's position, `intensity`,
// `wavelength` and `id` are copies of the ray's input values, `surface` is a small integer.  The pack
// kernel folds them into one 64-bit word per row -- (index of the ray in the input RaySet) << 24 |
// (surface id + 1) -- after *verifying* row by row that the frame's values (all five columns, the
// generation included) are exactly the ones the host will reconstruct; the host then fills those five columns from its own copy of the rays while
// the other ten columns are still streaming over the bus (88 instead of 120 B per row on PCIe).
// Any row that does not verify (ids that are not consecutive, surface ids outside 24 bits, NaN
// metadata) makes the caller fall back to copying all fifteen columns.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include "../../include/pyrayt_b200.h"

namespace prt {

__global__ void __launch_bounds__(256) frame_pack_kernel(const double* __restrict__ frame, long long rows,
                                                         long long stride, const double* __restrict__ rays,
                                                         long long n_rays, long long ray_stride,
                                                         const long long* __restrict__ gen_off, int generations,
                                                         unsigned long long* __restrict__ packed,
                                                         unsigned long long* __restrict__ bad) {
  const double id0 = rays[12 * ray_stride];
  unsigned my_bad = 0;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows;
       r += (long long)gridDim.x * blockDim.x) {
    const double id = frame[4 * stride + r], sid = frame[5 * stride + r];
    const double rel = id - id0;
    const long long idx = (rel >= 0.0 && rel < 1099511627776.0) ? (long long)rel : -1;  // < 2^40
    const long long s = (sid >= -1.0 && sid < 16777215.0) ? (long long)sid : -2;
    // (the host rebuilds id as id0 + index: both directions of that identity are checked)
    bool ok = idx >= 0 && idx < n_rays && (double)idx == rel && id0 + (double)idx == id && s >= -1 && (double)s == sid;
    if (ok) {
      // generation of row r = the last g with gen_off[g] <= r (the host walks the same offsets); the column
      // must hold the ray's own generation value in generation 0 and g afterwards (pyrayt/_pyrayt.py:440-441)
      int lo = 0, hi = generations;  // gen_off[0] = 0 <= r < gen_off[generations] = rows
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (gen_off[mid] <= r) lo = mid; else hi = mid;
      }
      const double want_gen = (lo == 0) ? rays[8 * ray_stride + idx] : (double)lo;
      ok = rays[12 * ray_stride + idx] == id && rays[9 * ray_stride + idx] == frame[1 * stride + r] &&
           rays[10 * ray_stride + idx] == frame[2 * stride + r] && frame[0 * stride + r] == want_gen;
    }
    packed[r] = ok ? (((unsigned long long)idx << 24) | (unsigned long long)(s + 1)) : ~0ull;
    my_bad += ok ? 0u : 1u;
  }
  for (int o = 16; o; o >>= 1) my_bad += __shfl_xor_sync(0xffffffffu, my_bad, o);
  if ((threadIdx.x & 31) == 0 && my_bad) atomicAdd(bad, (unsigned long long)my_bad);
}

}  // namespace prt

// The rebuilt columns are written once and not read again by these threads: store them around the
// cache (no read-for-ownership of 12 GB of destination lines while the DMA engine writes next to them).
static inline void store_stream(double* p, double v) {
#if defined(__x86_64__)
  long long bits;
  std::memcpy(&bits, &v, 8);
  _mm_stream_si64(reinterpret_cast<long long*>(p), bits);
#else
  *p = v;
#endif
}

extern "C" {

cudaError_t prt_launch_frame_pack(const double* frame, long long rows, long long stride, const double* rays,
                                  long long n_rays, long long ray_stride, const long long* gen_off, int generations,
                                  unsigned long long* packed, unsigned long long* bad, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(bad, 0, sizeof(unsigned long long), st);
  if (e != cudaSuccess || rows == 0) return e;
  const long long want = (rows + 255) / 256;
  const unsigned grid = (unsigned)std::min<long long>(want, 148LL * 32);
  prt::frame_pack_kernel<<<grid, 256, 0, st>>>(frame, rows, stride, rays, n_rays, ray_stride, gen_off, generations,
                                               packed, bad);
  return cudaGetLastError();
}

// host side: columns generation / intensity / wavelength / id / surface of rows [0, rows) from the packed words
void prt_host_expand_rows(const uint64_t* packed, int64_t r0, int64_t r1, const int64_t* gen_off, int32_t generations,
                          const double* r_gen, const double* r_int, const double* r_wl, const double* r_id,
                          double* frame, int64_t frame_stride) {
  double* f_gen = frame + 0 * frame_stride;
  double* f_int = frame + 1 * frame_stride;
  double* f_wl = frame + 2 * frame_stride;
  double* f_id = frame + 4 * frame_stride;
  double* f_sid = frame + 5 * frame_stride;
  const double id0 = r_id[0];
  // generation of row r0: last g with gen_off[g] <= r0
  int g = (int)(std::upper_bound(gen_off, gen_off + generations + 1, r0) - gen_off) - 1;
  int64_t r = r0;
  while (r < r1) {
    while (g + 1 <= generations && gen_off[g + 1] <= r) ++g;
    const int64_t end = std::min<int64_t>(r1, g + 1 <= generations ? gen_off[g + 1] : r1);
    for (; r < end; ++r) {
      const uint64_t p = packed[r];
      const int64_t idx = (int64_t)(p >> 24);
      store_stream(f_gen + r, (g == 0) ? r_gen[idx] : (double)g);  // pyrayt/_pyrayt.py:440-441
      store_stream(f_int + r, r_int[idx]);
      store_stream(f_wl + r, r_wl[idx]);
      store_stream(f_id + r, id0 + (double)idx);  // ids are consecutive (verified by the pack kernel)
      store_stream(f_sid + r, (double)((int64_t)(p & 0xffffffu) - 1));
    }
  }
#if defined(__x86_64__)
  _mm_sfence();
#endif
}

}  // extern "C"
