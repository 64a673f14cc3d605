// C-ABI layer of pyrayt_b200 (include/pyrayt_b200.h): argument validation, scene
// encoding (postfix CSG -> preorder program with bounding-box skip targets) and
// kernel launches.  No torch types; device buffers belong to the caller.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pyrayt_b200.h"
#include "prt_scene.h"
#include "prt_encode.h"


extern "C" {
cudaError_t prt_launch_trace(const prt::TraceArgs* a, int record, int generic, int diagnose, cudaStream_t st);
cudaError_t prt_launch_trace_f32(const prt::TraceArgs* a, int record, int ordered, int generic, int n_leaves,
                                 int n_components, int n_aabb, cudaStream_t st);
cudaError_t prt_launch_gather_f32(const prt::GatherArgs* a, cudaStream_t st);
size_t prt_f32_smem_bytes(int blob_bytes, int n_leaves, int n_components, int n_aabb);
cudaError_t prt_launch_scan(const int* run_count, long long* run_base, long long n_tiles, int generation_limit,
                            long long* gen_offsets, cudaStream_t st);
cudaError_t prt_launch_gather(const prt::GatherArgs* a, int layout, cudaStream_t st);
cudaError_t prt_launch_intersect(const unsigned char* blob, int blob_bytes, int component, const double* rays,
                                 long long n, double* hits, long long* sids, int slots, cudaStream_t st);
cudaError_t prt_launch_source(const prt_source_desc* src, double* rays, long long n, long long stride,
                              long long first, cudaStream_t st);
cudaError_t prt_launch_fp64_probe(double* out, int blocks, int iters, cudaStream_t st);
cudaError_t prt_launch_render_hit(const unsigned char* blob, int blob_bytes, const double* rays, long long n,
                                  double* t_out, long long* sid_out, double* normals, cudaStream_t st);
cudaError_t prt_launch_spot_moments(const double* frame, long long rows, long long stride, int select, double value,
                                    long long rays_per_group, int n_groups, const double* center, double* out,
                                    int blocks, cudaStream_t st);
cudaError_t prt_launch_spot_centers(const double* sums, const double* center_in, int n_groups, double* center_out,
                                    cudaStream_t st);
cudaError_t prt_launch_axis_table(const double* frame, long long rows, long long stride, int select, double value,
                                  long long first_id, long long gen0_rows, int* block_count, long long* block_base,
                                  long long* total, double* table, long long table_stride, long long table_capacity,
                                  cudaStream_t st);
cudaError_t prt_launch_frame_pack(const double* frame, long long rows, long long stride, const double* rays,
                                  long long n_rays, long long ray_stride, const long long* gen_off, int generations,
                                  unsigned long long* packed, unsigned long long* bad, cudaStream_t st);
void prt_host_expand_rows(const uint64_t* packed, int64_t r0, int64_t r1, const int64_t* gen_off, int32_t generations,
                          const double* r_gen, const double* r_int, const double* r_wl, const double* r_id,
                          double* frame, int64_t frame_stride);
cudaError_t prt_launch_nearest(const unsigned char* blob, int blob_bytes, const double* rays, long long n, double* t_out,
                               long long* sid_out, double* normals, cudaStream_t st);
}

extern "C" cudaError_t prt_launch_wavefront(prt::WaveArgs* a, int generic, cudaEvent_t* events, cudaStream_t st);

struct prt_scene {
  int device = 0;
  unsigned char* d_blob = nullptr;
  int blob_bytes = 0;
  int n_leaves = 0;
  int n_components = 0;
  std::vector<int> comp_slots;
  int generic = 1;  // some component needs the interpreter for arbitrary CSG trees
  int ordered = 0;  // the encoder chose the ray-ordered traversal (many boxed components)
  int n_aabb = 0;   // node boxes of the interpreter's programs
  std::vector<unsigned char> staging;  // host copy of the last prt_scene_update (pageable -> the copy is synchronous enough)
};

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return PRT_ERR_CUDA;
}

}  // namespace

extern "C" {

int prt_abi_version(void) { return PRT_ABI_VERSION; }
const char* prt_last_error(void) { return g_err.c_str(); }
int prt_tile_rays(void) { return prt::kTileRays; }

int prt_scene_create(const prt_scene_desc* d, int device, prt_scene** out) {
  if (!d || !out) return fail(PRT_ERR_INVALID, "null argument");
  *out = nullptr;
  std::vector<unsigned char> blob;
  std::vector<int> comp_slots;
  {
    std::string err;
    const int rc = prt::encode_scene(d, blob, comp_slots, err);
    if (rc != PRT_OK) return fail(rc, err);
  }
  const int off = (int)blob.size();

  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  prt_scene* sc = new prt_scene();
  sc->device = device;
  sc->blob_bytes = off;
  sc->n_leaves = d->n_leaves;
  sc->n_components = d->n_components;
  sc->comp_slots = comp_slots;
  sc->generic = (reinterpret_cast<const prt::BlobHeader*>(blob.data())->flags & 2) ? 1 : 0;
  {
    const prt::BlobHeader* bh = reinterpret_cast<const prt::BlobHeader*>(blob.data());
    sc->ordered = (bh->n_boxed > 0 && (bh->flags & 4) && (bh->flags & 8)) ? 1 : 0;
    sc->n_aabb = bh->n_aabb;
  }
  e = cudaMalloc(&sc->d_blob, (size_t)off);
  if (e != cudaSuccess) {
    delete sc;
    return cuda_fail(e, "cudaMalloc(scene)");
  }
  e = cudaMemcpy(sc->d_blob, blob.data(), (size_t)off, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(sc->d_blob);
    delete sc;
    return cuda_fail(e, "cudaMemcpy(scene)");
  }
  *out = sc;
  return PRT_OK;
}

int prt_scene_update(prt_scene* sc, const prt_scene_desc* d, void* cuda_stream) {
  if (!sc || !d) return fail(PRT_ERR_INVALID, "null argument");
  std::vector<unsigned char> blob;
  std::vector<int> comp_slots;
  {
    std::string err;
    const int rc = prt::encode_scene(d, blob, comp_slots, err);
    if (rc != PRT_OK) return fail(rc, err);
  }
  cudaError_t e = cudaSetDevice(sc->device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  if ((int)blob.size() != sc->blob_bytes) {  // topology changed: new device buffer
    unsigned char* fresh = nullptr;
    e = cudaMalloc(&fresh, blob.size());
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(scene)");
    cudaStreamSynchronize((cudaStream_t)cuda_stream);  // kernels still reading the old blob
    cudaFree(sc->d_blob);
    sc->d_blob = fresh;
    sc->blob_bytes = (int)blob.size();
  }
  sc->staging = blob;  // keeps the host copy alive until the async copy has run
  e = cudaMemcpyAsync(sc->d_blob, sc->staging.data(), sc->staging.size(), cudaMemcpyHostToDevice,
                      (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(scene)");
  sc->n_leaves = d->n_leaves;
  sc->n_components = d->n_components;
  sc->comp_slots = comp_slots;
  sc->generic = (reinterpret_cast<const prt::BlobHeader*>(sc->staging.data())->flags & 2) ? 1 : 0;
  {
    const prt::BlobHeader* bh = reinterpret_cast<const prt::BlobHeader*>(sc->staging.data());
    sc->ordered = (bh->n_boxed > 0 && (bh->flags & 4) && (bh->flags & 8)) ? 1 : 0;
    sc->n_aabb = bh->n_aabb;
  }
  return PRT_OK;
}

void prt_scene_destroy(prt_scene* scene) {
  if (!scene) return;
  if (scene->d_blob) cudaFree(scene->d_blob);
  delete scene;
}

int prt_scene_n_leaves(const prt_scene* scene) { return scene ? scene->n_leaves : 0; }

int prt_trace(prt_scene* scene, const prt_params* p, const double* d_rays, int64_t n_rays, int64_t ray_stride,
              const prt_records* rec, prt_counters* d_counters, void* cuda_stream) {
  if (!scene || !p || !d_counters) return fail(PRT_ERR_INVALID, "null argument");
  if (n_rays < 0 || (n_rays > 0 && !d_rays)) return fail(PRT_ERR_INVALID, "bad ray buffer");
  if (ray_stride < n_rays) return fail(PRT_ERR_INVALID, "ray_stride < n_rays");
  if (p->generation_limit < 1) return fail(PRT_ERR_INVALID, "generation_limit must be >= 1");
  if (p->generation_limit > 65535) return fail(PRT_ERR_LIMIT, "generation_limit must be <= 65535");
  if (p->record_mode < PRT_RECORD_ALL || p->record_mode > PRT_RECORD_NONE)
    return fail(PRT_ERR_INVALID, "bad record_mode");
  const int64_t tiles = (n_rays + prt::kTileRays - 1) / prt::kTileRays;
  if (tiles > 0x7fffffffLL) return fail(PRT_ERR_LIMIT, "too many rays for one launch");
  const bool record = p->record_mode != PRT_RECORD_NONE;
  prt::TraceArgs a;
  std::memset(&a, 0, sizeof a);
  if (record) {
    if (!rec || !rec->d_stage || !rec->d_run_start || !rec->d_run_count)
      return fail(PRT_ERR_INVALID, "record workspace missing");
    if (rec->n_tiles < tiles) return fail(PRT_ERR_INVALID, "records.n_tiles too small");
    if (rec->capacity < 0) return fail(PRT_ERR_INVALID, "negative capacity");
    a.stage = rec->d_stage;
    a.capacity = rec->capacity;
    a.run_start = reinterpret_cast<long long*>(rec->d_run_start);
    a.run_count = rec->d_run_count;
    a.n_tiles = rec->n_tiles;
  }
  a.blob = scene->d_blob;
  a.blob_bytes = scene->blob_bytes;
  a.generation_limit = p->generation_limit;
  a.record_mode = p->record_mode;
  a.ray_offset = p->ray_offset;
  a.detector_sid = (double)p->detector_sid;
  a.rays = d_rays;
  a.n_rays = n_rays;
  a.stride = ray_stride;
  a.ctr = d_counters;
  if (p->flags & PRT_FLAG_FP32) {
    if (p->flags & PRT_FLAG_DIAGNOSE) return fail(PRT_ERR_UNSUPPORTED, "PRT_FLAG_DIAGNOSE is an FP64 diagnostic");
    if (prt_f32_smem_bytes(scene->blob_bytes, scene->n_leaves, scene->n_components, scene->n_aabb) > 52 * 1024)
      return fail(PRT_ERR_LIMIT, "scene too large for the FP32 fast mode's shared-memory staging");
    cudaError_t e32 = prt_launch_trace_f32(&a, record ? 1 : 0, scene->ordered, scene->generic, scene->n_leaves,
                                           scene->n_components, scene->n_aabb, (cudaStream_t)cuda_stream);
    if (e32 != cudaSuccess) return cuda_fail(e32, "trace kernel launch (fp32)");
    return PRT_OK;
  }
  cudaError_t e = prt_launch_trace(&a, record ? 1 : 0, scene->generic, (p->flags & PRT_FLAG_DIAGNOSE) ? 1 : 0,
                                   (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "trace kernel launch");
  return PRT_OK;
}

int prt_trace_wavefront(prt_scene* scene, const prt_params* p, const double* d_rays, int64_t n_rays,
                        int64_t ray_stride, const prt_wave_workspace* ws, double* d_frame, int64_t frame_stride,
                        int64_t capacity, int64_t* d_gen_offsets, prt_counters* d_counters, void** nearest_events,
                        void* cuda_stream) {
  if (!scene || !p || !d_counters || !ws || !d_gen_offsets) return fail(PRT_ERR_INVALID, "null argument");
  if (n_rays < 0 || (n_rays > 0 && !d_rays)) return fail(PRT_ERR_INVALID, "bad ray buffer");
  if (ray_stride < n_rays) return fail(PRT_ERR_INVALID, "ray_stride < n_rays");
  if (p->generation_limit < 1) return fail(PRT_ERR_INVALID, "generation_limit must be >= 1");
  if (p->generation_limit > 65535) return fail(PRT_ERR_LIMIT, "generation_limit must be <= 65535");
  if (p->record_mode != PRT_RECORD_ALL && p->record_mode != PRT_RECORD_SURFACE)
    return fail(PRT_ERR_INVALID, "the wavefront trace records rows: record_mode must be ALL or SURFACE");
  if (scene->blob_bytes > prt::kMaxSharedBlob) return fail(PRT_ERR_LIMIT, "scene too large for this entry point (prt_trace reads large scenes from global memory)");
  const int64_t tiles = (n_rays + prt_wave_tile() - 1) / prt_wave_tile();
  if (tiles > 0x7fffffffLL) return fail(PRT_ERR_LIMIT, "too many rays for one launch");
  if (ws->n_tiles < tiles) return fail(PRT_ERR_INVALID, "workspace.n_tiles too small");
  if (n_rays > 0 && (!ws->d_state || !ws->d_flag || !ws->d_hit_t || !ws->d_hit_leaf || !ws->d_tile_count ||
                     !ws->d_tile_base))
    return fail(PRT_ERR_INVALID, "workspace buffer missing");
  if (!ws->d_alive) return fail(PRT_ERR_INVALID, "workspace buffer missing");
  if (capacity < 0 || (capacity > 0 && !d_frame) || frame_stride < capacity) return fail(PRT_ERR_INVALID, "bad frame");
  prt::WaveArgs a;
  std::memset(&a, 0, sizeof a);
  a.blob = scene->d_blob;
  a.blob_bytes = scene->blob_bytes;
  a.generation_limit = p->generation_limit;
  a.record_mode = p->record_mode;
  a.ray_offset = p->ray_offset;
  a.detector_sid = (double)p->detector_sid;
  a.rays = d_rays;
  a.n_rays = n_rays;
  a.stride = ray_stride;
  a.st = ws->d_state;
  a.flag = ws->d_flag;
  a.hit_t = ws->d_hit_t;
  a.hit_leaf = ws->d_hit_leaf;
  a.blk_count = ws->d_tile_count;
  a.blk_base = reinterpret_cast<long long*>(ws->d_tile_base);
  a.alive = reinterpret_cast<long long*>(ws->d_alive);
  a.gen_off = reinterpret_cast<long long*>(d_gen_offsets);
  a.frame = d_frame;
  a.frame_stride = frame_stride;
  a.capacity = capacity;
  a.ctr = d_counters;
  cudaError_t e = prt_launch_wavefront(&a, scene->generic, reinterpret_cast<cudaEvent_t*>(nearest_events),
                                       (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "wavefront launch");
  return PRT_OK;
}

int prt_scan_runs(const prt_records* rec, int32_t generation_limit, int64_t* d_gen_offsets, void* cuda_stream) {
  if (!rec || !rec->d_run_count || !rec->d_run_base || !d_gen_offsets)
    return fail(PRT_ERR_INVALID, "null argument");
  if (generation_limit < 1) return fail(PRT_ERR_INVALID, "generation_limit must be >= 1");
  cudaError_t e = prt_launch_scan(rec->d_run_count, reinterpret_cast<long long*>(rec->d_run_base), rec->n_tiles,
                                  generation_limit, reinterpret_cast<long long*>(d_gen_offsets),
                                  (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "scan kernel launch");
  return PRT_OK;
}

int prt_gather_frame(prt_scene* scene, const prt_records* rec, const double* d_rays, int64_t n_rays,
                     int64_t ray_stride, int32_t generation_limit, const int64_t* d_gen_offsets, double* frame,
                     int64_t frame_stride, int64_t frame_capacity, int32_t layout, void* cuda_stream) {
  if (!scene) return fail(PRT_ERR_INVALID, "null scene");
  if (!rec || !rec->d_stage || !rec->d_run_start || !rec->d_run_count || !rec->d_run_base || !d_gen_offsets)
    return fail(PRT_ERR_INVALID, "null argument");
  if (!frame) return fail(PRT_ERR_INVALID, "null frame");
  if (n_rays < 0 || (n_rays > 0 && !d_rays) || ray_stride < n_rays) return fail(PRT_ERR_INVALID, "bad ray buffer");
  const bool f32_records = (layout & PRT_LAYOUT_FP32_RECORDS) != 0;
  layout &= ~PRT_LAYOUT_FP32_RECORDS;
  if (layout != 0 && layout != 1) return fail(PRT_ERR_INVALID, "bad layout");
  if (f32_records && layout != 0) return fail(PRT_ERR_UNSUPPORTED, "FP32 records expand to the column-major frame only");
  if (frame_capacity < 0 || (layout == 0 && frame_stride < frame_capacity)) return fail(PRT_ERR_INVALID, "bad frame");
  if (rec->n_tiles > 0x7fffffffLL) return fail(PRT_ERR_LIMIT, "too many tiles");
  if (rec->n_tiles * (int64_t)prt::kTileRays < n_rays) return fail(PRT_ERR_INVALID, "records.n_tiles too small");
  prt::GatherArgs a;
  std::memset(&a, 0, sizeof a);
  a.blob = scene->d_blob;
  a.rays = d_rays;
  a.n_rays = n_rays;
  a.ray_stride = ray_stride;
  a.stage = rec->d_stage;
  a.capacity = rec->capacity;
  a.run_start = reinterpret_cast<const long long*>(rec->d_run_start);
  a.run_count = rec->d_run_count;
  a.run_base = reinterpret_cast<const long long*>(rec->d_run_base);
  a.n_tiles = rec->n_tiles;  // the run tables are indexed g * n_tiles + tile; tiles without rays have empty runs
  a.gen_offsets = reinterpret_cast<const long long*>(d_gen_offsets);
  a.generation_limit = generation_limit;
  a.frame = frame;
  a.frame_stride = frame_stride;
  a.frame_capacity = frame_capacity;
  cudaError_t e = f32_records ? prt_launch_gather_f32(&a, (cudaStream_t)cuda_stream)
                              : prt_launch_gather(&a, layout, (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "gather kernel launch");
  return PRT_OK;
}

int prt_intersect(prt_scene* scene, int32_t component, const double* d_rays, int64_t n, double* d_hits,
                  int64_t* d_sids, int32_t* slots_out, void* cuda_stream) {
  if (!scene) return fail(PRT_ERR_INVALID, "null scene");
  if (component < 0 || component >= scene->n_components) return fail(PRT_ERR_INVALID, "component out of range");
  const int slots = scene->comp_slots[component];
  if (slots_out) *slots_out = slots;
  if (n == 0) return PRT_OK;
  if (!d_rays || !d_hits || !d_sids) return fail(PRT_ERR_INVALID, "null buffer");
  cudaError_t e = prt_launch_intersect(scene->d_blob, scene->blob_bytes, component, d_rays, n, d_hits,
                                       reinterpret_cast<long long*>(d_sids), slots, (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "intersect kernel launch");
  return PRT_OK;
}

int prt_nearest_hit(prt_scene* scene, const double* d_rays, int64_t n, double* d_t, int64_t* d_sid, double* d_normals,
                    void* cuda_stream) {
  if (!scene) return fail(PRT_ERR_INVALID, "null scene");
  if (n < 0) return fail(PRT_ERR_INVALID, "negative ray count");
  if (n == 0) return PRT_OK;
  if (!d_rays || !d_t || !d_sid) return fail(PRT_ERR_INVALID, "null buffer");
  cudaError_t e = prt_launch_nearest(scene->d_blob, scene->blob_bytes, d_rays, n, d_t,
                                     reinterpret_cast<long long*>(d_sid), d_normals, (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "nearest kernel launch");
  return PRT_OK;
}

int prt_render_hit(prt_scene* scene, const double* d_rays, int64_t n, double* d_t, int64_t* d_sid, double* d_normals,
                   void* cuda_stream) {
  if (!scene) return fail(PRT_ERR_INVALID, "null scene");
  if (n < 0) return fail(PRT_ERR_INVALID, "negative ray count");
  if (n == 0) return PRT_OK;
  if (!d_rays || !d_t || !d_sid) return fail(PRT_ERR_INVALID, "null buffer");
  cudaError_t e = prt_launch_render_hit(scene->d_blob, scene->blob_bytes, d_rays, n, d_t,
                                        reinterpret_cast<long long*>(d_sid), d_normals, (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "render hit kernel launch");
  return PRT_OK;
}

int prt_generate_source(const prt_source_desc* src, double* d_rays, int64_t n_rays, int64_t ray_stride,
                        int64_t first_index, void* cuda_stream) {
  if (!src || (!d_rays && n_rays > 0)) return fail(PRT_ERR_INVALID, "null argument");
  const bool synthetic = src->kind >= 1 && src->kind <= 3;
  const bool reference = src->kind >= 10 && src->kind <= 14;
  if (!synthetic && !reference) return fail(PRT_ERR_INVALID, "unknown source kind");
  if (reference && (first_index < 0 || (double)(first_index + n_rays) > src->p[2]))
    return fail(PRT_ERR_INVALID, "window [first_index, first_index + n_rays) exceeds the source's ray count p[2]");
  if (ray_stride < n_rays) return fail(PRT_ERR_INVALID, "ray_stride < n_rays");
  cudaError_t e = prt_launch_source(src, d_rays, n_rays, ray_stride, first_index, (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "source kernel launch");
  return PRT_OK;
}

int prt_frame_pack(const double* d_frame, int64_t rows, int64_t frame_stride, const double* d_rays, int64_t n_rays,
                   int64_t ray_stride, const int64_t* d_gen_offsets, int32_t generation_limit, uint64_t* d_packed,
                   uint64_t* d_bad, void* cuda_stream) {
  if (rows < 0 || (rows > 0 && (!d_frame || !d_packed)) || frame_stride < rows) return fail(PRT_ERR_INVALID, "bad frame");
  if (!d_bad) return fail(PRT_ERR_INVALID, "d_bad is NULL");
  if (rows > 0 && (n_rays < 1 || !d_rays || ray_stride < n_rays)) return fail(PRT_ERR_INVALID, "bad ray buffer");
  if (rows > 0 && (!d_gen_offsets || generation_limit < 1)) return fail(PRT_ERR_INVALID, "generation offsets missing");
  cudaError_t e = prt_launch_frame_pack(d_frame, rows, frame_stride, d_rays, n_rays, ray_stride,
                                        reinterpret_cast<const long long*>(d_gen_offsets), generation_limit,
                                        reinterpret_cast<unsigned long long*>(d_packed),
                                        reinterpret_cast<unsigned long long*>(d_bad), (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "frame pack launch");
  return PRT_OK;
}

int prt_host_expand_frame(const uint64_t* h_packed, int64_t rows, const int64_t* h_gen_offsets,
                          int32_t generation_limit, const double* h_ray_generation, const double* h_ray_intensity,
                          const double* h_ray_wavelength, const double* h_ray_id, double* h_frame,
                          int64_t frame_stride, int32_t threads) {
  if (rows < 0 || generation_limit < 1) return fail(PRT_ERR_INVALID, "bad arguments");
  if (rows == 0) return PRT_OK;
  if (!h_packed || !h_gen_offsets || !h_ray_generation || !h_ray_intensity || !h_ray_wavelength || !h_ray_id ||
      !h_frame || frame_stride < rows)
    return fail(PRT_ERR_INVALID, "null or short buffer");
  if (h_gen_offsets[0] != 0 || h_gen_offsets[generation_limit] != rows)
    return fail(PRT_ERR_INVALID, "generation offsets do not cover the frame");
  int t = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  t = std::max(1, std::min<int>(t, (int)std::min<int64_t>(64, (rows + 65535) / 65536)));
  std::vector<std::thread> pool;
  const int64_t chunk = (rows + t - 1) / t;
  for (int k = 0; k < t; ++k) {
    const int64_t r0 = (int64_t)k * chunk, r1 = std::min<int64_t>(rows, r0 + chunk);
    if (r0 >= r1) break;
    pool.emplace_back(prt_host_expand_rows, h_packed, r0, r1, h_gen_offsets, generation_limit, h_ray_generation,
                      h_ray_intensity, h_ray_wavelength, h_ray_id, h_frame, frame_stride);
  }
  for (auto& th : pool) th.join();
  return PRT_OK;
}

static bool bad_select(int32_t select) { return select < PRT_SELECT_ALL || select > PRT_SELECT_GENERATION; }

int prt_spot_moments(const double* d_frame, int64_t rows, int64_t frame_stride, int32_t select, double value,
                     int64_t rays_per_group, int32_t n_groups, const double* d_center, double* d_out,
                     int32_t blocks, void* cuda_stream) {
  if (rows < 0 || (rows > 0 && !d_frame) || frame_stride < rows) return fail(PRT_ERR_INVALID, "bad frame");
  if (!d_out) return fail(PRT_ERR_INVALID, "d_out is NULL");
  if (bad_select(select)) return fail(PRT_ERR_INVALID, "unknown row selection");
  if (n_groups < 1 || n_groups > PRT_SPOT_MAX_GROUPS)
    return fail(PRT_ERR_INVALID, "n_groups must be 1.." + std::to_string(PRT_SPOT_MAX_GROUPS));
  if (rays_per_group < 1) return fail(PRT_ERR_INVALID, "rays_per_group must be >= 1");
  cudaError_t e = prt_launch_spot_moments(d_frame, rows, frame_stride, select, value, rays_per_group, n_groups,
                                          d_center, d_out, blocks, (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "spot moments launch");
  return PRT_OK;
}

int prt_spot_centers(const double* d_sums, const double* d_center_in, int32_t n_groups, double* d_center_out,
                     void* cuda_stream) {
  if (!d_sums || !d_center_out) return fail(PRT_ERR_INVALID, "NULL buffer");
  if (n_groups < 1 || n_groups > PRT_SPOT_MAX_GROUPS) return fail(PRT_ERR_INVALID, "bad n_groups");
  cudaError_t e = prt_launch_spot_centers(d_sums, d_center_in, n_groups, d_center_out, (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "spot centers launch");
  return PRT_OK;
}

int64_t prt_axis_table_blocks(int64_t rows) { return rows <= 0 ? 0 : (rows + 1023) / 1024; }

int prt_axis_table(const double* d_frame, int64_t rows, int64_t frame_stride, int32_t select, double value,
                   int64_t first_id, int64_t gen0_rows, int32_t* d_block_count, int64_t* d_block_base,
                   int64_t* d_total, double* d_table, int64_t table_stride, int64_t table_capacity,
                   void* cuda_stream) {
  if (rows < 0 || (rows > 0 && !d_frame) || frame_stride < rows) return fail(PRT_ERR_INVALID, "bad frame");
  if (bad_select(select)) return fail(PRT_ERR_INVALID, "unknown row selection");
  if (!d_total) return fail(PRT_ERR_INVALID, "d_total is NULL");
  if (rows > 0 && (!d_block_count || !d_block_base)) return fail(PRT_ERR_INVALID, "workspace is NULL");
  if (table_capacity < 0 || (table_capacity > 0 && !d_table) || table_stride < table_capacity)
    return fail(PRT_ERR_INVALID, "bad table");
  if (gen0_rows < 0 || gen0_rows > rows) return fail(PRT_ERR_INVALID, "gen0_rows out of range");
  if (prt_axis_table_blocks(rows) > 0x7fffffffLL) return fail(PRT_ERR_INVALID, "frame too long for one call");
  cudaError_t e = prt_launch_axis_table(d_frame, rows, frame_stride, select, value, first_id, gen0_rows,
                                        d_block_count, (long long*)d_block_base, (long long*)d_total, d_table,
                                        table_stride, table_capacity, (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "axis table launch");
  return PRT_OK;
}

int prt_fp64_probe(double* d_scratch, int32_t blocks, int32_t iters, void* cuda_stream) {
  if (!d_scratch || blocks < 1 || iters < 1) return fail(PRT_ERR_INVALID, "bad probe arguments");
  cudaError_t e = prt_launch_fp64_probe(d_scratch, blocks, iters, (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "fp64 probe launch");
  return PRT_OK;
}

}  // extern "C"
