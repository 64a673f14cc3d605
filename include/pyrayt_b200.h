/*
 * pyrayt_b200 -- C ABI of the B200-native ray-propagation hot path.
 *
 * This is the drop-in boundary underneath PyRayT's Python API.  The reference
 * (rfrazier716/PyRayT v0.3.1) is pure Python/NumPy and has no FFI of its own;
 * each entry point below names the reference interface it replaces so a
 * maintainer can bind it with ctypes:
 *
 *   prt_scene_create     <- the live scene RayTracer holds: RayTracer.__init__
 *                           components + _surface_lut (pyrayt/_pyrayt.py:211-260),
 *                           TracerSurface matrices/primitives/materials
 *                           (tinygfx/g3d/world_objects.py:338-423),
 *                           CSGSurface trees + _aobb (tinygfx/g3d/csg.py:64-116)
 *   prt_trace            <- RayTracer._st_propagate + _st_interact generation loop
 *                           (pyrayt/_pyrayt.py:370-452) and everything under it:
 *                           TracerSurface.intersect / get_world_normals
 *                           (world_objects.py:360-418), CSGSurface.intersect +
 *                           array_csg (csg.py:13-160), primitives' intersect/normal
 *                           (primitives.py:241-741), reflect/refract
 *                           (operations.py:86-162), materials' trace/index_at
 *                           (pyrayt/materials.py:47-145) and the per-generation
 *                           record of _RayTraceDataframe.insert (_pyrayt.py:168-186)
 *   prt_scan_runs +
 *   prt_gather_frame     <- the (generation, id) row order produced by
 *                           DataFrame.append per generation (_pyrayt.py:186,:428-435)
 *   prt_intersect        <- component.intersect(rays) -> (hits(m,N), surface_ids(m,N))
 *                           (world_objects.py:360-383, csg.py:118-160)
 *   prt_generate_source  <- Source.generate_rays (pyrayt/components.py:481-496) for the
 *                           seeded synthetic benchmark sources
 *
 * Conventions: plain pointers and sizes, no C++ or torch types; every function
 * returns 0 on success or a negative prt_status and never throws; the message
 * of the last failure on the calling thread is prt_last_error().  All d_*
 * pointers are device pointers owned by the caller (the Python host allocates
 * them as torch tensors); the library performs no hidden allocation and no
 * stream synchronisation inside prt_trace / prt_scan_runs / prt_gather_frame /
 * prt_intersect / prt_generate_source: they enqueue work on `cuda_stream`
 * (a cudaStream_t passed as void*) and return.
 *
 * All arithmetic is IEEE float64 without FMA contraction (the reference computes in float64 with NumPy),
 * except under PRT_FLAG_FP32, the optional single-precision fast mode with a stated tolerance.
 */
#ifndef PYRAYT_B200_H
#define PYRAYT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRT_ABI_VERSION 3

/* rows of the reference RaySet, (13, N) float64 row-major (pyrayt/_pyrayt.py:13-144) */
#define PRT_RAY_ROWS 13
/* columns of the results frame (pyrayt/_pyrayt.py:15,:154-165):
 * generation,intensity,wavelength,index,id,surface,x0,y0,z0,x1,y1,z1,x_tilt,y_tilt,z_tilt */
#define PRT_FRAME_COLS 15
#define PRT_STAGE_COLS 9  /* doubles per staged record, see prt_records */
/* hit slots one component may produce = 2 x leaves of the component */
#define PRT_MAX_SLOTS 32
/* largest scene a handle accepts.  Scenes whose encoded form fits a thread block's share of shared memory
 * (about 100 KB: a few hundred leaves) are staged there by every block; larger ones (lenslet arrays, ...) are
 * read in place through L1 / L2 by prt_trace, prt_intersect, prt_nearest_hit and prt_render_hit
 * (prt_trace_wavefront and the FP32 fast mode return PRT_ERR_LIMIT for them) */
#define PRT_MAX_LEAVES 4096
#define PRT_MAX_NODES 8192

typedef enum prt_status {
  PRT_OK = 0,
  PRT_ERR_INVALID = -1,   /* bad argument / malformed scene            */
  PRT_ERR_CUDA = -2,      /* CUDA runtime failure (see prt_last_error) */
  PRT_ERR_LIMIT = -3,     /* scene exceeds PRT_MAX_*                   */
  PRT_ERR_UNSUPPORTED = -4
} prt_status;

/* primitive type codes: tinygfx/g3d/primitives.py Sphere:220 Paraboloid:299 Plane:422 Cube:501 Cylinder:621 */
typedef enum prt_prim {
  PRT_SPHERE = 1,     /* param[0]=radius                                             */
  PRT_PARABOLOID = 2, /* param[0]=focus, [1]=height                                  */
  PRT_PLANE = 3,      /* param[0]=width(x), [1]=length(y)                            */
  PRT_CUBE = 4,       /* param = x_lo,x_hi,y_lo,y_hi,z_lo,z_hi (Cube.axis_spans)     */
  PRT_CYLINDER = 5    /* param[0]=radius, [1]=h_min, [2]=h_max, [3]=capped (0/1)     */
} prt_prim;

/* postfix node kinds; the CSG values equal tinygfx.g3d.csg.Operation (csg.py:7-10) */
typedef enum prt_node_kind { PRT_LEAF = 0, PRT_UNION = 1, PRT_INTERSECT = 2, PRT_DIFFERENCE = 3 } prt_node_kind;

/* material kinds: pyrayt/materials.py absorber:47 mirror:58 BasicRefractor:102 SellmeierRefractor:121 */
typedef enum prt_material {
  PRT_MAT_ABSORBER = 0,
  PRT_MAT_MIRROR = 1,
  PRT_MAT_GLASS_CONST = 2,     /* matp[0] = refractive index                      */
  PRT_MAT_GLASS_SELLMEIER = 3, /* matp = b1,b2,b3,c1,c2,c3                        */
  PRT_MAT_UNTRACEABLE = 4      /* no .trace(): a hit raises in the reference (Q9) */
} prt_material;

/*
 * Host-side flat scene, SoA.  Components are listed in RayTracer order; each
 * component is a postfix CSG program over nodes [comp_node_begin[c],
 * comp_node_begin[c+1]).  A bare TracerSurface component is a single PRT_LEAF.
 */
typedef struct prt_scene_desc {
  int32_t n_components;
  int32_t n_nodes;
  int32_t n_leaves;
  int32_t reserved;
  const int32_t* comp_node_begin; /* [n_components+1]                                          */
  const int32_t* node_kind;       /* [n_nodes] prt_node_kind                                   */
  const int32_t* node_leaf;       /* [n_nodes] leaf index for PRT_LEAF nodes, else -1          */
  const double* node_aabb;        /* [n_nodes*6] CSGSurface._aobb.axis_spans (x_lo,x_hi,y_lo,..)*/
  const int32_t* leaf_type;       /* [n_leaves] prt_prim                                       */
  const double* leaf_obj;         /* [n_leaves*16] row-major 4x4 world->object matrix          */
  const double* leaf_param;       /* [n_leaves*6]                                              */
  const double* leaf_nscale;      /* [n_leaves] Intersectable._normal_scale (+1/-1)            */
  const int64_t* leaf_sid;        /* [n_leaves] CountedObject id written to the `surface` col  */
  const int32_t* leaf_mat;        /* [n_leaves] prt_material                                   */
  const double* leaf_matp;        /* [n_leaves*6]                                              */
} prt_scene_desc;

typedef struct prt_scene prt_scene; /* opaque, owns a small device copy of the flattened scene */

typedef enum prt_record_mode {
  PRT_RECORD_ALL = 0,     /* every segment (the reference frame)               */
  PRT_RECORD_SURFACE = 1, /* only rows whose surface id == params.detector_sid */
  PRT_RECORD_NONE = 2     /* counters only                                     */
} prt_record_mode;

typedef struct prt_params {
  int32_t generation_limit; /* RayTracer generation_limit (pyrayt/_pyrayt.py:212,:444)  */
  int32_t record_mode;      /* prt_record_mode                                          */
  int32_t flags;            /* PRT_FLAG_* bits, 0 = the plain FP64 trace                */
  int32_t reserved;
  double ray_offset;        /* RayTracer.ray_offset_value = 1e-6 (pyrayt/_pyrayt.py:190)*/
  int64_t detector_sid;     /* for PRT_RECORD_SURFACE                                   */
} prt_params;

/*
 * prt_params.flags
 * PRT_FLAG_DIAGNOSE: besides tracing, count the rays the north star excludes from the bit-exact id
 * contract -- rays "within 1e-9 of grazing or CSG seams".  Operational definition: in some generation the
 * nearest-hit search (_st_propagate, pyrayt/_pyrayt.py:370-392) answers differently when the ray's origin is
 * displaced by 1e-9 x max(1, |origin|_inf) perpendicular to its direction (four displacements: +-e1, +-e2,
 * e1 = unit(v x axis of the smallest |v_k|), e2 = (v x e1) / |v|).  A hit that turns into a miss or a
 * miss that turns into a hit counts the ray in grazing_rays, a hit on a different surface in seam_rays.
 * Five nearest-hit searches per generation instead of one: a diagnostic, not the timed path.
 */
#define PRT_FLAG_DIAGNOSE 1
/*
 * PRT_FLAG_FP32: the optional fast mode of the north star.  The whole generation loop -- transforms,
 * intersections, CSG merge, nearest hit, normals, Snell / Sellmeier -- runs in single precision with FMA
 * contraction and fast division / square root; the frame keeps its fifteen float64 columns.  Contract: the
 * frame agrees with the FP64 frame to 1e-5 of the scene scale (positions; 1e-5 absolute for the unit tilt
 * and the refractive index), surface and generation ids agree except for rays passing within that distance
 * of an edge or seam.  A ray cannot leave a surface by the reference's 1e-6 offset in single precision
 * (ulp(100) = 7.6e-6): instead, roots of the leaf just hit that lie within 2e-4 x max(1, |origin|) of the
 * origin are taken as the crossing just made (so features of the scene -- thicknesses, gaps between surfaces
 * of one leaf -- must be well above 2e-4 length units, as they must be well above the reference's own 1e-6
 * offset in FP64).  Every scene the FP64 path stages in shared memory is accepted
 * (arbitrary CSG trees run a single-precision interpreter); scenes too large for that return PRT_ERR_LIMIT.
 * Staged records are 40 bytes (5 of the PRT_STAGE_COLS columns): pass PRT_LAYOUT_FP32_RECORDS in
 * prt_gather_frame's layout.
 * Not available in prt_trace_wavefront.
 */
#define PRT_FLAG_FP32 2
#define PRT_LAYOUT_FP32_RECORDS 0x100

/* device-resident counters, zeroed by the caller before prt_trace */
typedef struct prt_counters {
  uint64_t rays;               /* rays handed in                                                     */
  uint64_t generations;        /* sum over rays of generations entered; x n_leaves = ray-surface tests */
  uint64_t segments;           /* rows the trace produced (whether or not capacity allowed the write)  */
  uint64_t rows_reserved;      /* append cursor: staging rows reserved                                 */
  uint64_t rows_dropped;       /* rows not written because staging capacity was exhausted              */
  uint64_t tie_rays;           /* rays whose CSG merge compared equal finite keys of different leaves; the
                                  closed-form merges of left-deep trees look for them under PRT_FLAG_DIAGNOSE only */
  uint64_t untraceable_hits;   /* nearest hit landed on a PRT_MAT_UNTRACEABLE surface                  */
  uint64_t bad_w;              /* rays whose homogeneous w rows are not (1, 0)                         */
  uint64_t nan_rays;           /* rays terminated because their direction became NaN                   */
  uint64_t limit_rays;         /* rays stopped by generation_limit                                     */
  uint64_t absorber_segments;  /* segments that ended on an absorber                                   */
  uint64_t mirror_segments;    /* segments that ended on a mirror (the rest are glass)                 */
  uint64_t grazing_rays;       /* PRT_FLAG_DIAGNOSE: rays with a generation whose hit <-> miss flips under a 1e-9 shift */
  uint64_t seam_rays;          /* PRT_FLAG_DIAGNOSE: rays with a generation whose hit surface changes under a 1e-9 shift */
  uint64_t reserved[2];
} prt_counters;

/*
 * Caller-owned record workspace.  The trace kernel processes rays in tiles of
 * prt_tile_rays() consecutive rays; at every generation a tile appends its
 * surviving rays' rows (in ray order) as one contiguous run to the
 * column-major staging buffer and notes (start, count) in the run table.
 */
typedef struct prt_records {
  double* d_stage;      /* [PRT_STAGE_COLS * capacity] staged records, column c row r at d_stage[c*capacity + r]:
                           start position (3), direction as traced (3), hit distance, refractive index before the
                           interaction, (slot of the ray in its tile) << 32 | leaf hit -- what only the trace knows
                           about a row; prt_gather_frame expands it into the 15 frame columns               */
  int64_t capacity;     /* rows                                                                           */
  int64_t* d_run_start; /* [generation_limit * n_tiles], index g*n_tiles + tile                           */
  int32_t* d_run_count; /* [generation_limit * n_tiles]                                                   */
  int64_t* d_run_base;  /* [generation_limit * n_tiles] filled by prt_scan_runs: final frame row of run   */
  int64_t n_tiles;      /* >= ceil(n_rays / prt_tile_rays())                                              */
} prt_records;

int prt_abi_version(void);
const char* prt_last_error(void);
int prt_tile_rays(void);

int prt_scene_create(const prt_scene_desc* host_scene, int device, prt_scene** out);
void prt_scene_destroy(prt_scene* scene);
int prt_scene_n_leaves(const prt_scene* scene);

/*
 * Trace n_rays rays through every generation.  d_rays is the reference RaySet
 * layout: row k of ray i at d_rays[k*ray_stride + i], rows = x,y,z,w(=1),
 * dx,dy,dz,w(=0),generation,intensity,wavelength,index,id.
 * d_run_count must be zeroed by the caller when record_mode != PRT_RECORD_NONE.
 */
int prt_trace(prt_scene* scene, const prt_params* params, const double* d_rays, int64_t n_rays,
              int64_t ray_stride, const prt_records* records, prt_counters* d_counters,
              void* cuda_stream);

/*
 * The same trace for large ray sets, one launch per generation ("wavefront"): launch g finishes
 * generation g-1 (interaction, row, new state) and starts generation g (nearest hit, per-tile count of
 * the rows it will produce); a scan between two launches turns the counts into positions, so every
 * row is written straight to its final (generation, id) place in d_frame (column-major, column c row
 * r at d_frame[c*frame_stride + r], at most `capacity` rows; further rows are counted in
 * d_counters->rows_dropped, and d_counters->rows_reserved receives the exact row count).  No staging
 * buffer, no ordering pass; the frame is bit-identical to prt_trace + prt_scan_runs + prt_gather_frame.
 * All 2 x generation_limit + 1 launches are enqueued by this call; launches for generations no ray
 * reaches return immediately on the device.  d_gen_offsets[g] (g = 0 .. generation_limit) receives
 * the first frame row of generation g, d_gen_offsets[generation_limit] the row count.  record_mode
 * PRT_RECORD_NONE is not supported here (use prt_trace).
 * step_events (optional, may be NULL): 2 x (generation_limit + 1) cudaEvent_t handles recorded
 * before and after each launch of the step kernel, for timing that kernel alone.
 */
typedef struct prt_wave_workspace {
  double* d_state;       /* [7 * n_rays] position, direction, refractive index of every ray        */
  int32_t* d_flag;       /* [n_rays]                                                               */
  double* d_hit_t;       /* [n_rays]                                                               */
  int32_t* d_hit_leaf;   /* [n_rays]                                                               */
  int32_t* d_tile_count; /* [2 * n_tiles] rows each tile writes, by generation parity              */
  int64_t* d_tile_base;  /* [2 * n_tiles]                                                          */
  int64_t* d_alive;      /* [generation_limit + 1] live rays entering each generation              */
  int64_t n_tiles;       /* >= ceil(n_rays / prt_wave_tile())                                      */
} prt_wave_workspace;

int prt_wave_tile(void);
int prt_trace_wavefront(prt_scene* scene, const prt_params* params, const double* d_rays, int64_t n_rays,
                        int64_t ray_stride, const prt_wave_workspace* ws, double* d_frame,
                        int64_t frame_stride, int64_t capacity, int64_t* d_gen_offsets,
                        prt_counters* d_counters, void** step_events, void* cuda_stream);

/*
 * Turn the run table into final frame positions: d_gen_offsets[g] (g = 0 ..
 * generation_limit) receives the first frame row of generation g and
 * d_gen_offsets[generation_limit] the total row count.  The caller may rewrite
 * d_gen_offsets before prt_gather_frame (multi-GPU: shift every generation by the
 * rows lower ranks contribute to it, from the all-gathered per-rank counts).
 */
int prt_scan_runs(const prt_records* records, int32_t generation_limit, int64_t* d_gen_offsets,
                  void* cuda_stream);

/*
 * Expand the staged records into the final frame in (generation, id) order: the row
 * _RayTraceDataframe.insert builds (pyrayt/_pyrayt.py:168-186) -- generation / intensity / wavelength / id
 * from the RaySet the trace was given (same d_rays, n_rays, ray_stride), surface id from the scene, end
 * point = start + direction * distance, unit tilt -- computed with the trace kernel's own functions, so the
 * frame is bit-identical to writing the columns in the trace.
 * layout 0: column-major, column c row r at frame[c*frame_stride + r]  (what pandas holds)
 * layout 1: row-major,    row r column c at frame[r*PRT_FRAME_COLS + c]
 * Rows at or beyond frame_capacity are not written (the caller sizes the frame from a hint, reads
 * d_gen_offsets[generation_limit] afterwards and repeats with a larger frame if it was short).
 * `frame` may be any device-accessible pointer, including pinned host memory.
 */
int prt_gather_frame(prt_scene* scene, const prt_records* records, const double* d_rays, int64_t n_rays,
                     int64_t ray_stride, int32_t generation_limit, const int64_t* d_gen_offsets, double* frame,
                     int64_t frame_stride, int64_t frame_capacity, int32_t layout, void* cuda_stream);

/*
 * Lean device -> host transfer of a frame.  Five of the fifteen columns can be rebuilt on the host from
 * the input rays: generation (from the row position), intensity, wavelength and id (copies of the ray's
 * input values) and surface (a small integer).  prt_frame_pack writes one word per row,
 * (index of the ray in d_rays) << 24 | (surface id + 1), after checking row by row that the frame holds
 * exactly what the host will rebuild (ids consecutive from d_rays' first id, intensity / wavelength
 * bit-equal to the ray's, surface id an integer in [-1, 2^24 - 2], generation = the ray's own value in the
 * rows before d_gen_offsets[1] and g in [d_gen_offsets[g], d_gen_offsets[g+1]); d_gen_offsets holds
 * generation_limit + 1 entries as written by prt_scan_runs); *d_bad receives the number of rows
 * that failed -- if it is not zero the caller must copy all fifteen columns instead.
 * prt_host_expand_frame (host code, `threads` worker threads) then fills columns 0, 1, 2, 4, 5 of a host
 * frame from the packed words, the per-generation row offsets and host copies of rows 8, 9, 10 and 12 of the RaySet
 * (generation, intensity, wavelength, id; each n_rays long), while the
 * other ten columns are copied as usual: 88 instead of 120 bytes per row cross the bus.
 */
int prt_frame_pack(const double* d_frame, int64_t rows, int64_t frame_stride, const double* d_rays, int64_t n_rays,
                   int64_t ray_stride, const int64_t* d_gen_offsets, int32_t generation_limit, uint64_t* d_packed,
                   uint64_t* d_bad, void* cuda_stream);
int prt_host_expand_frame(const uint64_t* h_packed, int64_t rows, const int64_t* h_gen_offsets,
                          int32_t generation_limit, const double* h_ray_generation, const double* h_ray_intensity,
                          const double* h_ray_wavelength, const double* h_ray_id, double* h_frame,
                          int64_t frame_stride, int32_t threads);

/*
 * component.intersect(rays): d_rays is (2,4,N) like the reference (row k of ray i
 * at d_rays[k*n + i]); writes hits (m,N) ascending with +inf padding and the
 * surface ids (m,N) exactly as the reference returns them -- the ids carried by
 * the +inf slots included (a missed leaf keeps its id, a CSG node whose box was
 * missed reports -1; tinygfx/g3d/csg.py:126-160).
 * m = 2 x leaves of the component (returned through *slots_out when non-NULL).
 */
int prt_intersect(prt_scene* scene, int32_t component, const double* d_rays, int64_t n,
                  double* d_hits, int64_t* d_sids, int32_t* slots_out, void* cuda_stream);

/*
 * Nearest positive hit over all components = RayTracer._st_propagate alone (pyrayt/_pyrayt.py:370-392);
 * the same loop the reference's renderers run per pixel (tinygfx/g3d/renderers.py:72-94,:188-210).
 * d_rays (2,4,N) as for prt_intersect.  d_t[i] = distance or +inf, d_sid[i] = surface id or -1;
 * d_normals (optional, may be NULL): (3,N) world normal at the hit (get_world_normals,
 * world_objects.py:401-418), NaN for a miss.
 */
int prt_nearest_hit(prt_scene* scene, const double* d_rays, int64_t n, double* d_t, int64_t* d_sid,
                    double* d_normals, void* cuda_stream);

/*
 * EdgeRender._st_propagate / ShadedRenderer._st_propagate (tinygfx/g3d/renderers.py:72-94,:188-210) exactly:
 * the loop of prt_nearest_hit, except that -- like the renderers -- distance and surface are read from the
 * unfiltered hit array at the argmin of the filtered one, so a pixel whose component hits are all behind
 * the camera reports the component's first (negative) hit instead of a miss.  Same arguments.
 */
int prt_render_hit(prt_scene* scene, const double* d_rays, int64_t n, double* d_t, int64_t* d_sid,
                   double* d_normals, void* cuda_stream);

/*
 * Re-encode a scene into an existing handle (same device buffer when the encoded size is unchanged):
 * for optimisation loops that move components or change radii between thousands of small traces
 * (examples/lens_design.ipynb cells 28-33).  The copy is enqueued on cuda_stream.
 */
int prt_scene_update(prt_scene* scene, const prt_scene_desc* host_scene, void* cuda_stream);

/* seeded synthetic sources of SURVEY.md 8(d); see pyrayt_b200/sources.py for the exact law */
typedef struct prt_source_desc {
  int32_t kind;        /* 1 = disk/field/wavelength fan (config 4), 2 = solid-angle cone (config 2),
                          3 = Lambertian cone (config 5);
                          the reference's deterministic sources (pyrayt/components.py:511-613):
                          10 = LineOfRays, 11 = CircleOfRays, 12 = ConeOfRays, 13 = WedgeOfRays, 14 = Lamp
                          (:616-654: the same Lambertian law driven by counter-based uniforms of `seed` and the
                          ray id instead of NumPy's global stream; origin[0] = width, origin[1] = length) with
                          p[0] = spacing | diameter | cone angle | wedge angle | max angle [rad], p[1] = wavelength,
                          p[2] = ray count of the source, p[3] = id of its first ray,
                          p[4..15] = rows 0..2 of the source's 4x4 world matrix; a call writes the window
                          [first_index, first_index + n_rays) of the source's p[2] rays (a rank's share of
                          a sharded source; first_index = 0, n_rays = p[2] for the whole source)         */
  int32_t reserved;
  uint64_t seed;
  double origin[3];
  double p[16];        /* kind-specific parameters                                                    */
} prt_source_desc;

int prt_generate_source(const prt_source_desc* src, double* d_rays, int64_t n_rays, int64_t ray_stride,
                        int64_t first_index, void* cuda_stream);

/*
 * On-device read-out of a results frame: the pandas reductions of
 * examples/lens_design.ipynb cells 11-20 run where the frame already is.  d_frame is the
 * column-major frame prt_gather_frame wrote (column c of row r at d_frame[c*frame_stride + r]).
 *
 * Row selection (the notebook's two filters): PRT_SELECT_SURFACE keeps rows with
 * surface == value (`results['surface'] == imager.get_id()`, cells 11/19/38), PRT_SELECT_GENERATION
 * rows with generation == value (`results['generation'] == max`, cells 12/15/20).
 */
enum prt_select { PRT_SELECT_ALL = 0, PRT_SELECT_SURFACE = 1, PRT_SELECT_GENERATION = 2 };
#define PRT_SPOT_COLS 16
#define PRT_SPOT_CENTER_COLS 4
#define PRT_SPOT_MAX_GROUPS 256

/*
 * Per-group moments of the selected rows; group = floor(id / rays_per_group), the reference's
 * source_id (RayTracer.calculate_source_ids, pyrayt/_pyrayt.py:316-327); rows whose group is
 * >= n_groups are ignored.  d_center (optional, may be NULL = all zero): per group
 * {cy, cz, cf, ct}; sums are taken of (value - centre) so a second call with the centroids of the
 * first gives cancellation-free second moments.  d_out: n_groups x PRT_SPOT_COLS doubles,
 *   0 n            1 S(y1-cy)      2 S(z1-cz)      3 S(y1-cy)^2    4 S(z1-cz)^2   5 S(y1-cy)(z1-cz)
 *   6 min y1       7 max y1        8 min z1        9 max z1
 *  10 n_f (rows with a finite focus)   11 S(f-cf)   12 S(f-cf)^2,  f = -x_tilt*y0/y_tilt + x0 (cell 12)
 *  13 S(y_tilt-ct) 14 S(y_tilt-ct)^2  15 S(sin(y_tilt)-ct)^2 (the coma metric of cell 20, ct = sin(angle))
 * Sums are accumulated with floating-point atomics: the order, hence the last bits, may differ
 * between calls.  `blocks` > 0 caps the grid (default: one block per 8192 rows).
 */
int prt_spot_moments(const double* d_frame, int64_t rows, int64_t frame_stride, int32_t select, double value,
                     int64_t rays_per_group, int32_t n_groups, const double* d_center, double* d_out,
                     int32_t blocks, void* cuda_stream);

/* centres for a second, centred prt_spot_moments call: d_center_out = d_center_in (or 0) + mean residual of d_sums */
int prt_spot_centers(const double* d_sums, const double* d_center_in, int32_t n_groups, double* d_center_out,
                     void* cuda_stream);

/*
 * The focus table of cells 12 and 15: one column per selected row, in frame order, rows
 * {id, radius, focus, wavelength}: radius = y0 of the ray's generation-0 row, found by bisection on
 * the id column of the first gen0_rows rows (generation-0 rows are in id order; rays that miss in
 * generation 0 have no row, so ids there need not be consecutive; first_id, the lowest ray id, only
 * seeds the search; NaN if the ray has no generation-0 row),
 * focus = -x_tilt*y0/y_tilt + x0.  Row k of output column j at d_table[k*table_stride + j]; at most
 * table_capacity columns are written.  Workspace: d_block_count / d_block_base hold
 * prt_axis_table_blocks(rows) entries; d_total[1] receives the number of selected rows (d_total[0] = 0).
 */
int64_t prt_axis_table_blocks(int64_t rows);
int prt_axis_table(const double* d_frame, int64_t rows, int64_t frame_stride, int32_t select, double value,
                   int64_t first_id, int64_t gen0_rows, int32_t* d_block_count, int64_t* d_block_base,
                   int64_t* d_total, double* d_table, int64_t table_stride, int64_t table_capacity,
                   void* cuda_stream);

/*
 * Measurement aid (no reference counterpart): launches `blocks` x 256 threads that each run
 * `iters` x 8 independent double-precision FMAs, so the caller can time the FP64 pipe's
 * FMA rate with CUDA events (flops = blocks*256*iters*16).  d_scratch: >= 1 double.
 */
int prt_fp64_probe(double* d_scratch, int32_t blocks, int32_t iters, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* PYRAYT_B200_H */
