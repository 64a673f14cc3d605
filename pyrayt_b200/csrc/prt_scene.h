// Device-side scene encoding shared by the host ABI layer and the kernels.
//
// The caller hands prt_scene_create() postfix CSG programs (include/pyrayt_b200.h).
// The encoder (prt_encode.h) turns each component into a Comp record: bare leaves and
// left-deep trees of two or three leaves (everything the reference's factories build)
// are evaluated in registers straight from that record; any other tree runs a *preorder*
// op program with skip targets, so that the bounding-box cull of CSGSurface.intersect
// (tinygfx/g3d/csg.py:126-133) can skip a whole subtree per ray.  The encoded scene is
// one contiguous blob that every thread block copies to shared memory once.
#pragma once
#include <stdint.h>

#include "../../include/pyrayt_b200.h"

namespace prt {

#ifndef PRT_TILE
#define PRT_TILE 256
#endif
constexpr int kTileRays = PRT_TILE;  // rays per tile == threads per block of the trace kernel
constexpr int kMaxSlots = 32;      // PRT_MAX_SLOTS
constexpr int kMaxDepth = 6;       // simultaneously live hit lists while evaluating one component
constexpr int kFrameCols = 15;
// A staged record (trace kernel -> ordering pass) holds what only the trace knows about a row: start
// position (3), direction as traced (3), hit distance, refractive index before the interaction, and one
// word with the leaf hit and the ray's slot in its tile.  The ordering pass rebuilds the 15 frame columns
// from it, bit for bit (same functions, same -fmad=false arithmetic): 72 instead of 120 bytes per row.
constexpr int kStageCols = 9;  // PRT_STAGE_COLS
// slack (world units) of the exact component pruning: hit parameters and box parameters of the
// same point differ by rounding only (~1e-13 at scene scale), the ray offset is 1e-6
constexpr double kCullMargin = 1e-7;
// Scenes with more boxed components than this are traversed in ray order (OrderEntry tables + bisection).
// Smaller scenes keep the list-order loop: every lane of a warp then looks at the same component in the same
// iteration, which keeps the primitive switch uniform when the few components are of different kinds.
constexpr int kOrderedMinBoxed = 8;
// Encoded scenes up to this size are staged in shared memory by every block of the trace kernel (two blocks
// per SM share 228 KB); larger ones are read from global memory through L1 / L2 (trace_kernel<.., GLOBAL>).
constexpr int kMaxSharedBlob = 100 * 1024;

enum OpKind : int {
  OP_LEAF = 0,        // a = leaf              : push the leaf's hit pair
  OP_ENTER = 1,       // a = aabb, b = skip op, c = flags : world-space box test; on miss push an empty list and
                      // jump to b.  c & 1 (root only): the box provably contains the solid -> pruning allowed
  OP_MERGE = 2,       // a = csg operation     : pop R, pop L, push array_csg(L, R)
  OP_MERGE_LEAF = 3   // a = csg operation, b = leaf : top := array_csg(top, leaf pair)  (right child is a leaf)
};

// evaluation strategy of a component, chosen by the encoder
enum Shape : int {
  SHAPE_GENERIC = 0,  // any tree: preorder interpreter with hit lists in local memory
  SHAPE_LEAF = 1,     // bare TracerSurface
  SHAPE_LEFT2 = 2,    // ENTER, LEAF a, MERGE_LEAF b                     (A op B)
  SHAPE_LEFT3 = 3     // ENTER, ENTER, LEAF a, MERGE_LEAF b, MERGE_LEAF c  ((A op1 B) op2 C)
};

// Everything nearest_hit needs about one component, in one place (no pointer chasing through the
// op list): shape, proven-bound flag, the op range for the generic interpreter and, for the
// left-deep shapes, the leaves, operations and both bounding boxes.
struct Comp {
  int shape;   // Shape
  int flags;   // bit 0: the root box provably contains the solid -> pruning allowed
               // bit 1: the solid is convex (INTERSECT of convex primitives): a ray that has just left
               //        it through one of its faces cannot hit it again before it changes direction
               // bit 2: SHAPE_LEAF only: root_box is a conservative world box of the bare surface (quick prune)
               // bit 3 / 4: SHAPE_LEFT3 of a capped Cylinder and two Spheres with the cylinder as leaf a (bit 3,
               //        thick_lens) or leaf c (bit 4, biconvex_lens): evaluated in one block by lens3_hits_fast
  int begin, end;                  // ops [begin, end)
  int leaf_a, leaf_b, leaf_c;      // SHAPE_LEAF: leaf_a; SHAPE_LEFT2: a, b; SHAPE_LEFT3: a, b, c
  int op1, op2;                    // (A op1 B) op2 C
  int tt;                          // truth table of F(inA, inB, inC) = op2(op1(inA, inB), inC), bit a | b<<1 | c<<2
                                   // (SHAPE_LEFT2: op1(inA, inB)); see left_deep_first_hit
  int pad[2];
  double root_box[6];
  double inner_box[6];             // SHAPE_LEFT3: box of (A op1 B)
};

struct Op {
  int kind, a, b, c;
};

// One entry of a traversal table.  For each of the six (axis k, direction sign s) pairs the components that
// carry a proven / conservative root box ("boxed": Comp.flags & 5) are listed in the order a ray travelling
// along that axis meets them: ascending near face in the signed coordinate u = x_k (s = 0, ray towards +k)
// or u = -x_k (s = 1).  pmfar_u is the running maximum of far_u over the entries up to and including this
// one, so "everything before here lies behind the ray" is a bisection on a monotone column.
struct OrderEntry {
  double near_u, far_u, pmfar_u;
  int comp, pad;
};

struct Leaf {
  double m[12];    // rows 0..2 of the world->object 4x4 (row-major 3x4)
  double prm[6];   // primitive parameters (see prt_prim)
  double matp[6];  // material parameters (see prt_material)
  double nscale;   // Intersectable._normal_scale
  double sid;      // surface id as it appears in the frame (float64 column)
  int type;        // prt_prim
  int mat;         // prt_material
  int comp;        // component this leaf belongs to
  int pad;
};

struct BlobHeader {
  int n_components, n_ops, n_leaves, n_aabb;
  int off_ops;     // Op[n_ops]
  int off_aabb;    // double[n_aabb*6]
  int off_leaves;  // Leaf[n_leaves]
  int total_bytes;
  int max_slots;   // largest component hit-list length
  int flags;       // bit 0: every bounding-box span is 0 or in [2^-823, 2^677) (fast slab test allowed)
                   // bit 1: some component has SHAPE_GENERIC (needs the interpreter kernel variant)
                   // bit 2: every component root box lies within +-1e6 (dominant-axis quick prune allowed)
                   // bit 3: walk the boxed components in the order the ray meets them (many components);
                   //        otherwise visit every component in list order, in lockstep across the warp
  int off_comps;   // Comp[n_components]
  int n_boxed;     // entries per traversal table
  int n_unboxed;   // components without a usable box: always evaluated
  int off_order;   // OrderEntry[6][n_boxed], table index 2 * axis + sign
  int off_unboxed; // int[n_unboxed]
  int off_bycomp;  // OrderEntry[6][n_components]: the same entries in list order (unboxed: near -inf, far +inf)
  int pad1, pad2;
};

// arguments of the trace kernel (filled by prt_trace)
struct TraceArgs {
  const unsigned char* blob;
  int blob_bytes;
  int generation_limit;
  int record_mode;
  double ray_offset;
  double detector_sid;
  const double* rays;
  long long n_rays;
  long long stride;
  double* stage;
  long long capacity;
  long long* run_start;
  int* run_count;
  long long n_tiles;
  prt_counters* ctr;
};

// arguments of the ordering pass (gather_kernel), filled by prt_gather_frame
struct GatherArgs {
  const unsigned char* blob;  // scene blob in device memory (leaf -> surface id)
  const double* rays;         // the traced RaySet: generation / intensity / wavelength / id columns come from it
  long long n_rays, ray_stride;
  const double* stage;
  long long capacity;
  const long long* run_start;
  const int* run_count;
  const long long* run_base;
  long long n_tiles;
  const long long* gen_offsets;
  int generation_limit;
  double* frame;
  long long frame_stride;
  long long frame_capacity;   // rows the frame can hold (rows beyond it are not written)
};

// arguments of the wavefront driver (prt_wavefront.cu), filled by prt_trace_wavefront
struct WaveArgs {
  const unsigned char* blob;
  int blob_bytes;
  int generation_limit;
  int record_mode;
  int g;
  double ray_offset;
  double detector_sid;
  const double* rays;
  long long n_rays, stride;
  double* st;  // rows p0,p1,p2,v0,v1,v2,nidx: st[k*n_rays + i]
  int* flag;
  double* hit_t;
  int* hit_leaf;
  int* blk_count;       // [2 * n_tiles] by generation parity
  long long* blk_base;  // [2 * n_tiles]
  long long* alive;    // [generation_limit + 1]: live rays entering generation g (alive[0] unused)
  long long* gen_off;  // [generation_limit + 1]
  long long n_tiles;
  double* frame;
  long long frame_stride, capacity;
  prt_counters* ctr;
};

}  // namespace prt
