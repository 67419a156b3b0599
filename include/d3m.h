/*
 * d3m.h -- C ABI of the B200-native NeuralRecon lifting hot path (libd3m.so).
 *
 * Drop-in boundary for two reference entry points of achao2013/deep3dmap (paths relative to the
 * reference tree):
 *   back_project(coords, origin, voxel_size, feats, KRcam)      deep3dmap/core/voxel/back_project.py:5-84
 *   TSDFVolume.__init__/integrate/get_volume                    deep3dmap/core/tsdf/tsdf_volume.py:14-307
 *   TSDFVolumeTorch.integrate (dataloader variant)              deep3dmap/core/tsdf/tsdf_volume.py:437-574
 * and the callers either side of that path:
 *   generate_grid / NeuConNet.upsample / get_target / level glue   deep3dmap/core/voxel/generate_grids.py:4-11,
 *                                                                deep3dmap/models/neucon_network.py:52-89, 113-207
 *   GRUFusion sparse<->dense movement, sparse_to_dense_*           deep3dmap/models/modulars/gru_fusion.py:51-181,
 *                                                                deep3dmap/core/utils/neucon_utils.py:114-131
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA-runtime types in the signatures.  `stream` is a
 *     cudaStream_t passed as void* (NULL = legacy default stream).
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`.
 *   - every call returns 0 on success; non-zero = D3M_ERR_* (argument errors) or 1000 + cudaError_t.
 *     `d3m_last_error()` returns a thread-local, human-readable message for the last failure.
 *   - calls are asynchronous on `stream` unless stated otherwise; they never synchronise the device
 *     except `d3m_tsdf_download` and `d3m_tsdf_create/destroy`.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     D3M_ERR_NO_DEVICE.
 */
#ifndef D3M_H_
#define D3M_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D3M_VERSION 111

enum {
  D3M_OK = 0,
  D3M_ERR_ARG = 1,        /* bad shape / NULL pointer / unsupported size */
  D3M_ERR_WORKSPACE = 2,  /* workspace too small */
  D3M_ERR_NO_DEVICE = 3,
  D3M_ERR_ALIGN = 4,      /* pointer not 16-byte aligned */
  D3M_ERR_CUDA = 1000     /* + cudaError_t */
};

/* dtype of the (N,4) [batch,x,y,z] coordinate rows.  The reference passes float32 from
 * generate_grid (core/voxel/generate_grids.py:9) and int64 when coords come from torch.nonzero
 * (models/modulars/gru_fusion.py:294-300); int32 is an extension for >2^24-voxel scenes. */
enum { D3M_COORDS_F32 = 0, D3M_COORDS_I64 = 1, D3M_COORDS_I32 = 2 };

int d3m_version(void);
const char* d3m_last_error(void);
/* number of CUDA devices visible (0 when none / driver missing); never fails */
int d3m_device_count(void);
/* the calling thread's current CUDA device (what `.cuda()` / pycuda.autoinit would pick in the reference), -1 without one */
int d3m_current_device(void);

/* Diagnostics used by bench.py (not part of the reference surface).
 *   d3m_kernel_launches: kernels this library has launched in this process since load.
 *   d3m_profile_begin / d3m_profile_end: while enabled, one CUDA-event pair is recorded on the launching
 *   stream around every kernel launch; _end synchronises, writes a JSON object
 *   {"<kernel>": {"n": launches, "ms": total device milliseconds}, ...} into `json_out` and disables it. */
int64_t d3m_kernel_launches(void);
int d3m_profile_begin(void);
int d3m_profile_end(char* json_out, size_t cap);

/* ---------------------------------------------------------------------------------------------
 * Feature-map layout.  The kernels read / write per-view maps channels-last: (V,B,H,W,C) so that one
 * texel is C contiguous floats (128-bit loads).  The reference hands over (V,B,C,H,W) as produced
 * by torch.stack (models/neucon_network.py:128); these two kernels convert n_maps = V*B maps.
 * ------------------------------------------------------------------------------------------- */
int d3m_feats_nchw_to_nhwc(const float* src, float* dst, int64_t n_maps, int C, int H, int W, void* stream);
int d3m_feats_nhwc_to_nchw(const float* src, float* dst, int64_t n_maps, int C, int H, int W, void* stream);

/* ---------------------------------------------------------------------------------------------
 * back_project forward  (replaces back_project.py:23-84)
 *   coords      (N,4) rows [batch, x, y, z] in voxel units, dtype per coords_kind
 *   origin      (B,3) float32 metres            voxel_size  float (python float in the reference)
 *   KRcam       (V,B,4,4) float32 world->pixel
 *   out         (N,C+1) float32: view-mean features | normalised mean depth   (written for every row;
 *               rows whose batch index is outside [0,B) are zero, as in the reference)
 *   count       (N,) float32: number of views that see the voxel (bit-exact contract)
 *   feats       the per-view maps, float32, in either layout (feats_layout):
 *                 D3M_FEATS_NCHW  (V,B,C,H,W) as torch.stack hands them over (models/neucon_network.py:128); then
 *                                 feats_nhwc_scratch (V*B*H*W*C floats, 16-byte aligned) receives the channels-last copy
 *                                 the gather reads (written by the same launch that clears the binning state);
 *                 D3M_FEATS_NHWC  (V,B,H,W,C), used in place; feats_nhwc_scratch may be NULL.
 *   cell_hist   NULL, or d3m_back_project_cell_hist_elems(N,B,V,H,W) int32 (16-byte aligned), contents arbitrary on entry:
 *               the BINNING STATE the deterministic backward starts from -- samples per bin = (bilinear cell (v,b,y0,x0),
 *               voxel-index bucket), histogrammed by the gather (which projects every voxel anyway) and already scanned
 *               on return.  Pass it to d3m_back_project_bwd for the SAME coords / KRcam (the autograd wrapper does when
 *               feats needs grad); backward may then run any number of times on it.
 * workspace: d3m_back_project_fwd_workspace(N,B,V,C) bytes, 256-byte aligned.
 * Launches per call: relayout+clear, gather (+ depth statistics when B == 1), scan+normalise -- three.
 * ------------------------------------------------------------------------------------------- */
enum { D3M_FEATS_NHWC = 0, D3M_FEATS_NCHW = 1 };
size_t d3m_back_project_fwd_workspace(int64_t N, int B, int V, int C);
size_t d3m_back_project_cell_hist_elems(int64_t N, int B, int V, int H, int W);
int d3m_back_project_fwd(const void* coords, int coords_kind, int64_t N, const float* origin, int B,
                         float voxel_size, const float* feats, int feats_layout, float* feats_nhwc_scratch,
                         int V, int C, int H, int W, const float* KRcam, float* out, float* count, int* cell_hist,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Voxel-range sharding (BASELINE config 5): every rank runs the gather on its contiguous slice of the coordinate
 * list; the only cross-voxel coupling of the reference, the per-fragment depth normalisation
 * (back_project.py:77-80), is split in two so that the caller can all-reduce three fp64 scalars per fragment
 * in between:
 *   _partial  as d3m_back_project_fwd, but leaves the RAW mean depth in out[:,C] and writes this slice's
 *             (sum z, sum z^2, #{z>0}) per fragment to depth_sums (B,3) float64 on the device;
 *   _finish   takes the (all-reduced) sums, derives mean / L2-norm and normalises out[:,C] in place.  `workspace`
 *             must be the buffer handed to _partial for the same slice, untouched in between. */
int d3m_back_project_fwd_partial(const void* coords, int coords_kind, int64_t N, const float* origin, int B,
                                 float voxel_size, const float* feats, int feats_layout, float* feats_nhwc_scratch,
                                 int V, int C, int H, int W, const float* KRcam, float* out, float* count,
                                 int* cell_hist, double* depth_sums, void* workspace, size_t workspace_bytes,
                                 void* stream);
int d3m_back_project_fwd_finish(int64_t N, int B, int C, const double* depth_sums, float* out, void* workspace,
                                size_t workspace_bytes, void* stream);
/* _partial with the all-gather of the view counts fused into the gather kernel: every voxel's count is also stored into
 * every rank's full-scene count buffer (peer_count_host: HOST array of `world` device pointers, IPC-mapped, float32) at the
 * voxel's global row -- begin + n for a contiguous range (block == 0), ((n / block) * world + rank) * block + n % block for
 * block-cyclic ranges.  Complete everywhere after the caller's next sync across the ranks (d3m_p2p_sync). */
typedef struct {
  void* const* peer_count_host;
  int world, rank;
  int64_t begin, block;
} d3m_count_exchange;
int d3m_back_project_fwd_partial_x(const void* coords, int coords_kind, int64_t N, const float* origin, int B,
                                   float voxel_size, const float* feats, int feats_layout, float* feats_nhwc_scratch,
                                   int V, int C, int H, int W, const float* KRcam, float* out, float* count,
                                   int* cell_hist, double* depth_sums, void* workspace, size_t workspace_bytes,
                                   const d3m_count_exchange* cx, void* stream);

/* ---------------------------------------------------------------------------------------------
 * back_project backward w.r.t. feats  (replaces autograd through back_project.py:55-73:
 * div backward, mask, grid_sampler_2d_backward).  Deterministic: samples are binned per texel with
 * integer atomics only, each bin is ordered by voxel index, and every texel is accumulated by one
 * lane group in a fixed order -- no floating-point atomics anywhere.  The result is a pure function of the inputs
 * (bit-identical run to run and independent of whether count / cell_hist were handed over).
 *   grad_out         (N,C+1) float32 (the depth column carries no gradient to feats)
 *   count            (N,) float32 as returned by d3m_back_project_fwd for the same inputs, or NULL
 *                    (then the view counts are recomputed by one extra kernel)
 *   cell_hist        binning state produced by d3m_back_project_fwd for the same inputs (its claim counters are used
 *                    and handed back cleared, so backward may run more than once), or NULL (then the state is rebuilt
 *                    in the workspace: clear + projection/histogram pass + scan, three extra launches)
 * Launches per call with count and cell_hist: fill + pre-division, order (rank sort of every bin), gather -- three.
 *   grad_feats       float32, fully overwritten; (V,B,H,W,C) when grad_nchw == 0, the reference's
 *                    (V,B,C,H,W) when grad_nchw != 0 (the gather kernel then stores channel-strided)
 * ------------------------------------------------------------------------------------------- */
size_t d3m_back_project_bwd_workspace(int64_t N, int B, int V, int C, int H, int W);
int d3m_back_project_bwd(const void* coords, int coords_kind, int64_t N, const float* origin, int B,
                         float voxel_size, int V, int C, int H, int W, const float* KRcam,
                         const float* grad_out, const float* count, int* cell_hist, float* grad_feats,
                         int grad_nchw, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Voxel-range sharding over the GPUs of one box (BASELINE config 5): backward with a fused exchange.
 * Every rank holds a slice of the voxel list and the (replicated) feature maps of ALL views, so its backward produces a
 * partial gradient of every view.  Instead of materialising that partial gradient and all-reducing it (118 MB at 64 views),
 * the gather kernel stores each texel tile straight into the staging buffer of the rank that OWNS the view
 * (owner(v) = v / ceil(V / world)), into the slot of the sending rank -- peer memory over NVLink, overlapped with the
 * gather itself.  After a barrier across the ranks (any collective on the same streams) every owner adds its `world`
 * slots in rank order: the reduce-scatter-by-view a view-parallel 2D backbone consumes, deterministic.
 *   d3m_p2p_*                   staging memory: cudaMalloc + CUDA IPC handle (64 bytes, exchanged by the caller's process
 *                               group), opened once per peer.
 *   d3m_back_project_bwd_exchange   as d3m_back_project_bwd, but grad goes to peer_staging_host[owner] (host array of
 *                               `world` device pointers; entry `rank` is this rank's own buffer).  Staging layout of one
 *                               rank: (world, ceil(V/world), B, H, W, C) float32, channels-last.
 *   d3m_grad_slots_sum          owner side: out (n_views, B, C, H, W) = sum over the `world` slots, ascending rank order.
 * ------------------------------------------------------------------------------------------- */
int d3m_p2p_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64_host);
int d3m_p2p_open(const unsigned char* handle64_host, void** dev_ptr);
int d3m_p2p_close(void* dev_ptr);
int d3m_p2p_free(void* dev_ptr);
int d3m_back_project_bwd_exchange(const void* coords, int coords_kind, int64_t N, const float* origin, int B,
                                  float voxel_size, int V, int C, int H, int W, const float* KRcam,
                                  const float* grad_out, const float* count, int* cell_hist,
                                  void* const* peer_staging_host, int world, int rank, void* workspace,
                                  size_t workspace_bytes, void* stream);
int d3m_grad_slots_sum(const float* staging, int world, int views_per_owner, int n_views, int B, int C, int H, int W,
                       float* out_nchw, void* stream);
/* All-gather of per-voxel rows (view counts / occupancy for the next coarse-to-fine level, neucon_network.py:132,180-196)
 * as peer stores: this rank's n_local rows of row_bytes bytes go into EVERY rank's full buffer (peer_dst_dev_table: DEVICE
 * array of `world` device pointers) at their global positions -- begin + i for contiguous ranges (block == 0), or the
 * block-cyclic map ((i / block) * world + rank) * block + i % block.  Complete on every rank after the caller's next
 * collective on the same streams. */
int d3m_p2p_scatter_rows(const void* src, int64_t n_local, int row_bytes, int64_t begin, int64_t block,
                         void* const* peer_dst_dev_table, int world, int rank, void* stream);
/* All-reduce (sum, ascending rank order) of n <= 192 doubles AND barrier across the ranks of one box through peer memory:
 * one small kernel per rank writes its payload + a flag (`epoch`, the caller's call counter, > 0 and equal on all ranks)
 * into every rank's mailbox and waits for the others' flags.  Peer stores of earlier kernels on the same stream (count
 * rows, gradient slots) have arrived wherever the flag has.  mailbox_dev_table: DEVICE array of `world` pointers to the
 * ranks' mailboxes (d3m_p2p_sync_mailbox_bytes(world) bytes each, zero-filled once).  n == 0: barrier only. */
size_t d3m_p2p_sync_mailbox_bytes(int world);
int d3m_p2p_sync(void* const* mailbox_dev_table, int world, int rank, unsigned long long epoch, const double* payload, int n,
                 double* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * TSDF fusion  (replaces TSDFVolume, tsdf_volume.py:10-307, and TSDFVolumeTorch :485-574)
 * The handle owns three (X,Y,Z) C-order float32 volumes on `device` (tsdf=1, weight=0, color=0:
 * tsdf_volume.py:50-53), a pinned staging ring for host frames and the frame workspace.
 * ------------------------------------------------------------------------------------------- */
typedef struct d3m_tsdf d3m_tsdf;

/* integrate() arithmetic flavours */
enum {
  D3M_TSDF_KERNEL_SEMANTICS = 0, /* the reference CUDA kernel :68-126: roundf, cam_z<0 rejects, depth==0 rejects,
                                    R^T(p-t) from cam_pose, nvcc FMA contraction of that source */
  D3M_TSDF_TORCH_SEMANTICS = 1,  /* TSDFVolumeTorch :437-482: half-to-even, cam_z>0, depth>0, `pose` argument
                                    is the fp32 world->camera matrix inverse(cam_pose), no contraction */
  D3M_TSDF_WITH_COLOR = 2        /* flag (OR-ed): also run the colour average of :130-141, which the
                                    reference kernel never reaches (:129); default keeps colour == 0 */
};

int d3m_tsdf_create(int dim_x, int dim_y, int dim_z, const float* origin3_host, float voxel_size,
                    float trunc_margin, int device, d3m_tsdf** out_handle);
/* x-slab of a larger volume (multi-GPU slab sharding): the handle owns planes [x_begin, x_begin+dim_x_local) of a
 * volume whose voxel (0,0,0) sits at origin3_host; world positions are computed from the GLOBAL voxel index, so
 * the slabs of all ranks concatenated along x are bit-identical to the unsharded volume. */
int d3m_tsdf_create_slab(int dim_x_local, int dim_y, int dim_z, int x_begin, const float* origin3_host,
                         float voxel_size, float trunc_margin, int device, d3m_tsdf** out_handle);
int d3m_tsdf_destroy(d3m_tsdf* h);
/* device ordinal the handle's volumes live on.  Every d3m_tsdf_* call makes that device current for its own duration
 * and restores the caller's current device before returning. */
int d3m_tsdf_device(d3m_tsdf* h);
int d3m_tsdf_reset(d3m_tsdf* h, void* stream);
/* Re-use a handle for another volume of the same dimensions: new origin / voxel size / truncation, volumes reset
 * (tsdf = 1, weight = colour = 0).  The dataloader transform (transforms_seq.py:355-357) builds three small
 * volumes per sample; creating and destroying handles costs ~10 ms each (cudaMalloc / cudaMallocHost / cudaFree). */
int d3m_tsdf_rebase(d3m_tsdf* h, const float* origin3_host, float voxel_size, float trunc_margin, void* stream);

/* One frame from HOST memory -- the shape of TSDFVolume.integrate (:210-256): depth (H,W) float32
 * metres, optional colour (H,W) float32 already folded as b*65536+g*256+r (:223-227), intrinsics 3x3
 * and pose 4x4 row-major float32.  Copies through the pinned ring, then runs the frame kernels. */
int d3m_tsdf_integrate_host(d3m_tsdf* h, const float* depth_host, const float* color_host, int H, int W,
                            const float* intr9_host, const float* pose16_host, float obs_weight,
                            int flags, void* stream);

/* n_frames frames already resident on the device, integrated in order inside ONE launch (each voxel
 * block keeps its tsdf/weight in registers across the frames that touch it).
 *   depth (n_frames,H,W); color (n_frames,H,W) or NULL; intr9_host (n_frames,9) or (1,9) when
 *   intr_per_frame == 0; pose16_host (n_frames,16); obs_weight_host (n_frames) or NULL (= 1). */
int d3m_tsdf_integrate_device(d3m_tsdf* h, const float* depth, const float* color, int n_frames, int H,
                              int W, const float* intr9_host, int intr_per_frame,
                              const float* pose16_host, const float* obs_weight_host, int flags,
                              void* stream);

/* device pointers of the volumes (valid until destroy) and a blocking copy to host (get_volume, :302-307) */
int d3m_tsdf_volumes(d3m_tsdf* h, float** tsdf, float** weight, float** color);
int d3m_tsdf_download(d3m_tsdf* h, float* tsdf_host, float* weight_host, float* color_host, void* stream);
/* voxels whose weight changed in the last integrate call is not tracked; this returns the number of
 * tile launches of the last call (diagnostics for bench.py's gpu_launches) */
int d3m_tsdf_last_launches(d3m_tsdf* h);

/* Host (pageable) -> device copy of `bytes` bytes, stream-ordered on `stream`.  The frames and scene volumes of this path
 * arrive as pageable numpy / torch CPU memory (tools/data_gen/scannet.py:84-100 hands cv2 frames to integrate();
 * datasets/pipelines/transforms_seq.py:343-396 hands `tsdf_list_full`), for which a plain cudaMemcpy is a single-threaded
 * staging copy (~10 GB/s).  Here the buffer is cut into chunks that a small thread pool copies into a two-slot pinned ring
 * while the previous chunk is in flight on the copy engine.  Returns when the last chunk is staged: `host_src` may be
 * reused at once, the data is in `dev_dst` for everything ordered after the call on `stream`. */
int d3m_upload(const void* host_src, void* dev_dst, size_t bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Mesh export: marching cubes over a dense (X,Y,Z) float32 volume on the device (replaces the
 * skimage.measure.marching_cubes[_lewiner](tsdf_vol, level=0) calls of tsdf_volume.py:315,335 and
 * core/utils/neucon_utils.py:177).  Two passes around the ordered compaction d3m_compact:
 *   d3m_mc_flags   edge_flags (X*Y*Z, 3) uint8: the edge leaving voxel v along +x/+y/+z crosses the level ("inside" = value
 *                  < level); tri_flags (X*Y*Z, K) uint8, K = d3m_mc_max_triangles_per_cube(): slot k of the cube whose low
 *                  corner is v holds a triangle
 *   d3m_mc_emit    edge_list / tri_list = d3m_compact of those flags (int64, ascending); writes verts (n_verts,3) in voxel
 *                  coordinates, unit normals (n_verts,3) pointing towards larger values, faces (n_faces,3) int32 vertex ids
 *                  (counter-clockwise seen from the larger-value side), and edge_to_vertex (X*Y*Z*3 int32 scratch).
 * Vertices are the level crossings of the grid edges (linear interpolation), shared between faces; the case table is
 * generated by deep3dmap_b200/mc_tables.py (ambiguous faces always separate the inside corners: watertight).
 * ------------------------------------------------------------------------------------------- */
int d3m_mc_max_triangles_per_cube(void);
int d3m_mc_flags(const float* volume, int X, int Y, int Z, float level, uint8_t* edge_flags, uint8_t* tri_flags, void* stream);
int d3m_mc_emit(const float* volume, int X, int Y, int Z, float level, const int64_t* edge_list, int64_t n_verts,
                const int64_t* tri_list, int64_t n_faces, int* edge_to_vertex, float* verts, float* normals, int* faces,
                void* stream);

/* =============================================================================================
 * SURVEY section 8 row f1, ground-truth side of the dataloader transform
 * (deep3dmap/datasets/pipelines/transforms_seq.py:343-398, SeqRandomTransformSpace.transform).
 * ========================================================================================== */

/* :365-366  occ[i] = (tsdf[i] < hi) & (tsdf[i] > lo) & (weight[i] > min_weight)  on device arrays of n voxels
 * (the reference uses lo = -0.999, hi = 0.999, min_weight = 1 on the volumes of TSDFVolumeTorch). */
int d3m_tsdf_occupancy(const float* tsdf, const float* weight, int64_t n, float lo, float hi, float min_weight,
                       uint8_t* occ, void* stream);

/* :343-396  re-sample the full-scene TSDF `tsdf_full` (X,Y,Z) of level l on the transformed fragment grid:
 *   c = ((transform[:3,:] @ [idx*step*voxel_size + vol_origin_partial, 1]) - old_origin) / voxel_size / step
 *   g = 2*c/(dim-1) - 1 ;  nearest and trilinear 3-D grid_sample (align_corners=False, zero padding);
 *   out = trilinear where |nearest| < 1 else nearest ;  out = 1 where any |g| >= 1.
 * out is (nx,ny,nz) = voxel_dim / step, step = 2^l; the three small parameter arrays are HOST pointers
 * (transform12 = the first three rows of the 4x4 matrix, row-major). */
int d3m_gt_recrop(const float* tsdf_full, int X, int Y, int Z, int nx, int ny, int nz, int step, float voxel_size,
                  const float* vol_origin_partial3_host, const float* transform12_host,
                  const float* old_origin3_host, float* out, void* stream);

/* =============================================================================================
 * the steps either side of back_project in the coarse-to-fine level loop
 * (deep3dmap/models/neucon_network.py:113-207).  Integer / index work is bit-exact.
 * ========================================================================================== */

/* generate_grid (core/voxel/generate_grids.py:4-11) and the [b | x y z] rows built from it for every fragment
 * (models/neucon_network.py:118-122).  Per axis the grid holds arange(0, n, interval), x slowest, z fastest.
 *   coords  NULL or (B*n, 4) float32 rows [b, x, y, z], fragment-major;  grid3  NULL or (3, n) float32 planes. */
int d3m_grid_coords(int nx, int ny, int nz, int interval, int B, float* coords, float* grid3, void* stream);

/* NeuConNet.upsample (models/neucon_network.py:68-89): every voxel becomes `num` (1..8) children, child i adding
 * `interval` to the axes of pos_list[i-1]; features are repeated.  Either half may be skipped with NULL.
 *   pre_coords (N,4) of coords_kind -> up_coords (N*num,4) same dtype;  pre_feat (N,C) -> up_feat (N*num,C). */
int d3m_upsample(const void* pre_coords, int coords_kind, const float* pre_feat, int64_t N, int C, int interval,
                 int num, void* up_coords, float* up_feat, void* stream);

/* models/neucon_network.py:143-154: r = [xyz*voxel_size + origin[b], 1] @ world_to_aligned_camera[b,:3,:]^T written
 * as rows [rx, ry, rz, b] float32 (batch index moved to the last column).  world_to_aligned_camera is (B,4,4). */
int d3m_aligned_camera_coords(const void* coords, int coords_kind, int64_t N, const float* origin, int B,
                              float voxel_size, const float* world_to_aligned_camera, float* r_coords, void* stream);

/* NeuConNet.get_target (models/neucon_network.py:52-65): tsdf_out[n] = tsdf_vol[b, x//d, y//d, z//d] and the same for
 * the boolean occupancy volume, volumes (B,X,Y,Z).  Negative indices wrap once like torch; rows that are still out of
 * range (torch raises IndexError) are counted in *bad_rows (device int) and left unwritten. */
int d3m_gather_targets(const void* coords, int coords_kind, int64_t N, int divisor, const float* tsdf_vol,
                       const uint8_t* occ_vol, int B, int X, int Y, int Z, float* tsdf_out, uint8_t* occ_out,
                       int* bad_rows, void* stream);

/* models/neucon_network.py:181-182: flags[n] = occ[n*occ_stride] > threshold  and  grid_mask[n], the mask being an
 * explicit bool array and / or `count[n] > min_count` (":132 grid_mask = count > 1"); NULL = not applied. */
int d3m_occupancy_flags(const float* occ, int64_t occ_stride, const float* count, float min_count,
                        const uint8_t* grid_mask, float threshold, int64_t N, uint8_t* flags, void* stream);

/* Ordered stream compaction == torch.nonzero(flags) / boolean-mask indexing: out[k] = position (or values[position])
 * of the k-th set (invert != 0: cleared) flag, *total_dev = their number.  `out` must hold N entries.  Deterministic. */
size_t d3m_compact_workspace(int64_t N);
int d3m_compact(const uint8_t* flags, int64_t N, int invert, const int64_t* values, int64_t* out, int64_t* total_dev,
                void* workspace, size_t workspace_bytes, void* stream);

/* models/neucon_network.py:190-194 (training-time subsampling): keep[r] = 1 for r < n_keep, then keep[choice[j]] = 0;
 * compacting the index list with `keep` equals `occupancy[ind[choice]] = False`.  *bad counts out-of-range choices. */
int d3m_drop_ranks(const int64_t* choice, int64_t n_choice, int64_t n_keep, uint8_t* keep, int* bad, void* stream);

/* dst[m] = src[ind[m]] for rows of row_bytes (multiple of 4) bytes; and the fused
 * cat([a[ind], b[ind], ...], dim=1) of models/neucon_network.py:203-207 for up to 4 float32 sources
 * (ind == NULL: identity, i.e. a plain row-wise concatenation). */
int d3m_gather_rows(const void* src, int64_t row_bytes, const int64_t* ind, int64_t M, void* dst, void* stream);
int d3m_gather_concat(const float* const* srcs_host, const int* widths_host, int n_src, const int64_t* ind, int64_t M,
                      float* dst, void* stream);

/* rows per fragment (models/neucon_network.py:197-201 aborts the level when a fragment has none) */
int d3m_batch_counts(const void* coords, int coords_kind, int64_t N, int B, int64_t* counts, void* stream);

/* =============================================================================================
 * sparse <-> dense movement of the GRU-fusion global volume
 * (deep3dmap/models/modulars/gru_fusion.py:51-181, deep3dmap/core/utils/neucon_utils.py:114-131).
 * Coordinates are int64 (M,3) rows, volumes are (X,Y,Z,c) float32 C-order.
 * ========================================================================================== */

/* sparse_to_dense_torch / sparse_to_dense_channel: dense = full(default); dense[locs] = values ((M,c) rows, or
 * scalar_value when values == NULL).  With a workspace (d3m_sparse_to_dense_workspace bytes) duplicate locations are
 * resolved deterministically -- the last row wins, as on the reference's CPU path; without one the caller guarantees
 * unique locations.  Out-of-range rows are counted in *bad_rows and skipped. */
size_t d3m_sparse_to_dense_workspace(int X, int Y, int Z);
int d3m_sparse_to_dense(const int64_t* locs, int64_t M, const float* values, float scalar_value, int c,
                        float default_val, int X, int Y, int Z, float* dense, int* bad_rows, void* workspace,
                        size_t workspace_bytes, void* stream);

/* gru_fusion.py:83-91: shifted = global_coords - relative_origin; valid = inside the fragment bounding volume and,
 * when occupied_volume (X,Y,Z) is given (FUSION.FULL == False), occupied_volume[shifted] != 0. */
int d3m_fbv_mask(const int64_t* global_coords, int64_t M, const int64_t* relative_origin3_host, int X, int Y, int Z,
                 const float* occupied_volume, int64_t* shifted, uint8_t* valid, void* stream);

/* gru_fusion.py:100-106: flags[i] = any_c(pred(a[i,c])) | any_c(pred(b[i,c])), pred = (v != 0) for mode 0 and
 * (|v| < 1) for mode 1; vol_b may be NULL.  Compact the flags to get torch.nonzero's row-major order. */
int d3m_dense_union_flags(const float* vol_a, const float* vol_b, int64_t n_vox, int c, int mode, uint8_t* flags,
                          void* stream);

/* linear index -> rows ((x,y,z) + add3) * multiplier, optionally prefixed by batch_index (gru_fusion.py:145, 288) */
int d3m_unravel_coords(const int64_t* linear, int64_t M, int Y, int Z, const int64_t* add3_host, int64_t multiplier,
                       int with_batch, int64_t batch_index, int64_t* out, void* stream);

/* out[k,:] = volume[coords[k]] (gru_fusion.py:256-261); out-of-range rows counted in *bad_rows */
int d3m_dense_gather(const float* volume, int X, int Y, int Z, int c, const int64_t* coords, int64_t K, float* out,
                     int* bad_rows, void* stream);

/* dst = src + add3 on (M,3) int64 rows (gru_fusion.py:135) */
int d3m_coords_add(const int64_t* src, int64_t M, const int64_t* add3_host, int64_t* dst, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* D3M_H_ */
