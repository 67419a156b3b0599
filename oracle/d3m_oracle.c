/*
 * d3m_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference algorithms of the
 * NeuralRecon lifting hot path.  Never imported by the product package `deep3dmap_b200`; only
 * `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs use it.
 *
 * Parity status: the reference ships NO tests or golden vectors for this path,
 * so the oracle is pinned against outputs of the reference's own Python files executed in the build
 * container on synthetic inputs (`oracle/gen_golden.py` -> `tests/golden/*.npz`) and, for the TSDF
 * kernel, against the reference's own CUDA string compiled verbatim (`oracle/build_ref.py`).
 *
 * Every function cites the reference lines it restates.  All arithmetic is fp32 with explicit
 * rounding points: compile with -ffp-contract=off so that the ONLY fused multiply-adds are the
 * fmaf() calls written below.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -mavx2 -mfma -fopenmp -shared -fPIC
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { ORC_COORDS_F32 = 0, ORC_COORDS_I64 = 1, ORC_COORDS_I32 = 2 };
/* bilinear-sum flavour: 0 = aten CPU kernel (separate mul/add, GridSamplerKernel.cpp),
 *                       1 = aten CUDA kernel (out_acc += val*w contracted to FMA, GridSampler.cu) */
enum { ORC_INTERP_MULADD = 0, ORC_INTERP_FMA = 1 };
/* OR-ed into interp_mode: leave the raw mean depth (back_project.py:76) in out[:,C], i.e. stop before the
 * per-fragment normalisation of :77-80 (used by the voxel-range sharding tests, which all-reduce its sums) */
enum { ORC_RAW_DEPTH = 0x100 };

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* coords row -> (batch index or -1, xyz as float).  `back_project.py:29-30`: a row belongs to batch b
 * iff coords[:,0] == b; float coords are compared as floats, integer coords as integers. */
static inline int load_coord(const void* coords, int kind, int64_t n, int B, float xyz[3]) {
  int b = -1;
  if (kind == ORC_COORDS_F32) {
    const float* c = (const float*)coords + 4 * n;
    float bf = c[0];
    if (bf >= 0.0f && bf < (float)B && bf == floorf(bf)) b = (int)bf;
    xyz[0] = c[1]; xyz[1] = c[2]; xyz[2] = c[3];
  } else if (kind == ORC_COORDS_I64) {
    const int64_t* c = (const int64_t*)coords + 4 * n;
    if (c[0] >= 0 && c[0] < B) b = (int)c[0];
    xyz[0] = (float)c[1]; xyz[1] = (float)c[2]; xyz[2] = (float)c[3];
  } else {
    const int32_t* c = (const int32_t*)coords + 4 * n;
    if (c[0] >= 0 && c[0] < B) b = (int)c[0];
    xyz[0] = (float)c[1]; xyz[1] = (float)c[2]; xyz[2] = (float)c[3];
  }
  return b;
}

typedef struct {
  int valid;      /* in-frustum mask, back_project.py:49-51 */
  float z;        /* im_z */
  int x0, y0;     /* floor of un-normalised sample position */
  float fx, fy;   /* ix - x0, iy - y0 (exact) */
} orc_sample;

/* One voxel-view projection.  Restates back_project.py:37-51 and the align_corners=True
 * un-normalisation of aten grid_sampler (`((coord+1)/2)*(size-1)`).
 *   grid = coords*voxel_size + origin            two roundings (:37)
 *   im_p = KRcam @ [grid;1]                       K=4 dot product; torch CPU bmm == the FMA chain below
 *   im_x = px/pz, im_y = py/pz, im_z = pz          (:45-47)
 *   g    = 2*im_x/(w-1) - 1                        three roundings (:49)
 *   mask = |gx|<=1 & |gy|<=1 & im_z>0              (:50-51)                                   */
static inline orc_sample project(const float xyz[3], const float* org, float vs, const float* P, int H, int W) {
  orc_sample s;
  float gx = xyz[0] * vs; gx = gx + org[0];
  float gy = xyz[1] * vs; gy = gy + org[1];
  float gz = xyz[2] * vs; gz = gz + org[2];
  float p[3];
  for (int r = 0; r < 3; ++r) {
    float a = P[4 * r + 0] * gx;
    a = fmaf(P[4 * r + 1], gy, a);
    a = fmaf(P[4 * r + 2], gz, a);
    a = fmaf(P[4 * r + 3], 1.0f, a);
    p[r] = a;
  }
  float u = p[0] / p[2];
  float v = p[1] / p[2];
  s.z = p[2];
  float nx = (2.0f * u) / (float)(W - 1); nx = nx - 1.0f;
  float ny = (2.0f * v) / (float)(H - 1); ny = ny - 1.0f;
  s.valid = (fabsf(nx) <= 1.0f) && (fabsf(ny) <= 1.0f) && (s.z > 0.0f);
  s.x0 = s.y0 = 0; s.fx = s.fy = 0.0f;
  if (s.valid) {
    float ix = ((nx + 1.0f) / 2.0f) * (float)(W - 1);
    float iy = ((ny + 1.0f) / 2.0f) * (float)(H - 1);
    float fx0 = floorf(ix), fy0 = floorf(iy);
    s.x0 = (int)fx0; s.y0 = (int)fy0;
    s.fx = ix - fx0; s.fy = iy - fy0;
  }
  return s;
}

/* Forward.  Restates back_project.py:23-84 per voxel instead of per tensor op:
 *   features = grid_sample(bilinear, zeros, align_corners=True) (:55), invalid views zeroed (:61-62),
 *   count = mask.sum(0) (:64), sum over views in order / max(count,1) (:67-73),
 *   mean depth / max(count,1) (:76), normalised by mean and L2 norm over voxels with z>0 (:77-80).
 * feats (V,B,C,H,W), KR (V,B,4,4), out (N,C+1), count (N).  Rows whose batch index is outside
 * [0,B) stay zero (:25-26, :29).  The two global reductions (:77-78) are accumulated in double and
 * rounded once (torch's fp32 pairwise order is not restated; tolerance covers the last-ulp gap). */
int orc_back_project_fwd(const void* coords, int ckind, int64_t N, const float* origin, int B, float vs,
                         const float* feats, int V, int C, int H, int W, const float* KR,
                         float* out, float* count, int interp_mode) {
  const int64_t HW = (int64_t)H * W;
  const int raw_depth = (interp_mode & ORC_RAW_DEPTH) != 0;
  interp_mode &= 0xff;
  memset(out, 0, sizeof(float) * (size_t)N * (C + 1));
  memset(count, 0, sizeof(float) * (size_t)N);
  int* batch_of = (int*)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
  if (!batch_of) return 1;
#pragma omp parallel
  {
    float* acc = (float*)malloc(sizeof(float) * (size_t)C);
#pragma omp for schedule(static)
    for (int64_t n = 0; n < N; ++n) {
      float xyz[3];
      int b = load_coord(coords, ckind, n, B, xyz);
      batch_of[n] = b;
      if (b < 0) continue;
      for (int c = 0; c < C; ++c) acc[c] = 0.0f;
      int cnt = 0;
      float zsum = 0.0f;
      for (int v = 0; v < V; ++v) {
        const float* P = KR + ((int64_t)v * B + b) * 16;
        orc_sample s = project(xyz, origin + 3 * b, vs, P, H, W);
        if (!s.valid) continue;
        cnt += 1;
        zsum = zsum + s.z;
        const float* fm = feats + ((int64_t)v * B + b) * C * HW;
        float wx1 = s.fx, wy1 = s.fy;
        float wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;     /* == (x0+1)-ix exactly */
        float nw = wx0 * wy0, ne = wx1 * wy0, sw = wx0 * wy1, se = wx1 * wy1;
        int x0 = s.x0, y0 = s.y0, x1 = x0 + 1, y1 = y0 + 1;
        int in_nw = (x0 >= 0 && x0 < W && y0 >= 0 && y0 < H);
        int in_ne = (x1 >= 0 && x1 < W && y0 >= 0 && y0 < H);
        int in_sw = (x0 >= 0 && x0 < W && y1 >= 0 && y1 < H);
        int in_se = (x1 >= 0 && x1 < W && y1 >= 0 && y1 < H);
        for (int c = 0; c < C; ++c) {
          const float* ch = fm + (int64_t)c * HW;
          float val;
          if (interp_mode == ORC_INTERP_FMA) {
            val = 0.0f;
            if (in_nw) val = fmaf(ch[(int64_t)y0 * W + x0], nw, val);
            if (in_ne) val = fmaf(ch[(int64_t)y0 * W + x1], ne, val);
            if (in_sw) val = fmaf(ch[(int64_t)y1 * W + x0], sw, val);
            if (in_se) val = fmaf(ch[(int64_t)y1 * W + x1], se, val);
          } else {
            float a = in_nw ? ch[(int64_t)y0 * W + x0] * nw : 0.0f;
            float bq = in_ne ? ch[(int64_t)y0 * W + x1] * ne : 0.0f;
            float cq = in_sw ? ch[(int64_t)y1 * W + x0] * sw : 0.0f;
            float dq = in_se ? ch[(int64_t)y1 * W + x1] * se : 0.0f;
            val = a + bq; val = val + cq; val = val + dq;
          }
          acc[c] = acc[c] + val;
        }
      }
      float div = (float)(cnt > 0 ? cnt : 1);
      float* o = out + (int64_t)n * (C + 1);
      for (int c = 0; c < C; ++c) o[c] = acc[c] / div;
      o[C] = zsum / div; /* mean depth, normalised below */
      count[n] = (float)cnt;
    }
    free(acc);
  }
  /* per-batch depth normalisation, back_project.py:77-80 */
  for (int b = 0; b < B && !raw_depth; ++b) {
    double sum = 0.0; int64_t np_ = 0;
    for (int64_t n = 0; n < N; ++n)
      if (batch_of[n] == b && out[n * (C + 1) + C] > 0.0f) { sum += (double)out[n * (C + 1) + C]; ++np_; }
    float mean = np_ > 0 ? (float)(sum / (double)np_) : NAN;
    double ssq = 0.0;
    for (int64_t n = 0; n < N; ++n)
      if (batch_of[n] == b && out[n * (C + 1) + C] > 0.0f) {
        float d = out[n * (C + 1) + C] - mean;
        ssq += (double)d * (double)d;
      }
    float sd = (float)sqrt(ssq) + 1e-5f;
    for (int64_t n = 0; n < N; ++n) {
      if (batch_of[n] != b) continue;
      float z = out[n * (C + 1) + C];
      out[n * (C + 1) + C] = (z > 0.0f) ? (z - mean) / sd : 0.0f;
    }
  }
  free(batch_of);
  return 0;
}

/* Backward w.r.t. feats (the only input that carries grad: back_project.py:55 via autograd;
 * coords come from a no_grad block `neucon_network.py:78`, KRcam/origin are data).
 *   d features_sum = grad_out[:, :C] / max(count,1)           (div backward of :72)
 *   masked to valid views                                      (backward of :61)
 *   grid_sampler_2d_backward: grad_input[v,c,y,x] += w_corner * g   (mul, then add)
 * Accumulation order = the aten CPU kernel's (GridSamplerKernel.cpp, bilinear backward): a fragment's
 * voxels are walked in ascending order in chunks of `chunk` (= the ISA vector width: 8 AVX2, 16
 * AVX-512); per chunk and channel the nw corners of all lanes are scattered first, then ne, sw, se.
 * chunk = 1 gives the plain per-voxel nw,ne,sw,se order.  Views are independent.
 * The depth channel grad_out[:, C] never reaches feats.  grad_feats (V,B,C,H,W) is overwritten. */
#define ORC_MAX_CHUNK 64
int orc_back_project_bwd(const void* coords, int ckind, int64_t N, const float* origin, int B, float vs,
                         int V, int C, int H, int W, const float* KR, const float* grad_out,
                         float* grad_feats, int chunk) {
  const int64_t HW = (int64_t)H * W;
  if (chunk < 1) chunk = 1;
  if (chunk > ORC_MAX_CHUNK) return 2;
  memset(grad_feats, 0, sizeof(float) * (size_t)V * B * C * HW);
  float* cntf = (float*)calloc((size_t)(N > 0 ? N : 1), sizeof(float));
  int64_t* order = (int64_t*)malloc(sizeof(int64_t) * (size_t)(N > 0 ? N : 1)); /* voxels grouped by fragment */
  int64_t* start = (int64_t*)calloc((size_t)B + 1, sizeof(int64_t));
  if (!cntf || !order || !start) return 1;
  for (int64_t n = 0; n < N; ++n) {
    float xyz[3];
    int b = load_coord(coords, ckind, n, B, xyz);
    if (b >= 0) start[b + 1]++;
  }
  for (int b = 0; b < B; ++b) start[b + 1] += start[b];
  {
    int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)B);
    for (int b = 0; b < B; ++b) cur[b] = start[b];
    for (int64_t n = 0; n < N; ++n) {
      float xyz[3];
      int b = load_coord(coords, ckind, n, B, xyz);
      if (b >= 0) order[cur[b]++] = n;
    }
    free(cur);
  }
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < N; ++n) {
    float xyz[3];
    int b = load_coord(coords, ckind, n, B, xyz);
    if (b < 0) continue;
    int cnt = 0;
    for (int v = 0; v < V; ++v)
      cnt += project(xyz, origin + 3 * b, vs, KR + ((int64_t)v * B + b) * 16, H, W).valid;
    cntf[n] = (float)(cnt > 0 ? cnt : 1);
  }
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int v = 0; v < V; ++v) {
    for (int b = 0; b < B; ++b) {
      float* gm = grad_feats + ((int64_t)v * B + b) * C * HW;
      const float* P = KR + ((int64_t)v * B + b) * 16;
      for (int64_t i0 = start[b]; i0 < start[b + 1]; i0 += chunk) {
        int len = (int)((start[b + 1] - i0) < chunk ? (start[b + 1] - i0) : chunk);
        orc_sample smp[ORC_MAX_CHUNK];
        float wgt[4][ORC_MAX_CHUNK];
        int64_t off[4][ORC_MAX_CHUNK]; /* -1 = corner out of bounds / lane masked */
        int any = 0;
        for (int l = 0; l < len; ++l) {
          float xyz[3];
          int64_t n = order[i0 + l];
          load_coord(coords, ckind, n, B, xyz);
          smp[l] = project(xyz, origin + 3 * b, vs, P, H, W);
          for (int k = 0; k < 4; ++k) off[k][l] = -1;
          if (!smp[l].valid) continue;
          any = 1;
          float wx1 = smp[l].fx, wy1 = smp[l].fy, wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
          wgt[0][l] = wx0 * wy0; wgt[1][l] = wx1 * wy0; wgt[2][l] = wx0 * wy1; wgt[3][l] = wx1 * wy1;
          int x0 = smp[l].x0, y0 = smp[l].y0, x1 = x0 + 1, y1 = y0 + 1;
          int in_x0 = (x0 >= 0 && x0 < W), in_x1 = (x1 >= 0 && x1 < W);
          int in_y0 = (y0 >= 0 && y0 < H), in_y1 = (y1 >= 0 && y1 < H);
          if (in_x0 && in_y0) off[0][l] = (int64_t)y0 * W + x0;
          if (in_x1 && in_y0) off[1][l] = (int64_t)y0 * W + x1;
          if (in_x0 && in_y1) off[2][l] = (int64_t)y1 * W + x0;
          if (in_x1 && in_y1) off[3][l] = (int64_t)y1 * W + x1;
        }
        if (!any) continue;
        for (int c = 0; c < C; ++c) {
          float* ch = gm + (int64_t)c * HW;
          float g[ORC_MAX_CHUNK];
          for (int l = 0; l < len; ++l) {
            int64_t n = order[i0 + l];
            g[l] = grad_out[n * (C + 1) + c] / cntf[n];
          }
          for (int k = 0; k < 4; ++k)
            for (int l = 0; l < len; ++l)
              if (off[k][l] >= 0) { float t = wgt[k][l] * g[l]; ch[off[k][l]] += t; }
        }
      }
    }
  }
  free(cntf); free(order); free(start);
  return 0;
}

/* Valid-sample census used by bench/roofline bookkeeping: S = number of in-frustum voxel-view pairs. */
int64_t orc_back_project_valid_samples(const void* coords, int ckind, int64_t N, const float* origin, int B,
                                       float vs, int V, int H, int W, const float* KR) {
  int64_t S = 0;
#pragma omp parallel for reduction(+ : S) schedule(static)
  for (int64_t n = 0; n < N; ++n) {
    float xyz[3];
    int b = load_coord(coords, ckind, n, B, xyz);
    if (b < 0) continue;
    for (int v = 0; v < V; ++v)
      S += project(xyz, origin + 3 * b, vs, KR + ((int64_t)v * B + b) * 16, H, W).valid;
  }
  return S;
}

/* ------------------------------------------------------------------------------------------------
 * TSDF integration.
 * ---------------------------------------------------------------------------------------------- */

/* PTX cvt.rzi.s32.f32: NaN -> 0, saturating (what `(int)` means inside the reference CUDA kernel). */
static inline int cvt_rzi_s32(float x) {
  if (x != x) return 0;
  if (x >= 2147483648.0f) return 2147483647;
  if (x <= -2147483648.0f) return (-2147483647 - 1);
  return (int)x;
}

enum {
  ORC_TSDF_ROUND_HALF_AWAY = 0, /* CUDA roundf, TSDFVolume GPU kernel tsdf_volume.py:105-106 */
  ORC_TSDF_ROUND_HALF_EVEN = 1  /* np.round / torch.round, CPU paths :194-195, :458-459      */
};

/* One frame into one volume, the semantics of the reference GPU kernel (tsdf_volume.py:68-126) with
 * the FMA contraction nvcc applies to that exact source (checked against the PTX of the verbatim
 * string, see oracle/build_ref.py):
 *   pt   = fma(voxel, voxel_size, origin)                                   (:94-96)
 *   tmp  = pt - t ;  cam_k = fma(tmp_z,R[2][k], fma(tmp_x,R[0][k], tmp_y*R[1][k]))   (:98-103)
 *   px   = (int)roundf(fma(fx, cam_x/cam_z, cx))                              (:105-106)
 *   skip: px,py outside image, cam_z<0 (:110), depth==0 (:114), depth-cam_z < -trunc (:119)
 *   dist = fmin(1, diff/trunc); w_new = w_old+obs; tsdf = fma(dist,obs, w_old*tsdf)/w_new  (:121-126)
 * Deliberate deviation (documented in DESIGN.md): the linear index is decomposed with integer
 * arithmetic; the reference's float decomposition (:89-91) mis-addresses voxels once Nvox > 2^24.
 * `with_color` enables the colour running average of :130-141 (dead code in the reference kernel).
 * Volumes are (X,Y,Z) C-order float32.  color_im is the folded b*65536+g*256+r image (:223-227). */
int orc_tsdf_integrate(float* tsdf, float* weight, float* color, int dx, int dy, int dz,
                       const float* origin, float voxel_size, const float* intr9, const float* pose16,
                       const float* depth, const float* color_im, int im_h, int im_w, float trunc,
                       float obs_weight, int with_color) {
  const float fx = intr9[0], cx = intr9[2], fy = intr9[4], cy = intr9[5];
  const float* T = pose16;
#pragma omp parallel for schedule(static)
  for (int x = 0; x < dx; ++x) {
    for (int y = 0; y < dy; ++y) {
      for (int z = 0; z < dz; ++z) {
        int64_t idx = ((int64_t)x * dy + y) * dz + z;
        float ptx = fmaf((float)x, voxel_size, origin[0]);
        float pty = fmaf((float)y, voxel_size, origin[1]);
        float ptz = fmaf(voxel_size, (float)z, origin[2]);
        float tx = ptx - T[3], ty = pty - T[7], tz = ptz - T[11];
        float camx = fmaf(tz, T[8], fmaf(tx, T[0], ty * T[4]));
        float camy = fmaf(tz, T[9], fmaf(tx, T[1], ty * T[5]));
        float camz = fmaf(tz, T[10], fmaf(tx, T[2], ty * T[6]));
        int px = cvt_rzi_s32(roundf(fmaf(fx, camx / camz, cx)));
        int py = cvt_rzi_s32(roundf(fmaf(fy, camy / camz, cy)));
        if (px < 0 || px >= im_w || py < 0 || py >= im_h || camz < 0.0f) continue;
        float d = depth[(int64_t)py * im_w + px];
        if (d == 0.0f) continue;
        float diff = d - camz;
        if (diff < -trunc) continue;
        float dist = fminf(diff / trunc, 1.0f);
        float w_old = weight[idx];
        float w_new = w_old + obs_weight;
        weight[idx] = w_new;
        tsdf[idx] = fmaf(dist, obs_weight, w_old * tsdf[idx]) / w_new;
        if (with_color) {
          /* tsdf_volume.py:130-141 made reachable; contraction as nvcc applies it to that source */
          float oc = color[idx];
          float ob = floorf(oc / 65536.0f);
          float t0 = oc - ob * 65536.0f;                 /* integers < 2^24: every step exact */
          float og = floorf(t0 / 256.0f);
          float orr = t0 - og * 256.0f;
          float nc = color_im[(int64_t)py * im_w + px];
          float nb = floorf(nc / 65536.0f);
          float t1 = nc - nb * 65536.0f;
          float ng = floorf(t1 / 256.0f);
          float nr = t1 - ng * 256.0f;
          /* PTX of the reachable variant: mul(obs,new) then fma(w_old, old, .) then div.rn, roundf, min */
          nb = fminf(roundf(fmaf(w_old, ob, obs_weight * nb) / w_new), 255.0f);
          ng = fminf(roundf(fmaf(w_old, og, obs_weight * ng) / w_new), 255.0f);
          nr = fminf(roundf(fmaf(w_old, orr, obs_weight * nr) / w_new), 255.0f);
          color[idx] = (nb * 65536.0f + ng * 256.0f) + nr;
        }
      }
    }
  }
  return 0;
}

/* TSDFVolumeTorch.integrate semantics (tsdf_volume.py:437-482):
 *   world_c = origin + voxel_size*coords                 (:523, fp32)
 *   cam_c   = inverse(cam_pose) @ [world_c;1]             (:451-452; `w2c` is passed in, 3x4 rows of
 *             the fp32 inverse; K=4 product evaluated as the same FMA chain as torch CPU matmul)
 *   pix     = round_half_even(cam_x*fx/cam_z + cx)        (:458-459, mul, div, add: three roundings)
 *   valid   : 0<=pix<size & cam_z>0 (:462);  depth>0 & depth-cam_z >= -trunc (:471)
 *   dist = min(diff/trunc,1); tsdf = (w_old*tsdf + obs*dist)/w_new  (:470-480, no contraction)  */
int orc_tsdf_integrate_torch(float* tsdf, float* weight, int dx, int dy, int dz, const float* origin,
                             float voxel_size, const float* intr9, const float* w2c12,
                             const float* depth, int im_h, int im_w, float trunc, float obs_weight) {
  const float fx = intr9[0], cx = intr9[2], fy = intr9[4], cy = intr9[5];
#pragma omp parallel for schedule(static)
  for (int x = 0; x < dx; ++x) {
    for (int y = 0; y < dy; ++y) {
      for (int z = 0; z < dz; ++z) {
        int64_t idx = ((int64_t)x * dy + y) * dz + z;
        float wx = voxel_size * (float)x; wx = origin[0] + wx;
        float wy = voxel_size * (float)y; wy = origin[1] + wy;
        float wz = voxel_size * (float)z; wz = origin[2] + wz;
        float cam[3];
        for (int r = 0; r < 3; ++r) {
          float a = w2c12[4 * r + 0] * wx;
          a = fmaf(w2c12[4 * r + 1], wy, a);
          a = fmaf(w2c12[4 * r + 2], wz, a);
          a = fmaf(w2c12[4 * r + 3], 1.0f, a);
          cam[r] = a;
        }
        float ux = cam[0] * fx; ux = ux / cam[2]; ux = ux + cx;
        float uy = cam[1] * fy; uy = uy / cam[2]; uy = uy + cy;
        float rx = nearbyintf(ux), ry = nearbyintf(uy); /* default rounding mode = half to even */
        if (!(rx >= 0.0f && rx < (float)im_w && ry >= 0.0f && ry < (float)im_h && cam[2] > 0.0f)) continue;
        int px = (int)rx, py = (int)ry;
        float d = depth[(int64_t)py * im_w + px];
        float diff = d - cam[2];
        if (!(d > 0.0f && diff >= -trunc)) continue;
        float dist = diff / trunc; if (dist > 1.0f) dist = 1.0f;
        float w_old = weight[idx];
        float w_new = w_old + obs_weight;
        float a = w_old * tsdf[idx];
        float b = obs_weight * dist;
        tsdf[idx] = (a + b) / w_new;
        weight[idx] = w_new;
      }
    }
  }
  return 0;
}
