// MultiScaleRoIAlign forward / backward — generic kernels.
//
// Replaces TV ops/poolers.py:147-227 (zeros + per level: where / gather / roi_align /
// scatter) by ONE launch over all RoIs of all levels: the FPN level is mapped in-kernel
// (roi_common.cuh), so there is no host sync, no per-level gather and no result scatter.
// These kernels handle every configuration (any pooled size, sampling ratio <= 0 = adaptive,
// aligned flag, NCHW or NHWC, fp32 or bf16); roi_align_fast.cu holds the tuned 7x7/sr=2 paths.
//
// Forward: one CTA per RoI.  The per-axis taps (low/high index, weights, in-range flag) of
// all PH*grid_h row samples and PW*grid_w column samples are tabulated once in shared memory —
// bilinear weights are separable — then threads sweep the [C, PH, PW] output contiguously so
// stores coalesce, and the 4-tap gathers of neighbouring bins hit the same L1 lines.
// Backward (algo 1): the same sweep scattering with fp32 atomics into zero-filled gradients.
#include "roi_common.cuh"
#include <stdlib.h>

namespace dgod {

constexpr int kRoiThreads = 256;
constexpr int kMaxTabSamples = 64;  // PH*grid_h and PW*grid_w up to this use the smem tables

template <typename T, bool NHWC>
__device__ __forceinline__ float feat_at(const T* __restrict__ base, int c, int C, int W, int y, int x) {
  // base points at the image: NCHW -> [C][H][W] plane set; NHWC -> [H][W][C]
  if (NHWC) return to_f32<T>(base[((size_t)y * W + x) * C + c]);
  return to_f32<T>(base[x + (size_t)W * y]);  // caller pre-offsets the channel plane
}

template <typename T, bool NHWC>
__global__ void __launch_bounds__(kRoiThreads)
msroi_fwd_generic_kernel(const RoiDev g, const float* __restrict__ rois, int n_rois,
                         T* __restrict__ out) {
  __shared__ AxisTap ytab[kMaxTabSamples], xtab[kMaxTabSamples];
  __shared__ RoiGeom s_geo;
  const int k = blockIdx.x;
  if (threadIdx.x == 0) s_geo = roi_geometry(g, rois + (size_t)k * 5);
  __syncthreads();
  const RoiGeom r = s_geo;
  const int PH = g.PH, PW = g.PW, C = g.C;
  const int n_out = C * PH * PW;
  T* __restrict__ o = out + (size_t)k * n_out;
  if (r.batch < 0 || r.batch >= g.B) {
    for (int i = threadIdx.x; i < n_out; i += blockDim.x) o[i] = from_f32<T>(0.f);
    return;
  }
  const int ny = PH * r.grid_h, nx = PW * r.grid_w;
  const bool tab = ny <= kMaxTabSamples && nx <= kMaxTabSamples;
  if (tab) {
    for (int i = threadIdx.x; i < ny; i += blockDim.x)
      ytab[i] = axis_tap(sample_coord(r.start_h, i / r.grid_h, r.bin_h, i % r.grid_h, r.grid_h), r.H);
    for (int i = threadIdx.x; i < nx; i += blockDim.x)
      xtab[i] = axis_tap(sample_coord(r.start_w, i / r.grid_w, r.bin_w, i % r.grid_w, r.grid_w), r.W);
  }
  __syncthreads();
  const T* __restrict__ img = reinterpret_cast<const T*>(g.feat[r.level]) + (size_t)r.batch * C * r.H * r.W;
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) {
    const int pw = i % PW, ph = (i / PW) % PH, c = i / (PW * PH);
    const T* __restrict__ plane = NHWC ? img : img + (size_t)c * r.H * r.W;
    float acc = 0.f;
    for (int iy = 0; iy < r.grid_h; ++iy) {
      const AxisTap ty = tab ? ytab[ph * r.grid_h + iy]
                             : axis_tap(sample_coord(r.start_h, ph, r.bin_h, iy, r.grid_h), r.H);
      for (int ix = 0; ix < r.grid_w; ++ix) {
        const AxisTap tx = tab ? xtab[pw * r.grid_w + ix]
                               : axis_tap(sample_coord(r.start_w, pw, r.bin_w, ix, r.grid_w), r.W);
        if (!(ty.valid && tx.valid)) continue;  // the CPU kernel's empty PreCalc entry
        const float v1 = feat_at<T, NHWC>(plane, c, C, r.W, ty.lo, tx.lo);
        const float v2 = feat_at<T, NHWC>(plane, c, C, r.W, ty.lo, tx.hi);
        const float v3 = feat_at<T, NHWC>(plane, c, C, r.W, ty.hi, tx.lo);
        const float v4 = feat_at<T, NHWC>(plane, c, C, r.W, ty.hi, tx.hi);
        const float w1 = __fmul_rn(ty.h, tx.h), w2 = __fmul_rn(ty.h, tx.l);
        const float w3 = __fmul_rn(ty.l, tx.h), w4 = __fmul_rn(ty.l, tx.l);
        const float s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1), __fmul_rn(w2, v2)),
                                            __fmul_rn(w3, v3)), __fmul_rn(w4, v4));
        acc = __fadd_rn(acc, s);
      }
    }
    o[i] = from_f32<T>(__fdiv_rn(acc, r.count));
  }
}

// Backward, algorithm 1: atomic scatter (fp32 gradients only).
template <bool NHWC>
__global__ void __launch_bounds__(kRoiThreads)
msroi_bwd_atomic_kernel(const RoiDev g, const float* __restrict__ grad_out,
                        const float* __restrict__ rois, int n_rois) {
  __shared__ AxisTap ytab[kMaxTabSamples], xtab[kMaxTabSamples];
  __shared__ RoiGeom s_geo;
  const int k = blockIdx.x;
  if (threadIdx.x == 0) s_geo = roi_geometry(g, rois + (size_t)k * 5);
  __syncthreads();
  const RoiGeom r = s_geo;
  if (r.batch < 0 || r.batch >= g.B) return;
  const int PH = g.PH, PW = g.PW, C = g.C;
  const int n_out = C * PH * PW;
  const int ny = PH * r.grid_h, nx = PW * r.grid_w;
  const bool tab = ny <= kMaxTabSamples && nx <= kMaxTabSamples;
  if (tab) {
    for (int i = threadIdx.x; i < ny; i += blockDim.x)
      ytab[i] = axis_tap(sample_coord(r.start_h, i / r.grid_h, r.bin_h, i % r.grid_h, r.grid_h), r.H);
    for (int i = threadIdx.x; i < nx; i += blockDim.x)
      xtab[i] = axis_tap(sample_coord(r.start_w, i / r.grid_w, r.bin_w, i % r.grid_w, r.grid_w), r.W);
  }
  __syncthreads();
  const float cnt = (float)(r.grid_h * r.grid_w);  // backward divides by the raw product
  float* __restrict__ img = reinterpret_cast<float*>(g.gfeat[r.level]) + (size_t)r.batch * C * r.H * r.W;
  const float* __restrict__ go = grad_out + (size_t)k * n_out;
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) {
    const int pw = i % PW, ph = (i / PW) % PH, c = i / (PW * PH);
    const float gv = go[i];
    float* __restrict__ plane = NHWC ? img : img + (size_t)c * r.H * r.W;
    for (int iy = 0; iy < r.grid_h; ++iy) {
      const AxisTap ty = tab ? ytab[ph * r.grid_h + iy]
                             : axis_tap(sample_coord(r.start_h, ph, r.bin_h, iy, r.grid_h), r.H);
      for (int ix = 0; ix < r.grid_w; ++ix) {
        const AxisTap tx = tab ? xtab[pw * r.grid_w + ix]
                               : axis_tap(sample_coord(r.start_w, pw, r.bin_w, ix, r.grid_w), r.W);
        if (!(ty.valid && tx.valid)) continue;
        const float g1 = __fdiv_rn(__fmul_rn(gv, __fmul_rn(ty.h, tx.h)), cnt);
        const float g2 = __fdiv_rn(__fmul_rn(gv, __fmul_rn(ty.h, tx.l)), cnt);
        const float g3 = __fdiv_rn(__fmul_rn(gv, __fmul_rn(ty.l, tx.h)), cnt);
        const float g4 = __fdiv_rn(__fmul_rn(gv, __fmul_rn(ty.l, tx.l)), cnt);
        if (NHWC) {
          atomicAdd(plane + ((size_t)ty.lo * r.W + tx.lo) * C + c, g1);
          atomicAdd(plane + ((size_t)ty.lo * r.W + tx.hi) * C + c, g2);
          atomicAdd(plane + ((size_t)ty.hi * r.W + tx.lo) * C + c, g3);
          atomicAdd(plane + ((size_t)ty.hi * r.W + tx.hi) * C + c, g4);
        } else {
          atomicAdd(plane + (size_t)ty.lo * r.W + tx.lo, g1);
          atomicAdd(plane + (size_t)ty.lo * r.W + tx.hi, g2);
          atomicAdd(plane + (size_t)ty.hi * r.W + tx.lo, g3);
          atomicAdd(plane + (size_t)ty.hi * r.W + tx.hi, g4);
        }
      }
    }
  }
}

int fill_roi_dev(const dgod_roi_config* cfg, RoiDev& g) {
  DGOD_REQUIRE(cfg, "roi_align: cfg is null");
  DGOD_REQUIRE(cfg->n_levels >= 1 && cfg->n_levels <= DGOD_MAX_LEVELS, "roi_align: n_levels out of range");
  DGOD_REQUIRE(cfg->batch >= 0 && cfg->channels > 0, "roi_align: bad batch/channels");
  DGOD_REQUIRE(cfg->pooled_h > 0 && cfg->pooled_w > 0, "roi_align: pooled size must be positive");
  DGOD_REQUIRE(cfg->dtype == DGOD_F32 || cfg->dtype == DGOD_BF16, "roi_align: unsupported dtype");
  DGOD_REQUIRE(cfg->n_levels == 1 || cfg->k_max - cfg->k_min + 1 == cfg->n_levels,
               "roi_align: k_min/k_max do not match n_levels");
  g.n_levels = cfg->n_levels; g.B = cfg->batch; g.C = cfg->channels;
  g.PH = cfg->pooled_h; g.PW = cfg->pooled_w; g.sr = cfg->sampling_ratio; g.aligned = cfg->aligned;
  g.k_min = cfg->k_min; g.k_max = cfg->k_max; g.channels_last = cfg->channels_last;
  g.s0 = cfg->canonical_scale; g.lvl0 = cfg->canonical_level; g.eps = cfg->eps;
  for (int l = 0; l < DGOD_MAX_LEVELS; ++l) {
    g.feat[l] = nullptr; g.gfeat[l] = nullptr; g.H[l] = g.W[l] = 0; g.scale[l] = 0.f;
  }
  for (int l = 0; l < cfg->n_levels; ++l) {
    DGOD_REQUIRE(cfg->height[l] > 0 && cfg->width[l] > 0, "roi_align: empty feature level");
    g.H[l] = cfg->height[l]; g.W[l] = cfg->width[l]; g.scale[l] = cfg->spatial_scale[l];
  }
  return DGOD_OK;
}

// tuned paths (roi_align_fast.cu); return DGOD_OK + *handled = 1 when they took the call
int msroi_fwd_fast(const dgod_roi_config* cfg, const RoiDev& g, const float* rois, int n_rois,
                   void* out, cudaStream_t st, int* handled);
int msroi_bwd_fast(const dgod_roi_config* cfg, const RoiDev& g, const void* grad_out,
                   const float* rois, int n_rois, const int32_t* roi_img_offsets, void* workspace,
                   size_t workspace_bytes, cudaStream_t st, int* handled);
size_t msroi_bwd_workspace(int n_rois);
int msroi_bwd_red(const dgod_roi_config* cfg, const RoiDev& g, const void* grad_out, const float* rois,
                  int n_rois, cudaStream_t st, int* handled);
// TMA paths (roi_align_tma.cu): channels_last, 7x7 bins, sampling ratio 1..2
int msroi_fwd_tma(const dgod_roi_config* cfg, const RoiDev& g, const float* rois, int n_rois, void* out, void* workspace,
                  size_t workspace_bytes, cudaStream_t st, int* handled);
int msroi_bwd_tma(const dgod_roi_config* cfg, const RoiDev& g, const void* grad_out, const float* rois, int n_rois,
                  void* workspace, size_t workspace_bytes, cudaStream_t st, int* handled);
size_t msroi_tma_workspace(int n_rois);
// owner-computes backward (roi_align_own.cu): channels_last, 7x7 bins, sampling ratio 1..2, C % 64 == 0
int msroi_bwd_own(const dgod_roi_config* cfg, const RoiDev& g, const void* grad_out, const float* rois, int n_rois,
                  const int32_t* roi_img_offsets, void* workspace, size_t workspace_bytes, cudaStream_t st, int* handled, int claim);
size_t msroi_own_workspace(const RoiDev& g, int n_rois);

// DGOD_FWD_ALGO=1 keeps the forward on the table-driven kernel (A/B measurements); default: TMA path first.
static int fwd_algo_from_env() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DGOD_FWD_ALGO");
    v = e ? atoi(e) : 0;
  }
  return v;
}

}  // namespace dgod

using namespace dgod;

extern "C" size_t dgod_msroi_align_fwd_workspace_bytes(int n_rois) { return msroi_tma_workspace(n_rois); }

extern "C" int dgod_msroi_align_fwd(const dgod_roi_config* cfg, const void* const* feats,
                                    const float* rois, int n_rois, void* out, void* workspace,
                                    size_t workspace_bytes, dgod_stream_t stream) {
  RoiDev g;
  int rc = fill_roi_dev(cfg, g);
  if (rc) return rc;
  DGOD_REQUIRE(n_rois >= 0, "roi_align: negative n_rois");
  if (n_rois == 0) return DGOD_OK;
  DGOD_REQUIRE(feats && rois && out, "roi_align: null pointer");
  for (int l = 0; l < g.n_levels; ++l) {
    DGOD_REQUIRE(feats[l] || g.B == 0, "roi_align: null feature pointer");
    g.feat[l] = feats[l];
  }
  cudaStream_t st = (cudaStream_t)stream;
  int handled = 0;
  if (fwd_algo_from_env() == 0) {
    rc = msroi_fwd_tma(cfg, g, rois, n_rois, out, workspace, workspace_bytes, st, &handled);
    if (rc || handled) return rc;
  }
  rc = msroi_fwd_fast(cfg, g, rois, n_rois, out, st, &handled);
  if (rc || handled) return rc;
  if (cfg->dtype == DGOD_F32) {
    if (g.channels_last) msroi_fwd_generic_kernel<float, true><<<n_rois, kRoiThreads, 0, st>>>(g, rois, n_rois, (float*)out);
    else msroi_fwd_generic_kernel<float, false><<<n_rois, kRoiThreads, 0, st>>>(g, rois, n_rois, (float*)out);
  } else {
    if (g.channels_last) msroi_fwd_generic_kernel<__nv_bfloat16, true><<<n_rois, kRoiThreads, 0, st>>>(g, rois, n_rois, (__nv_bfloat16*)out);
    else msroi_fwd_generic_kernel<__nv_bfloat16, false><<<n_rois, kRoiThreads, 0, st>>>(g, rois, n_rois, (__nv_bfloat16*)out);
  }
  DGOD_LAUNCHED();
  return DGOD_OK;
}

extern "C" size_t dgod_msroi_align_bwd_workspace_bytes(int n_rois) {
  const size_t a = msroi_bwd_workspace(n_rois), b = msroi_tma_workspace(n_rois);
  return a > b ? a : b;
}

extern "C" size_t dgod_msroi_align_bwd_workspace_bytes_cfg(const dgod_roi_config* cfg, int n_rois) {
  size_t need = dgod_msroi_align_bwd_workspace_bytes(n_rois);
  RoiDev g;
  if (cfg && fill_roi_dev(cfg, g) == DGOD_OK) {
    const size_t own = msroi_own_workspace(g, n_rois);
    if (own > need) need = own;
  }
  return need;
}

extern "C" int dgod_msroi_align_bwd(const dgod_roi_config* cfg, const void* grad_out,
                                    const float* rois, int n_rois,
                                    const int32_t* roi_img_offsets, void* const* grad_feats,
                                    int algo, void* workspace, size_t workspace_bytes,
                                    dgod_stream_t stream) {
  RoiDev g;
  int rc = fill_roi_dev(cfg, g);
  if (rc) return rc;
  DGOD_REQUIRE(n_rois >= 0, "roi_align: negative n_rois");
  DGOD_REQUIRE(algo >= 0 && algo <= 5, "roi_align: unknown backward algorithm");
  DGOD_REQUIRE(grad_feats, "roi_align: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t esz = cfg->dtype == DGOD_F32 ? 4 : 2;
  for (int l = 0; l < g.n_levels; ++l) {
    DGOD_REQUIRE(grad_feats[l] || g.B == 0, "roi_align: null gradient pointer");
    g.gfeat[l] = grad_feats[l];
  }
  if (g.B == 0) return DGOD_OK;
  DGOD_REQUIRE(n_rois == 0 || (grad_out && rois), "roi_align: null pointer");
  // algo 0 (auto): the owner-computes kernel when the shape allows (channels_last, 7x7, sr 1..2, C % 64 == 0) and the
  // workspace was sized with dgod_msroi_align_bwd_workspace_bytes_cfg; else the TMA bulk-reduce path; else
  // fp32 gradients take the scatter path (16-byte vector reductions on channels_last, scalar atomics
  // on NCHW) and bf16 gradients the deterministic tile gather (fp32 accumulation, one rounding).
  if (algo == 0 || algo == 4 || algo == 5) {
    // 4: work items dealt round-robin (the default); 5: claimed from a global counter — for steps in which other
    // streams' kernels (an NCCL all-reduce overlapping backward) hold SMs when this persistent grid starts
    int handled = 0;
    rc = msroi_bwd_own(cfg, g, grad_out, rois, n_rois, roi_img_offsets, workspace, workspace_bytes, st, &handled, algo == 5);
    if (rc || handled) return rc;
    DGOD_REQUIRE(algo == 0, "roi_align: the owner-computes backward does not support this configuration or workspace");
  }
  if (algo == 0 || algo == 3) {
    int handled = 0;
    rc = msroi_bwd_tma(cfg, g, grad_out, rois, n_rois, workspace, workspace_bytes, st, &handled);
    if (rc || handled) return rc;
    DGOD_REQUIRE(algo == 0, "roi_align: the TMA backward does not support this configuration");
  }
  const bool want_tile = algo == 2 || (algo == 0 && cfg->dtype != DGOD_F32);
  if (want_tile && n_rois > 0) {
    int handled = 0;
    rc = msroi_bwd_fast(cfg, g, grad_out, rois, n_rois, roi_img_offsets, workspace, workspace_bytes, st, &handled);
    if (rc || handled) return rc;
    DGOD_REQUIRE(algo == 0, "roi_align: the tile-gather backward does not support this configuration");
  }
  for (int l = 0; l < g.n_levels; ++l)
    DGOD_CUDA(cudaMemsetAsync(g.gfeat[l], 0, (size_t)g.B * g.C * g.H[l] * g.W[l] * esz, st));
  if (n_rois == 0) return DGOD_OK;
  DGOD_REQUIRE(cfg->dtype == DGOD_F32, "roi_align: the atomic backward supports fp32 gradients only");
  {
    int handled = 0;
    rc = msroi_bwd_red(cfg, g, grad_out, rois, n_rois, st, &handled);   // NHWC: 16-byte vector reductions
    if (rc || handled) return rc;
  }
  if (g.channels_last) msroi_bwd_atomic_kernel<true><<<n_rois, kRoiThreads, 0, st>>>(g, (const float*)grad_out, rois, n_rois);
  else msroi_bwd_atomic_kernel<false><<<n_rois, kRoiThreads, 0, st>>>(g, (const float*)grad_out, rois, n_rois);
  DGOD_LAUNCHED();
  return DGOD_OK;
}
