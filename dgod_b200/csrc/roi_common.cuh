// Per-RoI geometry shared by the MultiScaleRoIAlign kernels.
//
// Restates torchvision's CPU roi_align kernel (the op behind TV ops/roi_align.py:204-260; its
// Python transcription is TV ops/roi_align.py:115-200, which lacks the out-of-range sample
// skip — SURVEY.md §8c) and the FPN level heuristic of TV ops/poolers.py:47-84.
// Operation order and the absence of FMA contraction follow the CPU kernel so that fp32
// results agree to the last bit wherever the accumulation order is the same.
#pragma once
#include "common.cuh"

namespace dgod {

struct RoiDev {
  const void* feat[DGOD_MAX_LEVELS];
  void* gfeat[DGOD_MAX_LEVELS];
  int n_levels, B, C, PH, PW, sr, aligned, k_min, k_max, channels_last;
  int H[DGOD_MAX_LEVELS], W[DGOD_MAX_LEVELS];
  float scale[DGOD_MAX_LEVELS];
  float s0, lvl0, eps;
};

struct RoiGeom {
  int level, batch;      // batch < 0 or >= B marks an unusable RoI (output zeros)
  int H, W;
  float start_w, start_h, bin_w, bin_h;
  int grid_h, grid_w;
  float count;
};

// LevelMapper.__call__ (TV ops/poolers.py:73-84):
//   floor(lvl0 + log2(sqrt(area) / s0) + eps) clamped to [k_min, k_max], minus k_min.
__device__ __forceinline__ int map_level(const RoiDev& g, float x1, float y1, float x2, float y2) {
  if (g.n_levels == 1) return 0;
  const float area = box_area_exact(x1, y1, x2, y2);
  const float s = __fsqrt_rn(area);
  const float q = __fdiv_rn(s, g.s0);
  const float lg = (float)log2((double)q);  // exact for powers of two, correctly rounded otherwise
  float t = floorf(__fadd_rn(__fadd_rn(g.lvl0, lg), g.eps));
  t = fminf(fmaxf(t, (float)g.k_min), (float)g.k_max);  // NaN -> k_min
  if (!(t == t)) t = (float)g.k_min;
  return (int)t - g.k_min;
}

__device__ __forceinline__ RoiGeom roi_geometry(const RoiDev& g, const float* __restrict__ roi) {
  RoiGeom r;
  const float bidx = roi[0];
  const float x1 = roi[1], y1 = roi[2], x2 = roi[3], y2 = roi[4];
  r.batch = (int)bidx;
  r.level = map_level(g, x1, y1, x2, y2);
  r.H = g.H[r.level];
  r.W = g.W[r.level];
  const float sc = g.scale[r.level];
  const float off = g.aligned ? 0.5f : 0.f;
  r.start_w = __fsub_rn(__fmul_rn(x1, sc), off);
  r.start_h = __fsub_rn(__fmul_rn(y1, sc), off);
  const float end_w = __fsub_rn(__fmul_rn(x2, sc), off);
  const float end_h = __fsub_rn(__fmul_rn(y2, sc), off);
  float rw = __fsub_rn(end_w, r.start_w), rh = __fsub_rn(end_h, r.start_h);
  if (!g.aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
  r.bin_h = __fdiv_rn(rh, (float)g.PH);
  r.bin_w = __fdiv_rn(rw, (float)g.PW);
  r.grid_h = g.sr > 0 ? g.sr : (int)ceilf(__fdiv_rn(rh, (float)g.PH));
  r.grid_w = g.sr > 0 ? g.sr : (int)ceilf(__fdiv_rn(rw, (float)g.PW));
  r.count = fmaxf((float)(r.grid_h * r.grid_w), 1.f);
  return r;
}

// One axis of bilinear_interpolate / pre_calc_for_bilinear_interpolate: sample coordinate ->
// (low index, high index, low weight l, high weight h, in-range flag).
struct AxisTap {
  int lo, hi;
  float l, h;  // value = h * v[lo] + l * v[hi]
  int valid;
};

__device__ __forceinline__ float sample_coord(float start, int p, float bin, int i, int grid) {
  // start + p*bin + (i + .5f) * bin / grid, left to right as in the CPU kernel
  return __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)),
                   __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)grid));
}

__device__ __forceinline__ AxisTap axis_tap(float v, int size) {
  AxisTap t;
  t.valid = !(v < -1.0f || v > (float)size);
  if (v <= 0.f) v = 0.f;
  int lo = (int)v, hi;
  if (lo >= size - 1) { hi = lo = size - 1; v = (float)lo; } else { hi = lo + 1; }
  t.lo = lo; t.hi = hi;
  t.l = __fsub_rn(v, (float)lo);
  t.h = __fsub_rn(1.f, t.l);
  return t;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

}  // namespace dgod
