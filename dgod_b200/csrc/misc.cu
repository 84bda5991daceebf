// Small element-wise kernels of the hot path: gradient-reversal scale (A13), BoxCoder.decode
// (A1, box-head weights) and the candidate stage of postprocess_detections (A9).
//
// Reference behaviour restated:
//   GRLayer.backward          DGcommon.py:40-42        out = grad.neg() * 0.1
//   BoxCoder.decode_single    TV models/detection/_utils.py:186-224
//   postprocess_detections    TV models/detection/roi_heads.py:692-724 (softmax, clip, drop the
//                             background column, score > thresh, remove_small_boxes(1e-2))
// All three are streaming kernels: every byte is read or written exactly once, 128-bit accesses.
#include "common.cuh"

namespace dgod {

// ---------------------------------------------------------------------------- GRL
__global__ void __launch_bounds__(256)
grl_scale_f32_kernel(const float* __restrict__ g, float* __restrict__ out, long long n, float alpha) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n4 = n >> 2;
  const float4* __restrict__ g4 = reinterpret_cast<const float4*>(g);
  float4* __restrict__ o4 = reinterpret_cast<float4*>(out);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = g4[i];
    v.x = __fmul_rn(-v.x, alpha); v.y = __fmul_rn(-v.y, alpha);
    v.z = __fmul_rn(-v.z, alpha); v.w = __fmul_rn(-v.w, alpha);
    o4[i] = v;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __fmul_rn(-g[i], alpha);
}

__global__ void __launch_bounds__(256)
grl_scale_bf16_kernel(const __nv_bfloat16* __restrict__ g, __nv_bfloat16* __restrict__ out,
                      long long n, float alpha) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n8 = n >> 3;
  const uint4* __restrict__ g8 = reinterpret_cast<const uint4*>(g);
  uint4* __restrict__ o8 = reinterpret_cast<uint4*>(out);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    uint4 v = g8[i];
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __bfloat1622float2(h[j]);
      h[j] = __floats2bfloat162_rn(__fmul_rn(-f.x, alpha), __fmul_rn(-f.y, alpha));
    }
    o8[i] = v;
  }
  for (long long i = (n8 << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __float2bfloat16_rn(__fmul_rn(-__bfloat162float(g[i]), alpha));
}

__global__ void __launch_bounds__(256)
grl_scale_scalar_kernel(const void* __restrict__ g, void* __restrict__ out, long long n, float alpha,
                        int dtype) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (dtype == DGOD_F32)
      ((float*)out)[i] = __fmul_rn(-((const float*)g)[i], alpha);
    else
      ((__nv_bfloat16*)out)[i] =
          __float2bfloat16_rn(__fmul_rn(-__bfloat162float(((const __nv_bfloat16*)g)[i]), alpha));
  }
}

// ---------------------------------------------------------------------------- decode
struct DecodeW { float wx, wy, ww, wh, clip; };

__device__ __forceinline__ float4 decode_one(const float4 d, const float4 b, const DecodeW w) {
  const float bw = __fsub_rn(b.z, b.x), bh = __fsub_rn(b.w, b.y);
  const float cx = __fadd_rn(b.x, __fmul_rn(0.5f, bw)), cy = __fadd_rn(b.y, __fmul_rn(0.5f, bh));
  const float dx = __fdiv_rn(d.x, w.wx), dy = __fdiv_rn(d.y, w.wy);
  const float dw = fminf(__fdiv_rn(d.z, w.ww), w.clip), dh = fminf(__fdiv_rn(d.w, w.wh), w.clip);
  const float pcx = __fadd_rn(__fmul_rn(dx, bw), cx), pcy = __fadd_rn(__fmul_rn(dy, bh), cy);
  const float pw = __fmul_rn((float)exp((double)dw), bw), ph = __fmul_rn((float)exp((double)dh), bh);
  const float hw = __fmul_rn(0.5f, pw), hh = __fmul_rn(0.5f, ph);
  return make_float4(__fsub_rn(pcx, hw), __fsub_rn(pcy, hh), __fadd_rn(pcx, hw), __fadd_rn(pcy, hh));
}

__global__ void __launch_bounds__(256)
box_decode_kernel(const float* __restrict__ rel, const float* __restrict__ boxes, int n, int n_cls,
                  DecodeW w, float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * n_cls) return;
  const int row = (int)(t / n_cls);
  const float4 d = __ldg(reinterpret_cast<const float4*>(rel) + t);
  reinterpret_cast<float4*>(out)[t] = decode_one(d, ld_box(boxes, row), w);
}

// ---------------------------------------------------------------------------- detection candidates
__global__ void __launch_bounds__(128)
detect_candidates_kernel(const float* __restrict__ logits, const float* __restrict__ reg,
                         const float* __restrict__ proposals, const int32_t* __restrict__ box_offsets,
                         const float* __restrict__ image_sizes, int n_img, int n_rows, int n_cls,
                         DecodeW w, float score_thresh, float min_size,
                         float* __restrict__ cand_boxes, float* __restrict__ cand_scores,
                         int64_t* __restrict__ cand_labels, uint8_t* __restrict__ cand_valid) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  int img = 0;
  {
    int lo = 0, hi = n_img;  // largest i with box_offsets[i] <= row
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (box_offsets[mid] <= row) lo = mid; else hi = mid;
    }
    img = lo;
  }
  const float img_h = image_sizes[2 * img], img_w = image_sizes[2 * img + 1];
  const float* __restrict__ lg = logits + (size_t)row * n_cls;
  float m = lg[0];
  for (int c = 1; c < n_cls; ++c) m = fmaxf(m, lg[c]);
  float sum = 0.f;
  for (int c = 0; c < n_cls; ++c) sum = __fadd_rn(sum, (float)exp((double)__fsub_rn(lg[c], m)));
  const float4 pb = ld_box(proposals, row);
  for (int c = 1; c < n_cls; ++c) {
    const float score = __fdiv_rn((float)exp((double)__fsub_rn(lg[c], m)), sum);
    float4 b = decode_one(__ldg(reinterpret_cast<const float4*>(reg) + (size_t)row * n_cls + c), pb, w);
    b.x = fminf(fmaxf(b.x, 0.f), img_w); b.z = fminf(fmaxf(b.z, 0.f), img_w);
    b.y = fminf(fmaxf(b.y, 0.f), img_h); b.w = fminf(fmaxf(b.w, 0.f), img_h);
    const bool ok = score > score_thresh && __fsub_rn(b.z, b.x) >= min_size &&
                    __fsub_rn(b.w, b.y) >= min_size;
    const size_t o = (size_t)row * (n_cls - 1) + (c - 1);
    reinterpret_cast<float4*>(cand_boxes)[o] = b;
    cand_scores[o] = score;
    cand_labels[o] = c;
    cand_valid[o] = ok ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------- NCHW -> NHWC
// [B][C][HW] -> [B][HW][C] through a 64x64 shared-memory tile (+1 padding: conflict-free both
// ways).  Reads are contiguous along HW, writes contiguous along C: pure HBM traffic, 2*n*s bytes.
// Lets NCHW feature maps (torchvision's default layout) take the channels_last TMA RoIAlign path.
template <typename T>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const T* __restrict__ src, T* __restrict__ dst, int C, int HW) {
  __shared__ T tile[64][65];
  const int b = blockIdx.z;
  const int hw0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const T* __restrict__ s = src + (size_t)b * C * HW;
  T* __restrict__ d = dst + (size_t)b * C * HW;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;       // 64 x 4
#pragma unroll 4
  for (int i = ty; i < 64; i += 4) {
    const int c = c0 + i, hw = hw0 + tx;
    if (c < C && hw < HW) tile[i][tx] = s[(size_t)c * HW + hw];
  }
  __syncthreads();
#pragma unroll 4
  for (int i = ty; i < 64; i += 4) {
    const int hw = hw0 + i, c = c0 + tx;
    if (c < C && hw < HW) d[(size_t)hw * C + c] = tile[tx][i];
  }
}

}  // namespace dgod

using namespace dgod;

extern "C" int dgod_nchw_to_nhwc(const void* src, void* dst, int batch, int channels, int hw, int dtype,
                                 dgod_stream_t stream) {
  DGOD_REQUIRE(batch >= 0 && channels >= 0 && hw >= 0, "dgod_nchw_to_nhwc: negative size");
  DGOD_REQUIRE(dtype == DGOD_F32 || dtype == DGOD_BF16, "dgod_nchw_to_nhwc: unsupported dtype");
  if (batch == 0 || channels == 0 || hw == 0) return DGOD_OK;
  DGOD_REQUIRE(src && dst, "dgod_nchw_to_nhwc: null pointer");
  DGOD_REQUIRE(batch <= 65535 && cdiv(channels, 64) <= 65535, "dgod_nchw_to_nhwc: grid too large");
  dim3 grid(cdiv(hw, 64), cdiv(channels, 64), batch);
  if (dtype == DGOD_F32)
    nchw_to_nhwc_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)src, (float*)dst, channels, hw);
  else
    nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst,
                                                                           channels, hw);
  DGOD_LAUNCHED();
  return DGOD_OK;
}

extern "C" int dgod_grl_scale(const void* grad, void* out, int64_t n, float alpha, int dtype,
                              dgod_stream_t stream) {
  DGOD_REQUIRE(n >= 0, "dgod_grl_scale: negative size");
  DGOD_REQUIRE(dtype == DGOD_F32 || dtype == DGOD_BF16, "dgod_grl_scale: unsupported dtype");
  if (n == 0) return DGOD_OK;
  DGOD_REQUIRE(grad && out, "dgod_grl_scale: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const bool aligned = (((uintptr_t)grad | (uintptr_t)out) & 15u) == 0;
  const long long vec = dtype == DGOD_F32 ? 4 : 8;
  long long blocks = (n / vec + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;  // grid-stride: 16 CTAs per SM
  if (!aligned)
    grl_scale_scalar_kernel<<<(int)blocks, 256, 0, st>>>(grad, out, n, alpha, dtype);
  else if (dtype == DGOD_F32)
    grl_scale_f32_kernel<<<(int)blocks, 256, 0, st>>>((const float*)grad, (float*)out, n, alpha);
  else
    grl_scale_bf16_kernel<<<(int)blocks, 256, 0, st>>>((const __nv_bfloat16*)grad, (__nv_bfloat16*)out, n, alpha);
  DGOD_LAUNCHED();
  return DGOD_OK;
}

extern "C" int dgod_box_decode(const float* rel_codes, const float* boxes, int n, int n_cls,
                               float wx, float wy, float ww, float wh, float xform_clip,
                               float* out, dgod_stream_t stream) {
  DGOD_REQUIRE(n >= 0 && n_cls >= 0, "dgod_box_decode: negative size");
  if (n == 0 || n_cls == 0) return DGOD_OK;
  DGOD_REQUIRE(rel_codes && boxes && out, "dgod_box_decode: null pointer");
  DecodeW w{wx, wy, ww, wh, xform_clip};
  box_decode_kernel<<<cdiv((long long)n * n_cls, 256), 256, 0, (cudaStream_t)stream>>>(
      rel_codes, boxes, n, n_cls, w, out);
  DGOD_LAUNCHED();
  return DGOD_OK;
}

extern "C" int dgod_detect_candidates(const float* class_logits, const float* box_regression,
                                      const float* proposals, const int32_t* box_offsets,
                                      const float* image_sizes, int n_img, int n_rows, int n_cls,
                                      float wx, float wy, float ww, float wh, float xform_clip,
                                      float score_thresh, float min_size, float* cand_boxes,
                                      float* cand_scores, int64_t* cand_labels,
                                      uint8_t* cand_valid, dgod_stream_t stream) {
  DGOD_REQUIRE(n_img >= 0 && n_rows >= 0 && n_cls >= 2, "dgod_detect_candidates: bad sizes");
  if (n_img == 0 || n_rows == 0) return DGOD_OK;
  DGOD_REQUIRE(class_logits && box_regression && proposals && box_offsets && image_sizes &&
                   cand_boxes && cand_scores && cand_labels && cand_valid,
               "dgod_detect_candidates: null pointer");
  DecodeW w{wx, wy, ww, wh, xform_clip};
  detect_candidates_kernel<<<cdiv(n_rows, 128), 128, 0, (cudaStream_t)stream>>>(
      class_logits, box_regression, proposals, box_offsets, image_sizes, n_img, n_rows, n_cls, w,
      score_thresh, min_size, cand_boxes, cand_scores, cand_labels, cand_valid);
  DGOD_LAUNCHED();
  return DGOD_OK;
}
