// Input side of the detector: GeneralizedRCNNTransform.forward
// (TV models/detection/transform.py:102-153, constructed at fasterrcnn.py:439-441 / fcos.py:483) —
//   normalize   (image - mean[:,None,None]) / std[:,None,None]                      (transform.py:155-166)
//   resize      F.interpolate(bilinear, align_corners=False, recompute_scale_factor=True)  (transform.py:25-83)
//   batch       zero-padded copy into [B, C, H_pad, W_pad], sizes rounded up to 32     (transform.py:237-255)
// as ONE launch for the whole batch.  The reference runs ~50 launches per batch (a normalize pair, an
// interpolate and a padded copy per image): 1.7 ms on a B200, host-bound.  Algorithmic bytes:
// sum_i C*h_i*w_i*4 read + B*C*H_pad*W_pad*4 written (102 + 60 MB at 8 x 3x800x1333 -> 608x1024): HBM-bound.
//
// Source coordinates follow ATen's upsample_bilinear2d (area_pixel_compute_source_index, align_corners
// false): src = max(fma(scale, dst+0.5, -0.5), 0) with scale = in/out in fp32, i0 = (int)src, i1 = i0 + (i0 < in-1),
// l1 = src - i0, l0 = 1 - l1, value = l0y*(l0x*v00 + l1x*v01) + l1y*(l0x*v10 + l1x*v11) on the normalised taps.
#include "common.cuh"

namespace dgod {

constexpr int kMaxBatchImages = 16;   // per launch (pointers and sizes travel as kernel parameters)

struct ImageBatchParams {
  const void* img[kMaxBatchImages];     // fp32 in [0,1], or uint8 in 0..255 (divided by 255 on load: DrivingDataset.py:53)
  int in_h[kMaxBatchImages], in_w[kMaxBatchImages], out_h[kMaxBatchImages], out_w[kMaxBatchImages];
  float mean[4], std[4];
  int channels, pad_h, pad_w;
};

constexpr int kPxPerThread = 4;       // consecutive output pixels per thread: one 128-bit store per channel

template <typename In> __device__ __forceinline__ float ld_px(const In* p);
template <> __device__ __forceinline__ float ld_px<float>(const float* p) { return __ldg(p); }
// the dataset's `image / 255.0` (DrivingDataset.py:53): one IEEE division of the byte value, bit-identical to the host's
template <> __device__ __forceinline__ float ld_px<uint8_t>(const uint8_t* p) { return __fdiv_rn((float)__ldg(p), 255.f); }

template <typename In>
__global__ void __launch_bounds__(256)
image_batch_kernel(const ImageBatchParams p, float* __restrict__ out, int first_image) {
  const int b = blockIdx.z, y = blockIdx.y;
  const int xb = (blockIdx.x * blockDim.x + threadIdx.x) * kPxPerThread;      // pad_w is a multiple of 4
  if (xb >= p.pad_w) return;
  const int ih = p.in_h[b], iw = p.in_w[b], oh = p.out_h[b], ow = p.out_w[b];
  float* dst = out + ((size_t)(first_image + b) * p.channels * p.pad_h + y) * p.pad_w + xb;
  const size_t cstride_out = (size_t)p.pad_h * p.pad_w;
  if (y >= oh || xb >= ow) {          // the padding of batch_images
    for (int c = 0; c < p.channels; ++c) *reinterpret_cast<float4*>(dst + c * cstride_out) = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float sy = (float)ih / (float)oh, sx = (float)iw / (float)ow;
  // ATen evaluates scale*(dst+0.5)-0.5 as one fused multiply-add (both its CPU and CUDA kernels contract it)
  const float fy = fmaxf(__fmaf_rn(sy, (float)y + 0.5f, -0.5f), 0.f);
  const int y0 = (int)fy;
  const int y1 = y0 + (y0 < ih - 1 ? 1 : 0);
  const float ly1 = fminf(fmaxf(fy - (float)y0, 0.f), 1.f), ly0 = 1.f - ly1;
  int x0[kPxPerThread], x1[kPxPerThread];
  float lx0[kPxPerThread], lx1[kPxPerThread];
#pragma unroll
  for (int q = 0; q < kPxPerThread; ++q) {
    const float fx = fmaxf(__fmaf_rn(sx, (float)(xb + q) + 0.5f, -0.5f), 0.f);
    x0[q] = min((int)fx, iw - 1);                     // columns past out_w (zeroed below) must still read in range
    x1[q] = x0[q] + (x0[q] < iw - 1 ? 1 : 0);
    lx1[q] = fminf(fmaxf(fx - (float)x0[q], 0.f), 1.f);
    lx0[q] = 1.f - lx1[q];
  }
  const In* src = reinterpret_cast<const In*>(p.img[b]);
  const size_t cstride_in = (size_t)ih * iw;
  const In* r0 = src + (size_t)y0 * iw;
  const In* r1 = src + (size_t)y1 * iw;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    if (c < p.channels) {
      float v[kPxPerThread][4];
#pragma unroll
      for (int q = 0; q < kPxPerThread; ++q) {        // the 16 taps of this channel in flight before the first blend
        v[q][0] = ld_px<In>(r0 + c * cstride_in + x0[q]); v[q][1] = ld_px<In>(r0 + c * cstride_in + x1[q]);
        v[q][2] = ld_px<In>(r1 + c * cstride_in + x0[q]); v[q][3] = ld_px<In>(r1 + c * cstride_in + x1[q]);
      }
      const float m = p.mean[c], sd = p.std[c];
      float o[kPxPerThread];
#pragma unroll
      for (int q = 0; q < kPxPerThread; ++q) {
        float v00 = v[q][0], v01 = v[q][1], v10 = v[q][2], v11 = v[q][3];
        if (m != 0.f || sd != 1.f) {   // (v - 0) / 1 == v exactly: DGFRCNN's transform (fasterrcnn.py:439-441) skips the IEEE divisions
          v00 = (v00 - m) / sd; v01 = (v01 - m) / sd; v10 = (v10 - m) / sd; v11 = (v11 - m) / sd;
        }
        o[q] = (xb + q < ow) ? ly0 * (lx0[q] * v00 + lx1[q] * v01) + ly1 * (lx0[q] * v10 + lx1[q] * v11) : 0.f;
      }
      *reinterpret_cast<float4*>(dst + c * cstride_out) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

}  // namespace dgod

using namespace dgod;

static int image_batch_impl(const void* const* images, bool u8, const int* in_h, const int* in_w, const int* out_h,
                            const int* out_w, int n_img, int channels, const float* mean, const float* std,
                            float* out, int pad_h, int pad_w, dgod_stream_t stream) {
  DGOD_REQUIRE(n_img >= 0 && channels >= 1 && channels <= 4 && pad_h > 0 && pad_w > 0, "dgod_image_batch: bad size");
  DGOD_REQUIRE(pad_w % 4 == 0 && ((uintptr_t)out & 15) == 0, "dgod_image_batch: pad_w must be a multiple of 4 and out 16-byte aligned");
  if (n_img == 0) return DGOD_OK;
  DGOD_REQUIRE(images && in_h && in_w && out_h && out_w && mean && std && out, "dgod_image_batch: null pointer");
  for (int first = 0; first < n_img; first += kMaxBatchImages) {
    const int n = n_img - first < kMaxBatchImages ? n_img - first : kMaxBatchImages;
    ImageBatchParams p = {};
    for (int i = 0; i < n; ++i) {
      const int k = first + i;
      DGOD_REQUIRE(images[k] && in_h[k] > 0 && in_w[k] > 0 && out_h[k] > 0 && out_w[k] > 0 && out_h[k] <= pad_h &&
                       out_w[k] <= pad_w,
                   "dgod_image_batch: image %d has an empty or oversized shape", k);
      p.img[i] = images[k];
      p.in_h[i] = in_h[k]; p.in_w[i] = in_w[k]; p.out_h[i] = out_h[k]; p.out_w[i] = out_w[k];
    }
    for (int c = 0; c < channels; ++c) {
      DGOD_REQUIRE(std[c] != 0.f, "dgod_image_batch: zero std");
      p.mean[c] = mean[c];
      p.std[c] = std[c];
    }
    p.channels = channels; p.pad_h = pad_h; p.pad_w = pad_w;
    dim3 grid(cdiv(pad_w, 256 * kPxPerThread), pad_h, n);
    if (u8) image_batch_kernel<uint8_t><<<grid, 256, 0, (cudaStream_t)stream>>>(p, out, first);
    else image_batch_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(p, out, first);
    DGOD_LAUNCHED();
  }
  return DGOD_OK;
}

extern "C" int dgod_image_batch(const float* const* images, const int* in_h, const int* in_w, const int* out_h,
                                const int* out_w, int n_img, int channels, const float* mean, const float* std,
                                float* out, int pad_h, int pad_w, dgod_stream_t stream) {
  return image_batch_impl(reinterpret_cast<const void* const*>(images), false, in_h, in_w, out_h, out_w, n_img, channels, mean,
                          std, out, pad_h, pad_w, stream);
}

extern "C" int dgod_image_batch_u8(const uint8_t* const* images, const int* in_h, const int* in_w, const int* out_h,
                                   const int* out_w, int n_img, int channels, const float* mean, const float* std,
                                   float* out, int pad_h, int pad_w, dgod_stream_t stream) {
  return image_batch_impl(reinterpret_cast<const void* const*>(images), true, in_h, in_w, out_h, out_w, n_img, channels, mean,
                          std, out, pad_h, pad_w, stream);
}
