// FCOS eval post-processing up to the NMS: fcos.py:576-597 for a whole batch in ONE launch.
//
// Per image and pyramid level the reference runs ~12 ATen kernels: score = sqrt(sigmoid(cls) * sigmoid(ctrness))
// over [locations x classes], `> score_thresh`, top-k 1000, index split into (location, class), BoxLinearCoder
// decode (fcos.py:72-100, normalize_by_size), clip to the image.  Here one CTA owns an (image, level) pair:
//   1. a 4-pass 8-bit radix select over the score bits finds the k-th largest passing score (scores are positive
//      floats, so their bit patterns order like the values); scores are recomputed per pass instead of stored;
//   2. the survivors (all scores above the k-th, and the lowest-index ones among equals) are gathered and sorted in
//      shared memory: descending score, ascending flat index (location * classes + class) on equal scores — torch
//      leaves that order implementation-defined;
//   3. the k survivors are decoded, clipped and written to fixed-capacity outputs [n_img, n_levels * topk] plus a
//      validity mask, which is what the segmented NMS (dgod_nms_batched) takes: no host synchronisation.
// The sigmoid uses expf: scores agree with the CPU path to ~1 ulp (tests: 1e-6 relative, survivor SET identical up
// to scores that close to the threshold / the k-th score); decode and clip are exact fp32.
//
// Latency-bound (a few MB per batch); the point is 1 launch instead of ~12 x levels x images.
#include "common.cuh"

namespace dgod {

constexpr int kPostThreads = 1024;
constexpr int kPostMaxK = 1024;        // topk_candidates handled per (image, level)

__device__ __forceinline__ float sigmoid_f(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }

// score bits of flat element e of this (image, level), 0 when it does not pass the threshold
__device__ __forceinline__ unsigned fcos_key(const float* __restrict__ cls, const float* __restrict__ ctr, int e, int C,
                                             float thresh) {
  const int loc = e / C;
  const float s = __fsqrt_rn(__fmul_rn(sigmoid_f(__ldg(cls + e)), sigmoid_f(__ldg(ctr + loc))));
  return s > thresh ? __float_as_uint(s) : 0u;
}

__global__ void __launch_bounds__(kPostThreads)
fcos_candidates_kernel(const float* __restrict__ cls_logits, const float* __restrict__ bbox_regression,
                       const float* __restrict__ bbox_ctrness, const float* __restrict__ anchors, int n_anchors, int C,
                       const int* __restrict__ level_offsets, int n_levels, const float* __restrict__ image_sizes,
                       float thresh, int topk, float* __restrict__ out_boxes, float* __restrict__ out_scores,
                       int64_t* __restrict__ out_labels, uint8_t* __restrict__ out_valid, int32_t* __restrict__ out_count) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_key[kPostMaxK];
  __shared__ int s_idx[kPostMaxK];
  __shared__ unsigned s_prefix, s_remaining, s_n, s_warp[kPostThreads / 32], s_eq_base;
  const int level = blockIdx.x, img = blockIdx.y, tid = threadIdx.x;
  const int a0 = level_offsets[level], n_loc = level_offsets[level + 1] - a0;
  const int n = n_loc * C;
  const float* cls = cls_logits + ((size_t)img * n_anchors + a0) * C;
  const float* ctr = bbox_ctrness + (size_t)img * n_anchors + a0;
  const size_t out0 = ((size_t)img * n_levels + level) * topk;

  // ---- 1. radix select of the k-th largest passing score
  unsigned prefix = 0, mask = 0, remaining = 0, k = 0;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += kPostThreads) hist[i] = 0;
    __syncthreads();
    for (int e = tid; e < n; e += kPostThreads) {
      const unsigned key = fcos_key(cls, ctr, e, C, thresh);
      if (key && (key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      if (pass == 0) {
        unsigned total = 0;
        for (int b = 0; b < 256; ++b) total += hist[b];
        s_n = total;
        s_remaining = min((unsigned)topk, total);
      }
      unsigned rem = s_remaining, b = 255;
      if (rem > 0) {
        for (;; --b) {
          if (hist[b] >= rem) break;
          rem -= hist[b];
          if (b == 0) break;
        }
      }
      s_prefix = prefix | (b << shift);
      s_remaining = rem;
    }
    __syncthreads();
    if (pass == 0) k = min((unsigned)topk, s_n);
    prefix = s_prefix;
    remaining = s_remaining;
    mask |= 255u << shift;
    if (k == 0) break;
    __syncthreads();
  }
  if (tid == 0) out_count[img * n_levels + level] = (int)k;
  for (int i = tid; i < topk; i += kPostThreads) out_valid[out0 + i] = i < (int)k;
  if (k == 0) {
    for (int i = tid; i < topk; i += kPostThreads) {
      reinterpret_cast<float4*>(out_boxes)[out0 + i] = make_float4(0.f, 0.f, 0.f, 0.f);
      out_scores[out0 + i] = 0.f;
      out_labels[out0 + i] = 0;
    }
    return;
  }
  const unsigned kth = prefix;             // bit pattern of the k-th largest score; `remaining` of the equals survive

  // ---- 2. gather: everything above the k-th score, then the lowest-index equals
  for (int i = tid; i < kPostMaxK; i += kPostThreads) { s_key[i] = 0; s_idx[i] = 0x7fffffff; }
  if (tid == 0) { s_n = 0; s_eq_base = 0; }
  __syncthreads();
  const unsigned n_above = k - remaining;
  for (int e0 = 0; e0 < n; e0 += kPostThreads) {
    const int e = e0 + tid;
    const unsigned key = e < n ? fcos_key(cls, ctr, e, C, thresh) : 0u;
    if (key > kth) {
      const unsigned slot = atomicAdd(&s_n, 1u);
      s_key[slot] = key; s_idx[slot] = e;
    }
    const bool eq = key == kth;
    if (__syncthreads_or(eq)) {            // rank the equals of this chunk in index order
      const unsigned bal = __ballot_sync(0xffffffffu, eq);
      if ((tid & 31) == 0) s_warp[tid >> 5] = __popc(bal);
      __syncthreads();
      unsigned before = s_eq_base;
      for (int w = 0; w < (tid >> 5); ++w) before += s_warp[w];
      const unsigned rank = before + __popc(bal & ((1u << (tid & 31)) - 1u));
      if (eq && rank < remaining) { s_key[n_above + rank] = key; s_idx[n_above + rank] = e; }
      __syncthreads();
      if (tid == 0) {
        unsigned tot = 0;
        for (int w = 0; w < kPostThreads / 32; ++w) tot += s_warp[w];
        s_eq_base += tot;
      }
      __syncthreads();
    }
  }
  __syncthreads();

  // ---- 3. bitonic sort of the kPostMaxK slots: descending score, ascending flat index on ties (empty slots last)
  for (int size = 2; size <= kPostMaxK; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const int i = tid, j = i ^ stride;
      if (j > i) {
        const bool up = (i & size) == 0;
        const unsigned ki = s_key[i], kj = s_key[j];
        const int xi = s_idx[i], xj = s_idx[j];
        const bool i_first = ki > kj || (ki == kj && xi < xj);     // i belongs before j in the final order
        if (i_first != up) { s_key[i] = kj; s_key[j] = ki; s_idx[i] = xj; s_idx[j] = xi; }
      }
      __syncthreads();
    }
  }

  // ---- 4. decode (fcos.py:72-100), clip (TV ops/boxes.py:149-182), write
  const float img_h = image_sizes[2 * img], img_w = image_sizes[2 * img + 1];
  for (int i = tid; i < topk; i += kPostThreads) {
    float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
    float score = 0.f;
    long long label = 0;
    if (i < (int)k) {
      const int e = s_idx[i], loc = e / C;
      label = e - loc * C;
      score = __uint_as_float(s_key[i]);
      const float4 a = ld_box(anchors, a0 + loc);
      const float4 r = __ldg(reinterpret_cast<const float4*>(bbox_regression) + (size_t)img * n_anchors + a0 + loc);
      const float cx = __fmul_rn(0.5f, __fadd_rn(a.x, a.z)), cy = __fmul_rn(0.5f, __fadd_rn(a.y, a.w));
      const float w = __fsub_rn(a.z, a.x), h = __fsub_rn(a.w, a.y);
      box.x = __fsub_rn(cx, __fmul_rn(r.x, w));
      box.y = __fsub_rn(cy, __fmul_rn(r.y, h));
      box.z = __fadd_rn(cx, __fmul_rn(r.z, w));
      box.w = __fadd_rn(cy, __fmul_rn(r.w, h));
      box.x = fminf(fmaxf(box.x, 0.f), img_w); box.z = fminf(fmaxf(box.z, 0.f), img_w);
      box.y = fminf(fmaxf(box.y, 0.f), img_h); box.w = fminf(fmaxf(box.w, 0.f), img_h);
    }
    reinterpret_cast<float4*>(out_boxes)[out0 + i] = box;
    out_scores[out0 + i] = score;
    out_labels[out0 + i] = label;
  }
}

}  // namespace dgod

using namespace dgod;

extern "C" int dgod_fcos_candidates(const float* cls_logits, const float* bbox_regression, const float* bbox_ctrness,
                                    const float* anchors, int n_anchors, int num_classes,
                                    const int32_t* level_offsets, int n_levels, int n_img, const float* image_sizes,
                                    float score_thresh, int topk, float* out_boxes, float* out_scores,
                                    int64_t* out_labels, uint8_t* out_valid, int32_t* out_count, dgod_stream_t stream) {
  DGOD_REQUIRE(n_anchors >= 0 && n_img >= 0 && n_levels >= 0 && num_classes > 0, "dgod_fcos_candidates: bad size");
  DGOD_REQUIRE(topk >= 1 && topk <= kPostMaxK, "dgod_fcos_candidates: topk must be in 1..1024");
  DGOD_REQUIRE(score_thresh >= 0.f, "dgod_fcos_candidates: the score threshold must be >= 0");
  if (n_img == 0 || n_levels == 0) return DGOD_OK;
  DGOD_REQUIRE(cls_logits && bbox_regression && bbox_ctrness && anchors && level_offsets && image_sizes && out_boxes &&
                   out_scores && out_labels && out_valid && out_count,
               "dgod_fcos_candidates: null pointer");
  DGOD_REQUIRE(((uintptr_t)bbox_regression & 15) == 0 && ((uintptr_t)anchors & 15) == 0 && ((uintptr_t)out_boxes & 15) == 0,
               "dgod_fcos_candidates: box tensors must be 16-byte aligned");
  fcos_candidates_kernel<<<dim3(n_levels, n_img), kPostThreads, 0, (cudaStream_t)stream>>>(
      cls_logits, bbox_regression, bbox_ctrness, anchors, n_anchors, num_classes, level_offsets, n_levels, image_sizes,
      score_thresh, topk, out_boxes, out_scores, out_labels, out_valid, out_count);
  DGOD_LAUNCHED();
  return DGOD_OK;
}
