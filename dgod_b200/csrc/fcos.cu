// FCOS location -> ground-truth assignment and target gather.
//
// Restates fcos.py:510-548 (FCOS.compute_loss) and fcos.py:136-158 (FCOSHead.compute_loss):
//   match(n,m) = max(|cx_n-gcx_m|, |cy_n-gcy_m|) < radius*size_n
//              & min(l,t,r,b) > 0
//              & lower_n < max(l,t,r,b) < upper_n         (lower 0 on the first level, upper inf on the last)
//   value(n,m) = match ? 1e8 - area_m : 0,   area_m = (y1-x1)*(y2-y1)   (sic, fcos.py:543)
//   matched_idx[n] = first arg-max_m value(n,m);  -1 when the maximum is < 1e-5
// All arithmetic is fp32 with the reference's operation order and no FMA contraction, so the
// int64 result is bit-exact (1e8 - area has an fp32 ulp of 8, i.e. real ties -> first index).
//
// One thread per location, ground truths of the image in shared memory; per image the kernel
// reads 16 B per location (anchors, shared by all images -> L2 resident) and writes 8 B
// (+8 +16 +4*classes for the optional targets): latency/HBM bound.
#include "common.cuh"

namespace dgod {

constexpr int kFcosThreads = 256;
constexpr int kFcosGtChunk = 256;

__global__ void __launch_bounds__(kFcosThreads)
fcos_assign_kernel(const float* __restrict__ anchors, int n_anchors, int n_first, int n_last,
                   float radius, const float* __restrict__ gt_boxes,
                   const int64_t* __restrict__ gt_labels, const int32_t* __restrict__ gt_offsets,
                   int64_t* __restrict__ matched_idx, int64_t* __restrict__ cls_targets,
                   float* __restrict__ box_targets, float* __restrict__ onehot, int num_classes) {
  __shared__ float4 s_box[kFcosGtChunk];
  __shared__ float2 s_ctr[kFcosGtChunk];
  __shared__ float s_val[kFcosGtChunk];  // 1e8 - area
  const int img = blockIdx.y;
  const int g0 = gt_offsets[img], g1 = gt_offsets[img + 1];
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = n < n_anchors;

  float acx = 0.f, acy = 0.f, asz = 0.f, lower = 0.f, upper = 0.f, rad = 0.f;
  if (active) {
    float4 a = ld_box(anchors, n);
    acx = __fdiv_rn(__fadd_rn(a.x, a.z), 2.f);   // fcos.py:520
    acy = __fdiv_rn(__fadd_rn(a.y, a.w), 2.f);
    asz = __fsub_rn(a.z, a.x);                   // fcos.py:521
    rad = __fmul_rn(radius, asz);                // fcos.py:525
    lower = n < n_first ? 0.f : __fmul_rn(asz, 4.f);                    // fcos.py:536-537
    upper = n >= n_anchors - n_last ? INFINITY : __fmul_rn(asz, 8.f);   // fcos.py:538-539
  }
  float best = 0.f;  // values are >= 0; an image without ground truth keeps idx -1 (fcos.py:512-516)
  int best_m = -1;
  for (int c0 = g0; c0 < g1; c0 += kFcosGtChunk) {
    const int cnt = min(kFcosGtChunk, g1 - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      float4 g = ld_box(gt_boxes, c0 + i);
      s_box[i] = g;
      s_ctr[i] = make_float2(__fdiv_rn(__fadd_rn(g.x, g.z), 2.f), __fdiv_rn(__fadd_rn(g.y, g.w), 2.f));
      float area = __fmul_rn(__fsub_rn(g.y, g.x), __fsub_rn(g.w, g.y));  // fcos.py:543 as written
      s_val[i] = __fsub_rn(1e8f, area);
    }
    __syncthreads();
    if (active) {
      for (int m = 0; m < cnt; ++m) {
        const float4 g = s_box[m];
        const float2 gc = s_ctr[m];
        float dx = fabsf(__fsub_rn(acx, gc.x)), dy = fabsf(__fsub_rn(acy, gc.y));
        bool match = fmaxf(dx, dy) < rad;
        float l = __fsub_rn(acx, g.x), t = __fsub_rn(acy, g.y);
        float r = __fsub_rn(g.z, acx), b = __fsub_rn(g.w, acy);
        float dmin = fminf(fminf(l, t), fminf(r, b));
        float dmax = fmaxf(fmaxf(l, t), fmaxf(r, b));
        match = match && (dmin > 0.f) && (dmax > lower) && (dmax < upper);
        float v = match ? __fmul_rn(1.f, s_val[m]) : __fmul_rn(0.f, s_val[m]);  // fcos.py:544
        if (best_m < 0 || v > best) { best = v; best_m = c0 - g0 + m; }
      }
    }
  }
  if (!active) return;
  long long idx = best_m;
  if (best_m < 0 || best < 1e-5f) idx = -1;  // fcos.py:546
  const long long o = (long long)img * n_anchors + n;
  if (matched_idx) matched_idx[o] = idx;

  // fcos.py:136-147 target gather (optional)
  const int n_gt = g1 - g0;
  long long cls = 0;
  float4 bt = make_float4(0, 0, 0, 0);
  if (n_gt > 1) {
    const long long cl = idx < 0 ? 0 : idx;
    if (gt_labels) cls = gt_labels[g0 + cl];
    bt = ld_box(gt_boxes, g0 + cl);
  }
  if (idx < 0) cls = -1;
  if (cls_targets) cls_targets[o] = cls;
  if (box_targets) reinterpret_cast<float4*>(box_targets)[o] = bt;
  if (onehot) {
    float* row = onehot + o * num_classes;
    for (int c = 0; c < num_classes; ++c) row[c] = (cls >= 0 && c == cls) ? 1.f : 0.f;  // fcos.py:157-158
  }
}

}  // namespace dgod

using namespace dgod;

extern "C" int dgod_fcos_assign(const float* anchors, int n_anchors, int n_first, int n_last,
                                double center_sampling_radius, const float* gt_boxes,
                                const int64_t* gt_labels, const int32_t* gt_offsets, int n_img,
                                int64_t* matched_idx, int64_t* cls_targets, float* box_targets,
                                float* onehot, int num_classes, dgod_stream_t stream) {
  DGOD_REQUIRE(n_anchors >= 0 && n_img >= 0 && n_first >= 0 && n_last >= 0,
               "dgod_fcos_assign: negative size");
  if (n_anchors == 0 || n_img == 0) return DGOD_OK;
  DGOD_REQUIRE(anchors && gt_offsets, "dgod_fcos_assign: null pointer");
  DGOD_REQUIRE(!onehot || num_classes > 0, "dgod_fcos_assign: onehot needs num_classes > 0");
  dim3 grid(cdiv(n_anchors, kFcosThreads), n_img);
  fcos_assign_kernel<<<grid, kFcosThreads, 0, (cudaStream_t)stream>>>(
      anchors, n_anchors, n_first, n_last, (float)center_sampling_radius, gt_boxes, gt_labels,
      gt_offsets, matched_idx, cls_targets, box_targets, onehot, num_classes);
  DGOD_LAUNCHED();
  return DGOD_OK;
}
