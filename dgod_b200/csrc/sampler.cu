// BalancedPositiveNegativeSampler on the device without host synchronisation (SURVEY.md §8f rank 1;
// TV models/detection/_utils.py:11-71, used by the RPN loss fasterrcnn.py:119-123 and by
// RoIHeads.subsample TV roi_heads.py:615-622 <- fasterrcnn.py:272).
//
// Upstream draws `randperm(#pos)[:num_pos]` and `randperm(#neg)[:num_neg]` per image — a uniformly random subset of
// the positives and of the negatives — and returns boolean masks that the callers turn into index lists with
// torch.where (a host synchronisation each).  Here every candidate carries a uniform random key; the subset is "the k
// smallest keys of the class", found per image by a 4-pass 8-bit radix select over the key bits in ONE launch for the
// whole batch, and the survivors are written as fixed-capacity index lists in ascending index order (the order
// torch.where would give) plus validity masks.  Equal keys are resolved by ascending index.  Keys can be injected,
// which is how the tests force the selection the reference made.
//
// One CTA per image: reads labels + keys 4 times per class (L2 resident: 155 520 anchors x 8 B per image),
// latency-bound, ~2 x 155 k-wide torch.topk launches + their temporaries replaced by one.
#include "common.cuh"

namespace dgod {

constexpr int kSampThreads = 1024;

template <typename L> __device__ __forceinline__ int label_class(const L* __restrict__ labels, long long i);
// 1: positive (label >= 1), 0: negative (label == 0), -1: ignored
template <> __device__ __forceinline__ int label_class<float>(const float* __restrict__ labels, long long i) {
  const float v = __ldg(labels + i);
  return v >= 1.f ? 1 : (v == 0.f ? 0 : -1);
}
template <> __device__ __forceinline__ int label_class<int64_t>(const int64_t* __restrict__ labels, long long i) {
  const long long v = __ldg(labels + i);
  return v >= 1 ? 1 : (v == 0 ? 0 : -1);
}

struct SampShared {
  unsigned hist[256];
  unsigned warp[kSampThreads / 32];
  unsigned prefix, remaining, total, base_above, base_eq;
};

// The `want` smallest keys among the elements of class `cls` of one image -> out_idx (ascending index), out_valid.
// Returns (to every thread) how many were selected.
template <typename L>
__device__ unsigned select_class(const L* __restrict__ labels, const float* __restrict__ keys, int n, int cls, unsigned want,
                                 int capacity, int64_t* __restrict__ out_idx, uint8_t* __restrict__ out_valid, SampShared& sh) {
  const int tid = threadIdx.x;
  unsigned prefix = 0, mask = 0, k = 0;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += kSampThreads) sh.hist[i] = 0;
    __syncthreads();
    for (int e = tid; e < n; e += kSampThreads) {
      if (label_class<L>(labels, e) == cls) {
        const unsigned key = float_ordered(__ldg(keys + e));
        if ((key & mask) == prefix) atomicAdd(&sh.hist[(key >> shift) & 255u], 1u);
      }
    }
    __syncthreads();
    if (tid == 0) {
      if (pass == 0) {
        unsigned total = 0;
        for (int b = 0; b < 256; ++b) total += sh.hist[b];
        sh.total = total;
        sh.remaining = min(want, total);
      }
      unsigned rem = sh.remaining, b = 0;
      if (rem > 0) {
        for (; b < 255; ++b) {
          if (sh.hist[b] >= rem) break;
          rem -= sh.hist[b];
        }
      }
      sh.prefix = prefix | (b << shift);
      sh.remaining = rem;
    }
    __syncthreads();
    if (pass == 0) k = min(want, sh.total);
    prefix = sh.prefix;
    mask |= 255u << shift;
    if (k == 0) break;
    __syncthreads();
  }
  const unsigned kth = prefix, n_eq = sh.remaining;      // key of the k-th smallest; n_eq of the equals are taken
  __syncthreads();
  if (tid == 0) { sh.base_above = 0; sh.base_eq = 0; }
  __syncthreads();
  // ordered compaction: every element below the k-th key, the first n_eq equal ones, in ascending index
  for (int e0 = 0; e0 < n && k > 0; e0 += kSampThreads) {
    const int e = e0 + tid;
    bool below = false, eq = false;
    if (e < n && label_class<L>(labels, e) == cls) {
      const unsigned key = float_ordered(__ldg(keys + e));
      below = key < kth;
      eq = key == kth;
    }
    const unsigned bal_b = __ballot_sync(0xffffffffu, below), bal_e = __ballot_sync(0xffffffffu, eq);
    if ((tid & 31) == 0) sh.warp[tid >> 5] = __popc(bal_b) | (__popc(bal_e) << 16);
    __syncthreads();
    unsigned before_b = 0, before_e = 0;
    for (int w = 0; w < (tid >> 5); ++w) { before_b += sh.warp[w] & 0xffffu; before_e += sh.warp[w] >> 16; }
    const unsigned lane_lt = (1u << (tid & 31)) - 1u;
    const unsigned eq_rank = sh.base_eq + before_e + __popc(bal_e & lane_lt);
    const bool take = below || (eq && eq_rank < n_eq);
    // position = (#taken before this element): below-count so far + min(eq-count so far, n_eq)
    const unsigned below_before = sh.base_above + before_b + __popc(bal_b & lane_lt);
    const unsigned eq_before = min(eq_rank, n_eq);
    if (take) {
      const unsigned pos = below_before + eq_before;
      if ((int)pos < capacity) { out_idx[pos] = e; out_valid[pos] = 1; }
    }
    __syncthreads();
    if (tid == 0) {
      unsigned tb = 0, te = 0;
      for (int w = 0; w < kSampThreads / 32; ++w) { tb += sh.warp[w] & 0xffffu; te += sh.warp[w] >> 16; }
      sh.base_above += tb;
      sh.base_eq += te;
    }
    __syncthreads();
  }
  for (int i = (int)k + tid; i < capacity; i += kSampThreads) { out_idx[i] = 0; out_valid[i] = 0; }
  __syncthreads();
  return k;
}

template <typename L>
__global__ void __launch_bounds__(kSampThreads)
balanced_sample_kernel(const L* __restrict__ labels, const float* __restrict__ keys, int n, int num_pos, int batch_size,
                       int cap_pos, int cap_neg, int64_t* __restrict__ pos_idx, uint8_t* __restrict__ pos_valid,
                       int64_t* __restrict__ neg_idx, uint8_t* __restrict__ neg_valid, int32_t* __restrict__ counts) {
  __shared__ SampShared sh;
  const int img = blockIdx.x;
  labels += (size_t)img * n;
  keys += (size_t)img * n;
  // TV _utils.py:44-50: num_pos = min(#pos, batch * fraction); num_neg = min(#neg, batch - num_pos)
  const unsigned np = select_class<L>(labels, keys, n, 1, (unsigned)num_pos, cap_pos, pos_idx + (size_t)img * cap_pos,
                                      pos_valid + (size_t)img * cap_pos, sh);
  const unsigned nn = select_class<L>(labels, keys, n, 0, (unsigned)batch_size - np, cap_neg, neg_idx + (size_t)img * cap_neg,
                                      neg_valid + (size_t)img * cap_neg, sh);
  if (threadIdx.x == 0) { counts[2 * img] = (int)np; counts[2 * img + 1] = (int)nn; }
}

}  // namespace dgod

using namespace dgod;

extern "C" int dgod_balanced_sample(const void* labels, int labels_are_int64, const float* keys, int n_img, int n,
                                    int num_pos, int batch_size_per_image, int64_t* pos_idx, uint8_t* pos_valid,
                                    int64_t* neg_idx, uint8_t* neg_valid, int32_t* counts, dgod_stream_t stream) {
  DGOD_REQUIRE(n_img >= 0 && n >= 0 && num_pos >= 0 && batch_size_per_image >= num_pos, "dgod_balanced_sample: bad size");
  if (n_img == 0) return DGOD_OK;
  const int cap_pos = num_pos < n ? num_pos : n, cap_neg = batch_size_per_image < n ? batch_size_per_image : n;
  DGOD_REQUIRE(labels && keys && counts && (cap_pos == 0 || (pos_idx && pos_valid)) && (cap_neg == 0 || (neg_idx && neg_valid)),
               "dgod_balanced_sample: null pointer");
  if (labels_are_int64)
    balanced_sample_kernel<int64_t><<<n_img, kSampThreads, 0, (cudaStream_t)stream>>>(
        (const int64_t*)labels, keys, n, num_pos, batch_size_per_image, cap_pos, cap_neg, pos_idx, pos_valid, neg_idx, neg_valid,
        counts);
  else
    balanced_sample_kernel<float><<<n_img, kSampThreads, 0, (cudaStream_t)stream>>>(
        (const float*)labels, keys, n, num_pos, batch_size_per_image, cap_pos, cap_neg, pos_idx, pos_valid, neg_idx, neg_valid,
        counts);
  DGOD_LAUNCHED();
  return DGOD_OK;
}
