// Bitmask batched NMS.
//
// Reference behaviour restated (TV = torchvision 0.26.0):
//   nms                 TV ops/boxes.py:20-48 -> torchvision::nms CPU kernel: stable descending
//                       sort of the scores, greedy pass, suppress iff IoU > threshold with the
//                       fp32 IoU widened to double for the comparison (SURVEY.md §8c probes)
//   batched_nms         TV ops/boxes.py:51-120: coordinate-offset trick (<= 1000 boxes on CPU)
//                       or per-group loop ("vanilla"), both honoured via `offset_mode`
//
// Pipeline (all segments = images and all groups = levels/classes in one set of launches):
//   1. prepare   64-bit key = segment:16 | group:16 | ~ordered(score):32, value = input index
//   2. sort      bitonic, ascending (key, index)  => processing order, runs = (segment, group)
//   3. gather    boxes into processing order (+ the offset trick's fp32 shift), 32-bit run keys
//   4. mask      64x64 IoU tiles from shared memory -> per-row 64-bit suppression words, only
//                inside a run and above the diagonal (block-diagonal, upper-triangular)
//   5. scan      one CTA per run: resolve 64 candidates per step from the diagonal word, OR the
//                kept rows into the run's removed-bitmap in shared memory
//   6. re-key kept entries as segment | ~score, sort again, write segment-relative indices in
//      descending score order, truncated to max_out_per_seg
//
// Bytes: 28 B per box in and 8 B per kept box out are compulsory; the mask adds
// 2 * 8 B * sum_runs n_r^2/128 (write + read).  The pair evaluations (sum_runs n_r^2/2, ~25
// fp32 instructions each) are the real cost for large runs — see DESIGN.md for the roofline.
#include "nms_core.cuh"
#include "sort.cuh"
#include <utility>

namespace dgod {

constexpr unsigned long long kKeyMax = ~0ull;

// ---------------------------------------------------------------------------- mask
// One thread per (row, 64-column chunk): 256 threads = 64 rows x 4 consecutive column chunks,
// column chunk c = r + 4*blockIdx.y + tid/64, i.e. only chunks on or above the diagonal are ever
// launched.  Column boxes sit in shared memory and are read as warp broadcasts (a warp = 32 rows of
// one chunk).  The IoU > threshold test avoids the IEEE division on all but borderline pairs:
// with t = thr*union, inter > t(1+2^-20) implies fl(inter/union) > thr and inter < t(1-2^-20)
// implies fl(inter/union) < thr (each fp32 rounding moves a value by at most 2^-24 relative);
// only pairs in between take the exact quotient.  Non-overlapping pairs exit after 4 min/max.
constexpr int kMaskChunksPerCta = 4;

__device__ __forceinline__ bool iou_gt_exact(const float4 a, float area_a, const float4 b, float thr,
                                             bool skip_disjoint) {
  const float w = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
  const float h = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
  if (skip_disjoint && (w <= 0.f || h <= 0.f)) return false;   // IoU is 0 (or 0/0): never > thr >= 0
  const float inter = __fmul_rn(fmaxf(w, 0.f), fmaxf(h, 0.f));
  const float area_b = box_area_exact(b.x, b.y, b.z, b.w);
  const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  if (uni > 0.f && thr > 0.f) {
    const float t = __fmul_rn(thr, uni);
    if (inter > __fmul_rn(t, 1.00000095367431640625f)) return true;     // 1 + 2^-20
    if (inter < __fmul_rn(t, 0.99999904632568359375f)) return false;    // 1 - 2^-20
  }
  return __fdiv_rn(inter, uni) > thr;
}

__global__ void __launch_bounds__(256)
nms_mask_kernel(const float4* __restrict__ sbox, const uint32_t* __restrict__ runkey, int n_pos,
                float thr, unsigned long long* __restrict__ mask, int row_words,
                unsigned long long* __restrict__ diag_cols) {
  __shared__ float4 s_box[kMaskChunksPerCta][64];
  __shared__ uint32_t s_key[kMaskChunksPerCta][64];
  __shared__ uint32_t s_colbits[64][2];
  const int r = blockIdx.x;                                   // row chunk
  const int c0 = r + kMaskChunksPerCta * blockIdx.y;          // first column chunk of this CTA
  const int row0 = r * 64;
  const long long col0 = (long long)c0 * 64;
  if (row0 >= n_pos || col0 >= n_pos) return;
  // run keys are non-decreasing: if the first column is already past the run of the chunk's last
  // real row, no pair of this CTA shares a run.  Padding rows (kNoRun) never own pairs.
  int last = min(row0 + 63, n_pos - 1);
  uint32_t rk_last = runkey[last];
  while (rk_last == kNoRun && last > row0) rk_last = runkey[--last];
  if (rk_last == kNoRun || runkey[col0] > rk_last) {
    if (blockIdx.y == 0 && threadIdx.x < 64 && row0 + (int)threadIdx.x < n_pos) diag_cols[row0 + threadIdx.x] = 0ull;
    return;
  }

  {
    const int cc = threadIdx.x >> 6, b = threadIdx.x & 63;
    const long long q = col0 + threadIdx.x;
    float4 bx = make_float4(0, 0, 0, 0);
    uint32_t k = kNoRun;
    if (q < n_pos) { bx = sbox[q]; k = runkey[q]; }
    s_box[cc][b] = bx;
    s_key[cc][b] = k;
  }
  __syncthreads();

  const int cc = threadIdx.x >> 6, row = threadIdx.x & 63;
  const int p = row0 + row;
  const int w = c0 + cc - r;                                  // word index inside the mask row
  const bool diag = blockIdx.y == 0 && cc == 0;               // warps 0-1 of the first CTA: the diagonal tile
  const bool active = p < n_pos && w < row_words;
  const uint32_t rk = active ? runkey[p] : kNoRun;
  if (!diag && rk == kNoRun) return;
  unsigned long long bits = 0ull;
  const long long qbase = col0 + cc * 64;
  if (rk != kNoRun && qbase + 63 > p && s_key[cc][0] <= rk && s_key[cc][63] >= rk) {
    const float4 a = sbox[p];
    const float area_a = box_area_exact(a.x, a.y, a.z, a.w);
    const bool skip_disjoint = thr >= 0.f;
#pragma unroll 4
    for (int b = 0; b < 64; ++b) {
      if (s_key[cc][b] == rk && qbase + b > p && iou_gt_exact(a, area_a, s_box[cc][b], thr, skip_disjoint))
        bits |= (1ull << b);
    }
  }
  if (rk != kNoRun) mask[(size_t)p * row_words + w] = bits;
  if (diag) {
    // transpose of the diagonal 64x64 block: diag_cols[q] = rows of the chunk that suppress q
    // (what the scan's fixpoint resolve consumes); rows 0-31 ballot in warp 0, rows 32-63 in warp 1
#pragma unroll 8
    for (int b = 0; b < 64; ++b) {
      const uint32_t bal = __ballot_sync(0xffffffffu, (bits >> b) & 1ull);
      if ((threadIdx.x & 31) == 0) s_colbits[b][threadIdx.x >> 5] = bal;
    }
    // the two warps meet on a named barrier (the other six warps may already have left)
    asm volatile("bar.sync 1, 64;\n" ::);
    if (p < n_pos) diag_cols[p] = ((unsigned long long)s_colbits[row][1] << 32) | s_colbits[row][0];
  }
}

int launch_nms_mask(const float4* sbox, const uint32_t* runkey, int n_pos, int max_run_len,
                    float thr, unsigned long long* mask, unsigned long long* diag_cols, cudaStream_t st) {
  if (n_pos <= 0) return DGOD_OK;
  const int row_words = nms_mask_row_words(max_run_len);
  // a run of length L starting anywhere inside row chunk r reaches column chunk r + L/64 + 1 at most
  const int groups = (row_words + kMaskChunksPerCta - 1) / kMaskChunksPerCta;
  dim3 grid(cdiv(n_pos, 64), groups);
  nms_mask_kernel<<<grid, 256, 0, st>>>(sbox, runkey, n_pos, thr, mask, row_words, diag_cols);
  DGOD_LAUNCHED();
  return DGOD_OK;
}

// ---------------------------------------------------------------------------- scan
// One CTA per run (the CTA of the chunk in which the run starts).  Per 64-candidate chunk: the
// chunk's mask rows (all words up to the run's end) are brought into shared memory with cp.async —
// the next chunk's rows stream in while this one is resolved — thread 0 walks the alive bits of
// the diagonal word, then every thread ORs the kept rows of its word into the removed-bitmap.
constexpr int kScanThreads = 256;
constexpr int kScanThreadsLong = 1024;   // runs whose mask rows do not fit shared memory
constexpr int kScanMaxWords = 160;   // runs up to ~10k candidates use the staged path

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}

__global__ void __launch_bounds__(kScanThreadsLong)
nms_scan_kernel(const unsigned long long* __restrict__ mask, const unsigned long long* __restrict__ diag_cols,
                int row_words, const uint32_t* __restrict__ runkey, const uint8_t* __restrict__ alive, int n_pos,
                unsigned long long* __restrict__ keepbits, int32_t* __restrict__ compact_pos,
                int32_t* __restrict__ run_count, int staged) {
  extern __shared__ unsigned long long s_dyn[];     // removed[row_words] | rows[2][64][row_words] (staged)
  unsigned long long* s_removed = s_dyn;
  unsigned long long* s_rows = s_dyn + row_words;
  __shared__ uint32_t s_ballot[2];
  __shared__ unsigned long long s_kept;
  const int c = blockIdx.x, tid = threadIdx.x;
  const int p0 = c * 64;
  if (tid < 64) {
    const int p = p0 + tid;
    bool start = false;
    if (p < n_pos) {
      uint32_t k = runkey[p];
      start = k != kNoRun && (p == 0 || runkey[p - 1] != k);
    }
    uint32_t bal = __ballot_sync(0xffffffffu, start);
    if ((tid & 31) == 0) s_ballot[tid >> 5] = bal;
  }
  __syncthreads();
  unsigned long long starts = ((unsigned long long)s_ballot[1] << 32) | s_ballot[0];
  if (starts == 0ull) return;

  while (starts) {
    const int sb = __ffsll((long long)starts) - 1;
    starts &= starts - 1;
    const int s = p0 + sb;
    const uint32_t rk = runkey[s];
    int lo = s + 1, hi = n_pos;  // upper bound of rk in the non-decreasing key array
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (runkey[mid] <= rk) lo = mid + 1; else hi = mid;
    }
    const int e = lo;
    const int cs = s >> 6, ce = (e - 1) >> 6;
    __syncthreads();  // previous run done with the shared buffers
    for (int w = tid; w <= ce - cs && w < row_words; w += blockDim.x) s_removed[w] = 0ull;
    auto stage = [&](int cc, int buf) {      // rows of chunk cc, words 0..ce-cc, into rows[buf]
      const int nw = ce - cc + 1;
      unsigned long long* dst = s_rows + (size_t)buf * 64 * row_words;
      for (int i = tid; i < 64 * nw; i += blockDim.x) {
        const int row = i / nw, w = i - row * nw;
        const int p = cc * 64 + row;
        if (p >= s && p < e) cp_async8(dst + row * row_words + w, mask + (size_t)p * row_words + w);
      }
      asm volatile("cp.async.commit_group;\n" ::);
    };
    if (staged) stage(cs, 0);
    int count = 0;
    for (int cc = cs; cc <= ce; ++cc) {
      const int buf = (cc - cs) & 1;
      const unsigned long long* rows = s_rows + (size_t)buf * 64 * row_words;
      bool in = false;
      if (tid < 64) {
        const int p = cc * 64 + tid;
        in = p >= s && p < e && (!alive || alive[p]);
        uint32_t bal = __ballot_sync(0xffffffffu, in);
        if ((tid & 31) == 0) s_ballot[tid >> 5] = bal;
      }
      if (staged) asm volatile("cp.async.wait_group 0;\n" ::);
      __syncthreads();                                   // rows of chunk cc landed; removed[] is current
      if (staged && cc < ce) stage(cc + 1, buf ^ 1);     // overlaps with the resolve below
      if (tid < 32) {
        // Greedy keep set of the chunk as the fixpoint of
        //   K[b] = alive[b] and no a < b with K[a] and D[a][b]
        // iterated from K = alive: entry b is final after b+1 sweeps at the latest, in practice
        // after (longest suppression chain + 1) sweeps; each sweep is two ballots.
        const unsigned long long al = (((unsigned long long)s_ballot[1] << 32) | s_ballot[0]) & ~s_removed[cc - cs];
        const int q0 = cc * 64 + tid, q1 = q0 + 32;
        const unsigned long long col0 = q0 < n_pos ? diag_cols[q0] : 0ull;   // only rows a < b of the same run
        const unsigned long long col1 = q1 < n_pos ? diag_cols[q1] : 0ull;
        unsigned long long K = al;
        while (true) {
          const bool k0 = ((al >> tid) & 1ull) && ((col0 & K) == 0ull);
          const bool k1 = ((al >> (tid + 32)) & 1ull) && ((col1 & K) == 0ull);
          const unsigned long long Kn = ((unsigned long long)__ballot_sync(0xffffffffu, k1) << 32) | __ballot_sync(0xffffffffu, k0);
          if (Kn == K) break;
          K = Kn;
        }
        if (tid == 0) s_kept = K;
      }
      __syncthreads();
      const unsigned long long kept = s_kept;
      if (staged) {
        // 8 threads per word: each ORs every 8th kept row, then a 3-step shuffle OR
        for (int w0 = 1; cc + w0 <= ce; w0 += blockDim.x / 8) {
          const int w = w0 + (tid >> 3), part = tid & 7;
          unsigned long long acc = 0ull;
          if (cc + w <= ce) {
            unsigned long long kk = kept & (0x0101010101010101ull << part);
            while (kk) {
              const int b = __ffsll((long long)kk) - 1;
              kk &= kk - 1;
              acc |= rows[b * row_words + w];
            }
          }
          acc |= __shfl_xor_sync(0xffffffffu, acc, 1);
          acc |= __shfl_xor_sync(0xffffffffu, acc, 2);
          acc |= __shfl_xor_sync(0xffffffffu, acc, 4);
          if (part == 0 && cc + w <= ce) s_removed[cc - cs + w] |= acc;
        }
      } else {
        // long runs: the rows stay in global memory (L2).  8 threads per word, each with its up to 8 kept rows
        // in flight at once, then a 3-step shuffle OR; the launch uses 1024 threads for this path.
        for (int w0 = 1; cc + w0 <= ce; w0 += blockDim.x / 8) {
          const int w = w0 + (tid >> 3), part = tid & 7;
          unsigned long long acc = 0ull;
          if (cc + w <= ce) {
            unsigned long long kk = kept & (0x0101010101010101ull << part);
            unsigned long long v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              v[u] = 0ull;
              if (kk) {
                const int b = __ffsll((long long)kk) - 1;
                kk &= kk - 1;
                v[u] = mask[(size_t)(cc * 64 + b) * row_words + w];
              }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) acc |= v[u];
          }
          acc |= __shfl_xor_sync(0xffffffffu, acc, 1);
          acc |= __shfl_xor_sync(0xffffffffu, acc, 2);
          acc |= __shfl_xor_sync(0xffffffffu, acc, 4);
          if (part == 0 && cc + w <= ce) s_removed[cc - cs + w] |= acc;
        }
      }
      if (tid < 64 && ((kept >> tid) & 1ull) && compact_pos) {
        const int j = count + __popcll(kept & ((1ull << tid) - 1ull));
        compact_pos[s + j] = cc * 64 + tid;
      }
      if (tid == 0 && kept) atomicOr(&keepbits[cc], kept);
      count += __popcll(kept);
    }
    if (run_count && tid == 0) run_count[rk] = count;
  }
}

int launch_nms_scan(const unsigned long long* mask, const unsigned long long* diag_cols, const uint32_t* runkey,
                    const uint8_t* alive, int n_pos, int max_run_len, unsigned long long* keepbits,
                    int32_t* compact_pos, int32_t* run_count, cudaStream_t st) {
  if (n_pos <= 0) return DGOD_OK;
  const int row_words = nms_mask_row_words(max_run_len);
  const int staged = row_words <= kScanMaxWords;
  const size_t smem = (size_t)row_words * sizeof(unsigned long long) * (staged ? 1 + 2 * 64 : 1);
  static size_t attr_smem = 0;
  if (smem > 48 * 1024 && smem > attr_smem) {
    DGOD_CUDA(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  nms_scan_kernel<<<cdiv(n_pos, 64), staged ? kScanThreads : kScanThreadsLong, smem, st>>>(mask, diag_cols, row_words, runkey, alive, n_pos, keepbits,
                                                              compact_pos, run_count, staged);
  DGOD_LAUNCHED();
  return DGOD_OK;
}

// ---------------------------------------------------------------------------- generic pipeline
__device__ __forceinline__ int find_segment(const int32_t* __restrict__ seg_offsets, int n_seg, int p) {
  int lo = 0, hi = n_seg;  // largest s with seg_offsets[s] <= p
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (seg_offsets[mid] <= p) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
nms_prepare_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                   const int64_t* __restrict__ groups, const uint8_t* __restrict__ valid,
                   const int32_t* __restrict__ seg_offsets, int n_seg, int n_total, int n_pow2,
                   int offset_mode, unsigned long long* __restrict__ keys,
                   uint32_t* __restrict__ vals, uint32_t* __restrict__ seg_max,
                   int32_t* __restrict__ seg_live /* optional: live records per segment */,
                   int32_t* __restrict__ status) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pow2) return;     // n_pow2 is a multiple of the warp size or a single partial warp exits together
  unsigned long long key = kKeyMax;
  int live_seg = -1;
  if (p < n_total && (!valid || valid[p])) {
    const int seg = find_segment(seg_offsets, n_seg, p);
    long long g = groups ? groups[p] : 0;
    if (!offset_mode && (g < 0 || g > 65535)) { atomicOr(status, 1); g = 0; }
    if (offset_mode) {
      float4 b = ld_box(boxes, p);
      float m = fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w));  // boxes.max(), TV ops/boxes.py:99
      atomicMax(&seg_max[seg], float_ordered(m));
      g = 0;  // one run per segment; the groups act through the coordinate shift only
    }
    key = ((unsigned long long)seg << 48) | ((unsigned long long)g << 32) |
          (unsigned long long)(~float_ordered(scores[p] + 0.f));  // -0 sorts like +0
    live_seg = seg;
  }
  if (seg_live) {   // warp-aggregated count of the live records per segment
    const unsigned peers = __match_any_sync(__activemask(), live_seg);
    if (live_seg >= 0 && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&seg_live[live_seg], __popc(peers));
  }
  keys[p] = key;
  vals[p] = (uint32_t)p;
}

__global__ void __launch_bounds__(256)
nms_gather_kernel(const float* __restrict__ boxes, const int64_t* __restrict__ groups,
                  const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals,
                  const uint32_t* __restrict__ seg_max, int n_pow2, int offset_mode,
                  float4* __restrict__ sbox, uint32_t* __restrict__ runkey) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pow2) return;
  const unsigned long long key = keys[p];
  if (key == kKeyMax) { runkey[p] = kNoRun; return; }
  const uint32_t idx = vals[p];
  float4 b = ld_box(boxes, idx);
  if (offset_mode) {
    const int seg = (int)(key >> 48);
    const float maxc = float_from_ordered(seg_max[seg]);
    const float g = groups ? (float)groups[idx] : 0.f;           // idxs.to(boxes)
    const float off = __fmul_rn(g, __fadd_rn(maxc, 1.f));        // TV ops/boxes.py:100
    b.x = __fadd_rn(b.x, off); b.y = __fadd_rn(b.y, off);        // TV ops/boxes.py:101
    b.z = __fadd_rn(b.z, off); b.w = __fadd_rn(b.w, off);
  }
  sbox[p] = b;
  runkey[p] = (uint32_t)(key >> 32);
}

__global__ void __launch_bounds__(256)
nms_rekey_kernel(unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ keepbits,
                 int n_pow2, int32_t* __restrict__ seg_live /* optional: kept records per segment */) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pow2) return;
  const bool kept = (keepbits[p >> 6] >> (p & 63)) & 1ull;
  const unsigned long long key = keys[p];
  const bool live = kept && key != kKeyMax;
  keys[p] = live ? (key & ~(0xffffull << 32)) : kKeyMax;
  if (seg_live) {
    const int live_seg = live ? (int)(key >> 48) : -1;
    const unsigned peers = __match_any_sync(__activemask(), live_seg);
    if (live && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&seg_live[live_seg], __popc(peers));
  }
}

__global__ void __launch_bounds__(256)
nms_output_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals,
                  const int32_t* __restrict__ seg_offsets, int n_pow2, int out_stride,
                  int64_t* __restrict__ keep_out, int32_t* __restrict__ keep_count) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pow2) return;
  const unsigned long long key = keys[p];
  if (key == kKeyMax) return;
  const int seg = (int)(key >> 48);
  const unsigned long long first_key = (unsigned long long)seg << 48;
  int lo = 0, hi = p;  // lower bound of first_key in keys[0..p]
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (keys[mid] < first_key) lo = mid + 1; else hi = mid;
  }
  const int rank = p - lo;
  if (rank < out_stride)
    keep_out[(size_t)seg * out_stride + rank] = (int64_t)vals[p] - seg_offsets[seg];
  const unsigned long long next = (p + 1 < n_pow2) ? keys[p + 1] : kKeyMax;
  if (next == kKeyMax || (int)(next >> 48) != seg) keep_count[seg] = min(rank + 1, out_stride);
}

struct NmsBuffers {
  unsigned long long* keys;
  uint32_t* vals;
  unsigned long long* keys2;
  uint32_t* vals2;
  float4* sbox;
  uint32_t* runkey;
  unsigned long long* mask;
  unsigned long long* keepbits;
  unsigned long long* diag_cols;
  uint32_t* seg_max;
  int32_t* seg_live;
};

static size_t carve(Workspace& ws, NmsBuffers& b, int n_total, int n_seg, int max_seg_len) {
  const int P = next_pow2(n_total > 0 ? n_total : 1);
  b.keys = ws.take<unsigned long long>(P);
  b.vals = ws.take<uint32_t>(P);
  b.keys2 = ws.take<unsigned long long>(P);    // out-of-place buffers of the rank sort
  b.vals2 = ws.take<uint32_t>(P);
  b.sbox = ws.take<float4>(P);
  b.runkey = ws.take<uint32_t>(P);
  b.keepbits = ws.take<unsigned long long>(P / 64 + 1);
  b.diag_cols = ws.take<unsigned long long>(P);
  b.seg_max = ws.take<uint32_t>(n_seg > 0 ? n_seg : 1);
  b.seg_live = ws.take<int32_t>(2 * (size_t)(n_seg > 0 ? n_seg : 1));   // live per segment: first / second sort
  b.mask = ws.take<unsigned long long>((size_t)(n_total > 0 ? n_total : 1) * nms_mask_row_words(max_seg_len));
  return ws.used;
}

}  // namespace dgod

using namespace dgod;

extern "C" size_t dgod_nms_workspace_bytes(int n_total, int n_seg, int max_seg_len) {
  Workspace ws(nullptr, 0);
  NmsBuffers b;
  return carve(ws, b, n_total, n_seg, max_seg_len);
}

extern "C" int dgod_nms_batched(const float* boxes, const float* scores, const int64_t* groups,
                                const uint8_t* valid, const int32_t* seg_offsets, int n_seg,
                                int n_total, int max_seg_len, double iou_threshold,
                                int offset_mode, int max_out_per_seg, int64_t* keep_out,
                                int32_t* keep_count, int32_t* status, void* workspace,
                                size_t workspace_bytes, dgod_stream_t stream) {
  DGOD_REQUIRE(n_seg >= 0 && n_total >= 0 && max_seg_len >= 0, "dgod_nms_batched: negative size");
  DGOD_REQUIRE(n_seg < 65535, "dgod_nms_batched: at most 65534 segments per call");
  DGOD_REQUIRE(max_seg_len <= n_total, "dgod_nms_batched: max_seg_len > n_total");
  if (n_seg == 0) return DGOD_OK;
  DGOD_REQUIRE(keep_count && status && seg_offsets, "dgod_nms_batched: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  DGOD_CUDA(cudaMemsetAsync(keep_count, 0, (size_t)n_seg * sizeof(int32_t), st));
  DGOD_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
  if (n_total == 0) return DGOD_OK;
  DGOD_REQUIRE(boxes && scores && keep_out, "dgod_nms_batched: null pointer");
  Workspace ws(workspace, workspace_bytes);
  NmsBuffers b;
  carve(ws, b, n_total, n_seg, max_seg_len);
  if (!workspace || !ws.ok()) {
    set_error("dgod_nms_batched: workspace too small (%zu < %zu)", workspace_bytes, ws.used);
    return DGOD_ERR_WORKSPACE;
  }
  const int P = next_pow2(n_total);
  const int out_stride = max_out_per_seg > 0 ? max_out_per_seg : max_seg_len;
  const float thr = float_round_down(iou_threshold);
  DGOD_CUDA(cudaMemsetAsync(b.keepbits, 0, (size_t)(P / 64 + 1) * sizeof(unsigned long long), st));
  if (offset_mode) DGOD_CUDA(cudaMemsetAsync(b.seg_max, 0, (size_t)n_seg * sizeof(uint32_t), st));

  // segments of a training step (4-8 k candidates) are ordered by one rank-sort launch; very long
  // segments take the bitonic network
  const bool use_rank = max_seg_len <= kRankSortMaxSeg && n_seg <= kRankSortMaxNseg;
  if (use_rank) DGOD_CUDA(cudaMemsetAsync(b.seg_live, 0, 2 * (size_t)n_seg * sizeof(int32_t), st));
  nms_prepare_kernel<<<cdiv(P, 256), 256, 0, st>>>(boxes, scores, groups, valid, seg_offsets, n_seg,
                                                   n_total, P, offset_mode, b.keys, b.vals,
                                                   b.seg_max, use_rank ? b.seg_live : nullptr, status);
  DGOD_LAUNCHED();
  int rc;
  if (use_rank) {
    rc = rank_sort(b.keys, b.vals, seg_offsets, nullptr, b.seg_live, n_seg, n_total, P, max_seg_len, b.keys2, b.vals2, st);
    std::swap(b.keys, b.keys2);
    std::swap(b.vals, b.vals2);
  } else {
    rc = bitonic_sort(b.keys, b.vals, P, st);
  }
  if (rc) return rc;
  nms_gather_kernel<<<cdiv(P, 256), 256, 0, st>>>(boxes, groups, b.keys, b.vals, b.seg_max, P,
                                                  offset_mode, b.sbox, b.runkey);
  DGOD_LAUNCHED();
  // positions >= n_total are padding (kNoRun) and sort to the tail: only n_total positions matter
  rc = launch_nms_mask(b.sbox, b.runkey, n_total, max_seg_len, thr, b.mask, b.diag_cols, st);
  if (rc) return rc;
  rc = launch_nms_scan(b.mask, b.diag_cols, b.runkey, nullptr, n_total, max_seg_len, b.keepbits, nullptr, nullptr, st);
  if (rc) return rc;
  nms_rekey_kernel<<<cdiv(P, 256), 256, 0, st>>>(b.keys, b.keepbits, P, use_rank ? b.seg_live + n_seg : nullptr);
  DGOD_LAUNCHED();
  if (use_rank) {
    // the input is the first sort's output: segment s now lives in the live range of s
    rc = rank_sort(b.keys, b.vals, seg_offsets, b.seg_live, b.seg_live + n_seg, n_seg, n_total, P, max_seg_len, b.keys2, b.vals2, st);
    std::swap(b.keys, b.keys2);
    std::swap(b.vals, b.vals2);
  } else {
    rc = bitonic_sort(b.keys, b.vals, P, st);
  }
  if (rc) return rc;
  nms_output_kernel<<<cdiv(P, 256), 256, 0, st>>>(b.keys, b.vals, seg_offsets, P, out_stride,
                                                  keep_out, keep_count);
  DGOD_LAUNCHED();
  return DGOD_OK;
}
