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
//   1. order     one CTA per segment sorts (group, descending score, index) records in shared memory and
//                writes the boxes in processing order (+ the offset trick's fp32 shift) and 32-bit run
//                keys; runs = (segment, group)
//   2. mask      64x64 IoU tiles from shared memory -> per-row 64-bit suppression words, only
//                inside a run and above the diagonal (block-diagonal, upper-triangular)
//   3. scan      one CTA per run: resolve 64 candidates per step from the diagonal word, OR the
//                kept rows into the run's removed-bitmap in shared memory
//   4. emit      one CTA per segment sorts the kept records by (descending score, index) and writes
//                segment-relative indices, truncated to max_out_per_seg
// Segments longer than 16384 records take a global bitonic network instead of steps 1 and 4
// (prepare, sort, gather | re-key, sort, output).
//
// Bytes: 28 B per box in and 8 B per kept box out are compulsory; the mask adds
// 2 * 8 B * sum_runs n_r^2/128 (write + read).  The pair evaluations (sum_runs n_r^2/2, ~25
// fp32 instructions each) are the real cost for large runs — see DESIGN.md for the roofline.
#include "nms_core.cuh"
#include "bulk.cuh"
#include "sort.cuh"
#include <cooperative_groups.h>
#include <utility>
namespace cg = cooperative_groups;

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

__device__ __forceinline__ bool iou_gt_exact(const float4 a, float area_a, const float4 b, float area_b, float thr,
                                             bool skip_disjoint) {
  const float w = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
  const float h = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
  if (skip_disjoint && (w <= 0.f || h <= 0.f)) return false;   // IoU is 0 (or 0/0): never > thr >= 0
  const float inter = __fmul_rn(fmaxf(w, 0.f), fmaxf(h, 0.f));
  const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  if (uni > 0.f && thr > 0.f) {
    const float t = __fmul_rn(thr, uni);
    if (inter > __fmul_rn(t, 1.00000095367431640625f)) return true;     // 1 + 2^-20
    if (inter < __fmul_rn(t, 0.99999904632568359375f)) return false;    // 1 - 2^-20
  }
  return __fdiv_rn(inter, uni) > thr;
}

// The same decision without the early exit for disjoint pairs.  A warp holds 32 rows against ONE column box, so the
// exit only pays when none of the 32 pairs overlaps; with the clustered boxes of a detector nearly every iteration took
// both the exit test and the full path.  Here every pair runs ~17 arithmetic instructions and only the borderline /
// degenerate ones (a handful per million) branch to the IEEE division.  Bit-identical decisions: a disjoint pair has
// inter = 0 < t(1 - 2^-20) when union and threshold are positive, and takes the quotient (0 or NaN, never > thr)
// otherwise.
#ifndef DGOD_NMS_MASK_EARLY_EXIT
__device__ __forceinline__ bool iou_gt_exact_flat(const float4 a, float area_a, const float4 b, float area_b, float thr) {
  const float w = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
  const float h = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
  const float inter = __fmul_rn(fmaxf(w, 0.f), fmaxf(h, 0.f));
  const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  const float t = __fmul_rn(thr, uni);
  const bool hit = inter > __fmul_rn(t, 1.00000095367431640625f);       // 1 + 2^-20
  const bool miss = inter < __fmul_rn(t, 0.99999904632568359375f);      // 1 - 2^-20
  if ((hit || miss) && uni > 0.f && thr > 0.f) return hit;
  return __fdiv_rn(inter, uni) > thr;
}
#define IOU_GT(a, aa, b, ab, thr, skip) iou_gt_exact_flat(a, aa, b, ab, thr)
#else
#define IOU_GT(a, aa, b, ab, thr, skip) iou_gt_exact(a, aa, b, ab, thr, skip)
#endif

__global__ void __launch_bounds__(256)
nms_mask_kernel(const float4* __restrict__ sbox, const uint32_t* __restrict__ runkey, int n_pos,
                float thr, unsigned long long* __restrict__ mask, int row_words,
                unsigned long long* __restrict__ diag_cols, int pdl) {
  __shared__ float4 s_box[kMaskChunksPerCta][64];
  __shared__ float s_area[kMaskChunksPerCta][64];
  __shared__ uint32_t s_key[kMaskChunksPerCta][64];
  __shared__ uint32_t s_colbits[64][2];
  if (pdl) {                                                  // launched behind the kernel that wrote sbox / runkey
    pdl_wait();
    pdl_trigger();
  }
  const int r = blockIdx.x;                                   // row chunk
  const int c0 = r + kMaskChunksPerCta * blockIdx.y;          // first column chunk of this CTA
  const int row0 = r * 64;
  const long long col0 = (long long)c0 * 64;
  if (row0 >= n_pos || col0 >= n_pos) return;
  // run keys are non-decreasing: if the first column is already past the run of the chunk's last
  // real row, no pair of this CTA shares a run.  Dead rows never own pairs.
  int last = min(row0 + 63, n_pos - 1);
  uint32_t rk_last = runkey[last];
  while (is_norun(rk_last) && last > row0) rk_last = runkey[--last];
  if (is_norun(rk_last) || runkey[col0] > rk_last) {
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
    s_area[cc][b] = box_area_exact(bx.x, bx.y, bx.z, bx.w);
    s_key[cc][b] = k;
  }
  __syncthreads();

  const int cc = threadIdx.x >> 6, row = threadIdx.x & 63;
  const int p = row0 + row;
  const int w = c0 + cc - r;                                  // word index inside the mask row
  const bool diag = blockIdx.y == 0 && cc == 0;               // warps 0-1 of the first CTA: the diagonal tile
  const bool active = p < n_pos && w < row_words;
  const uint32_t rk = active ? runkey[p] : kNoRun;
  const bool dead = is_norun(rk);
  if (!diag && dead) return;
  unsigned long long bits = 0ull;
  const long long qbase = col0 + cc * 64;
  if (!dead && qbase + 63 > p && s_key[cc][0] <= rk && s_key[cc][63] >= rk) {
    const float4 a = sbox[p];
    const float area_a = box_area_exact(a.x, a.y, a.z, a.w);
    const bool skip_disjoint = thr >= 0.f;
    const int b_lo = qbase > p ? 0 : (int)(p - qbase) + 1;   // only columns behind the row
    if (s_key[cc][0] == rk && s_key[cc][63] == rk) {
      // the whole column chunk lies inside the row's run (all but the chunks at the ends of a run)
#pragma unroll 4
      for (int b = b_lo; b < 64; ++b)
        if (IOU_GT(a, area_a, s_box[cc][b], s_area[cc][b], thr, skip_disjoint)) bits |= (1ull << b);
    } else {
#pragma unroll 4
      for (int b = b_lo; b < 64; ++b)
        if (s_key[cc][b] == rk && IOU_GT(a, area_a, s_box[cc][b], s_area[cc][b], thr, skip_disjoint))
          bits |= (1ull << b);
    }
  }
  if (!dead) mask[((size_t)r * row_words + w) * 64 + row] = bits;        // chunk-major: [chunk][word][row]
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
  // Programmatic dependent launch pays for grids of a few waves (it hides the launch gap, ~4 us); griddepcontrol.wait
  // itself costs every CTA ~2 us, which a grid of a hundred waves (100 000 boxes) would pay a hundred times.
  if ((long long)grid.x * grid.y <= 16384) {
    DGOD_CUDA(launch_pdl(nms_mask_kernel, grid, dim3(256), 0, st, sbox, runkey, n_pos, thr, mask, row_words, diag_cols, 1));
  } else {
    nms_mask_kernel<<<grid, 256, 0, st>>>(sbox, runkey, n_pos, thr, mask, row_words, diag_cols, 0);
  }
  DGOD_LAUNCHED();
  return DGOD_OK;
}

// ---------------------------------------------------------------------------- scan
// One CTA per run (the CTA of the chunk in which the run starts).  A run is resolved 64 candidates
// (one chunk) at a time, in order; the only true dependency between consecutive chunks is
//     removed(c+1) = OR of the mask word for chunk c+1 over every kept row of chunks <= c,
// so the kernel keeps exactly that on its critical path:
//   * the chunks' mask rows (64 rows x row_words, contiguous in memory) stream through a ring of shared-memory
//     buffers, one bulk copy (TMA engine) per chunk, issued two to three chunks ahead;
//   * warp 0 resolves chunk c (fixpoint on the transposed diagonal block), then ORs the kept rows' word for
//     chunk c+1 only into a register ("carry") — diag_cols / alive of chunk c+1 were fetched a step earlier;
//   * the other warps, one step behind, OR the rows kept in chunk c-1 into the removed-bitmap words of the
//     chunks >= c+1.
// One __syncthreads per chunk.  Runs whose rows do not fit the ring (> ~8 k candidates) take the older kernel
// below, which reads the kept rows from global memory.
constexpr int kScanThreads = 256;
constexpr int kScanThreadsLong = 1024;   // runs whose mask rows do not fit shared memory
constexpr int kScanRingBytes = 200 * 1024;
constexpr int kScanDepthMax = 8;
constexpr int kScanMaxWords = kScanRingBytes / 3 / (64 * 8);   // ring of >= 3 chunks: runs up to ~8 k candidates

__global__ void __launch_bounds__(kScanThreads)
nms_scan_kernel(const unsigned long long* __restrict__ mask, const unsigned long long* __restrict__ diag_cols,
                int row_words, int depth, const uint32_t* __restrict__ runkey, const uint8_t* __restrict__ alive,
                int n_pos, unsigned long long* __restrict__ keepbits, int32_t* __restrict__ compact_pos,
                int32_t* __restrict__ run_count) {
  pdl_wait();                                                 // launched behind the mask kernel
  pdl_trigger();
  extern __shared__ __align__(128) unsigned long long s_dyn[];     // ring[depth][64][row_words] | removed[row_words]
  const size_t buf_words = (size_t)64 * row_words;
  unsigned long long* s_removed = s_dyn + (size_t)depth * buf_words;
  __shared__ __align__(8) unsigned long long s_full[kScanDepthMax];
  __shared__ uint32_t s_ballot[2];
  __shared__ unsigned long long s_kept[2];
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int p0 = c * 64;
  if (tid < 64) {
    const int p = p0 + tid;
    bool start = false;
    if (p < n_pos) {
      const uint32_t k = runkey[p];
      start = !is_norun(k) && (p == 0 || runkey[p - 1] != k);
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, start);
    if (lane == 0) s_ballot[tid >> 5] = bal;
  }
  if (tid == 0) {
    for (int i = 0; i < kScanDepthMax; ++i) mbar_init(&s_full[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  unsigned long long starts = ((unsigned long long)s_ballot[1] << 32) | s_ballot[0];
  if (starts == 0ull) return;
  const unsigned buf_bytes = (unsigned)(buf_words * sizeof(unsigned long long));
  unsigned issued = 0;       // chunks this CTA has streamed so far (all runs): chunk j sits in slot j % depth
  unsigned waits[kScanDepthMax] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};   // completed phases of every slot's barrier (uniform)
  auto phase_of = [&](unsigned slot) -> unsigned {
    unsigned ph = 0u;
#pragma unroll
    for (int d = 0; d < kScanDepthMax; ++d)
      if ((unsigned)d == slot) { ph = waits[d] & 1u; ++waits[d]; }
    return ph;
  };

  while (starts) {
    const int sb = __ffsll((long long)starts) - 1;
    starts &= starts - 1;
    const int s = p0 + sb;
    const uint32_t rk = runkey[s];
    int lo = s + 1, hi = n_pos;  // upper bound of rk in the non-decreasing key array
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (runkey[mid] <= rk) lo = mid + 1; else hi = mid;
    }
    const int e = lo;
    const int cs = s >> 6, ce = (e - 1) >> 6, n_chunks = ce - cs + 1;
    __syncthreads();  // previous run done with the shared buffers
    for (int w = tid; w < n_chunks && w < row_words; w += blockDim.x) s_removed[w] = 0ull;
    const unsigned base = issued;
    auto stream_chunk = [&](int i) {      // one thread: words 1 .. n_chunks-1-i of chunk cs + i (the columns inside the run)
      const unsigned j = base + (unsigned)i;
      const unsigned slot = j % (unsigned)depth;
      const unsigned bytes = (unsigned)(n_chunks - 1 - i) * 512u;
      if (bytes == 0u) return;             // the run's last chunk has no columns behind it
      mbar_expect_tx(&s_full[slot], bytes);
      bulk_load(s_dyn + (size_t)slot * buf_words + 64, mask + ((size_t)(cs + i) * row_words + 1) * 64, bytes, &s_full[slot]);
    };
    if (tid == 32)
      for (int i = 0; i < depth && i < n_chunks; ++i) stream_chunk(i);
    issued += (unsigned)n_chunks;

    // warp 0 state: inputs of the chunk it resolves next, fetched one step ahead
    unsigned long long col_a = 0ull, col_b = 0ull, carry = 0ull;
    bool in_a = false, in_b = false;
    auto fetch_inputs = [&](int cc) {
      const int q0 = cc * 64 + lane, q1 = q0 + 32;
      in_a = q0 >= s && q0 < e && (!alive || alive[q0]);
      in_b = q1 >= s && q1 < e && (!alive || alive[q1]);
      col_a = q0 < n_pos ? diag_cols[q0] : 0ull;   // only rows a < b of the same run
      col_b = q1 < n_pos ? diag_cols[q1] : 0ull;
    };
    if (tid < 32) fetch_inputs(cs);
    int count = 0;
    for (int i = 0; i < n_chunks; ++i) {
      const unsigned j = base + (unsigned)i;
      const unsigned slot = j % (unsigned)depth;
      __syncthreads();   // step i-1 finished everywhere: removed[] holds the rows kept up to chunk i-2, kept[i-1] is published
      if (tid == 32 && i >= 2 && i + depth - 2 < n_chunks) stream_chunk(i + depth - 2);   // slot of chunk i-2 is free
      if (i + 1 < n_chunks) mbar_wait(&s_full[slot], phase_of(slot));
      const unsigned long long* rows = s_dyn + (size_t)slot * buf_words;      // [word][row] of chunk i
      if (tid < 32) {
        // Greedy keep set of the chunk as the fixpoint of
        //   K[b] = alive[b] and no a < b with K[a] and D[a][b]
        // iterated from K = alive: entry b is final after b+1 sweeps at the latest, in practice
        // after (longest suppression chain + 1) sweeps; each sweep is two ballots.
        const unsigned long long al =
            (((unsigned long long)__ballot_sync(0xffffffffu, in_b) << 32) | __ballot_sync(0xffffffffu, in_a)) &
            ~(s_removed[i] | carry);
        const unsigned long long ca = col_a, cb = col_b;
        if (i + 1 < n_chunks) fetch_inputs(cs + i + 1);        // global loads in flight during the resolve
        unsigned long long K = al;
        while (true) {
          const bool k0 = ((al >> lane) & 1ull) && ((ca & K) == 0ull);
          const bool k1 = ((al >> (lane + 32)) & 1ull) && ((cb & K) == 0ull);
          const unsigned long long Kn = ((unsigned long long)__ballot_sync(0xffffffffu, k1) << 32) | __ballot_sync(0xffffffffu, k0);
          if (Kn == K) break;
          K = Kn;
        }
        if (lane == 0) s_kept[i & 1] = K;
        carry = 0ull;
        if (i + 1 < n_chunks) {          // word 1 of the chunk's rows = columns of chunk i+1
          unsigned long long v = 0ull;
          if ((K >> lane) & 1ull) v = rows[64 + lane];
          if ((K >> (lane + 32)) & 1ull) v |= rows[64 + lane + 32];
          const unsigned vlo = __reduce_or_sync(0xffffffffu, (unsigned)v);
          const unsigned vhi = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
          carry = ((unsigned long long)vhi << 32) | vlo;
        }
        if (compact_pos) {
          if ((K >> lane) & 1ull) compact_pos[s + count + __popcll(K & ((1ull << lane) - 1ull))] = (cs + i) * 64 + lane;
          if ((K >> (lane + 32)) & 1ull) compact_pos[s + count + __popcll(K & ((1ull << (lane + 32)) - 1ull))] = (cs + i) * 64 + lane + 32;
        }
        if (keepbits && lane == 0 && K) atomicOr(&keepbits[cs + i], K);
        count += __popcll(K);
      } else if (i >= 1 && i + 1 < n_chunks) {
        // rows kept in chunk i-1 -> removed[] of the chunks >= i+1 (their words >= 2): one warp per word, a lane ORs
        // its two rows, warp-wide OR reduction
        const unsigned long long kept = s_kept[(i - 1) & 1];
        if (kept) {
          const unsigned pslot = (j - 1u) % (unsigned)depth;
          const unsigned long long* prow = s_dyn + (size_t)pslot * buf_words;
          const bool ka = (kept >> lane) & 1ull, kb = (kept >> (lane + 32)) & 1ull;
          const int n_w = n_chunks - i - 1;                    // words 2 .. n_w + 1 of chunk i-1's rows
          for (int w = (tid >> 5) - 1; w < n_w; w += kScanThreads / 32 - 1) {
            unsigned long long v = 0ull;
            if (ka) v = prow[(size_t)(w + 2) * 64 + lane];
            if (kb) v |= prow[(size_t)(w + 2) * 64 + lane + 32];
            const unsigned vlo = __reduce_or_sync(0xffffffffu, (unsigned)v);
            const unsigned vhi = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
            if (lane == 0) s_removed[i + 1 + w] |= ((unsigned long long)vhi << 32) | vlo;
          }
        }
      }
    }
    if (run_count && tid == 0) run_count[rk] = count;
  }
}

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}

__global__ void __launch_bounds__(kScanThreadsLong)
nms_scan_long_kernel(const unsigned long long* __restrict__ mask, const unsigned long long* __restrict__ diag_cols,
                     int row_words, const uint32_t* __restrict__ runkey, const uint8_t* __restrict__ alive, int n_pos,
                     unsigned long long* __restrict__ keepbits, int32_t* __restrict__ compact_pos,
                     int32_t* __restrict__ run_count) {
  extern __shared__ unsigned long long s_dyn[];     // removed[row_words]
  unsigned long long* s_removed = s_dyn;
  __shared__ uint32_t s_ballot[2];
  __shared__ unsigned long long s_kept;
  const int c = blockIdx.x, tid = threadIdx.x;
  const int p0 = c * 64;
  if (tid < 64) {
    const int p = p0 + tid;
    bool start = false;
    if (p < n_pos) {
      uint32_t k = runkey[p];
      start = !is_norun(k) && (p == 0 || runkey[p - 1] != k);
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
    int count = 0;
    for (int cc = cs; cc <= ce; ++cc) {
      bool in = false;
      if (tid < 64) {
        const int p = cc * 64 + tid;
        in = p >= s && p < e && (!alive || alive[p]);
        uint32_t bal = __ballot_sync(0xffffffffu, in);
        if ((tid & 31) == 0) s_ballot[tid >> 5] = bal;
      }
      __syncthreads();                                   // removed[] is current
      if (tid < 32) {
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
      // the rows stay in global memory (L2).  8 threads per word, each with its up to 8 kept rows
      // in flight at once, then a 3-step shuffle OR
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
              v[u] = mask[((size_t)cc * row_words + w) * 64 + b];
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
      if (tid < 64 && ((kept >> tid) & 1ull) && compact_pos) {
        const int j = count + __popcll(kept & ((1ull << tid) - 1ull));
        compact_pos[s + j] = cc * 64 + tid;
      }
      if (keepbits && tid == 0 && kept) atomicOr(&keepbits[cc], kept);
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
  if (row_words <= kScanMaxWords) {
    const size_t buf = (size_t)64 * row_words * sizeof(unsigned long long);
    int depth = (int)((size_t)(100 * 1024) / buf);       // ~100 KB of rows in flight, two CTAs per SM
    depth = depth < 3 ? 3 : (depth > kScanDepthMax ? kScanDepthMax : depth);
    const size_t smem = buf * depth + (size_t)row_words * sizeof(unsigned long long);
    static size_t attr_smem = 0;
    if (smem > 48 * 1024 && smem > attr_smem) {
      DGOD_CUDA(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_smem = smem;
    }
    DGOD_CUDA(launch_pdl(nms_scan_kernel, dim3(cdiv(n_pos, 64)), dim3(kScanThreads), smem, st, mask, diag_cols, row_words, depth, runkey,
                         alive, n_pos, keepbits, compact_pos, run_count));
  } else {
    const size_t smem = (size_t)row_words * sizeof(unsigned long long);
    static size_t attr_smem_long = 0;
    if (smem > 48 * 1024 && smem > attr_smem_long) {
      DGOD_CUDA(cudaFuncSetAttribute(nms_scan_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_smem_long = smem;
    }
    // (a plain launch: behind a mask kernel of many waves, early-scheduled 1024-thread CTAs cost more than the launch gap)
    nms_scan_long_kernel<<<cdiv(n_pos, 64), kScanThreadsLong, smem, st>>>(mask, diag_cols, row_words, runkey, alive, n_pos,
                                                                        keepbits, compact_pos, run_count);
  }
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
                   int32_t* __restrict__ status) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pow2) return;     // n_pow2 is a multiple of the warp size or a single partial warp exits together
  unsigned long long key = kKeyMax;
  if (p < n_total && (!valid || valid[p])) {
    const int seg = find_segment(seg_offsets, n_seg, p);
    long long g = groups ? groups[p] : 0;
    if (!offset_mode && (g < 0 || g > 65534)) { atomicOr(status, 1); g = 0; }
    if (offset_mode) {
      float4 b = ld_box(boxes, p);
      float m = fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w));  // boxes.max(), TV ops/boxes.py:99
      atomicMax(&seg_max[seg], float_ordered(m));
      g = 0;  // one run per segment; the groups act through the coordinate shift only
    }
    key = ((unsigned long long)seg << 48) | ((unsigned long long)g << 32) |
          (unsigned long long)(~float_ordered(scores[p] + 0.f));  // -0 sorts like +0
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
nms_rekey_kernel(unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ keepbits, int n_pow2) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pow2) return;
  const bool kept = (keepbits[p >> 6] >> (p & 63)) & 1ull;
  const unsigned long long key = keys[p];
  const bool live = kept && key != kKeyMax;
  keys[p] = live ? (key & ~(0xffffull << 32)) : kKeyMax;
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

// ---------------------------------------------------------------------------- segment-sort pipeline
// Segments of up to 16384 records (every NMS call of a training / evaluation step) are ordered by ONE 8-CTA
// cluster each, in shared memory: a record is the 64-bit word
//     group:16 | ~ordered(score):32 | index inside the segment:14        (dead records: ~0)
// whose ascending order is the processing order (group, descending score, ascending index).  Each CTA sorts an
// eighth of the segment (bitonic network in its own shared memory), pulls the seven other sorted tiles through
// distributed shared memory and ranks its records against them by binary search — this replaces prepare + sort +
// gather.  Positions keep the segment's own range [seg_offsets[s], seg_offsets[s+1]): live records first, dead
// ones behind them with the run key dead_key(s).  After mask + scan a second kernel of the same shape orders the
// kept records by (descending score, index) and writes the output — 4 launches and no memset for the whole call.
constexpr int kSegSortMax = 16384;
constexpr int kSegCluster = 8;
constexpr int kSegSortThreads = 1024;

__device__ __forceinline__ void bitonic_sort_shared(unsigned long long* sk, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int q = threadIdx.x; q < (n_pow2 >> 1); q += blockDim.x) {
        const int i = 2 * q - (q & (j - 1));
        const int l = i + j;
        const bool asc = (i & k) == 0;
        const unsigned long long a = sk[i], b = sk[l];
        if ((a > b) == asc) { sk[i] = b; sk[l] = a; }
      }
      __syncthreads();
    }
  }
}

// Tile size of a segment of n records: a power of two >= 32 with kSegCluster * T >= n.
__host__ __device__ __forceinline__ int seg_tile(int n) {
  int t = 32;
  while (t * kSegCluster < n) t <<= 1;
  return t;
}

// The CTA's tile s_all[rank*T, +T) holds its records.  Sorts it, exchanges tiles across the cluster; on return every
// CTA holds all kSegCluster sorted tiles in s_all.
__device__ __forceinline__ void cluster_sort_tiles(cg::cluster_group& cluster, unsigned long long* s_all, int T, int rank) {
  __syncthreads();
  bitonic_sort_shared(s_all + rank * T, T);
  cluster.sync();                                        // every tile is sorted
  for (int r = 0; r < kSegCluster; ++r) {
    if (r == rank) continue;
    const unsigned long long* remote = cluster.map_shared_rank(s_all, r) + r * T;
    for (int i = threadIdx.x; i < T; i += blockDim.x) s_all[r * T + i] = remote[i];
  }
  __syncthreads();
}

// Global rank of the record at index i of the CTA's own tile: records of equal key (only the dead ones) keep the
// tile-major order.
__device__ __forceinline__ int cluster_rank(const unsigned long long* s_all, int T, int rank, int i) {
  const unsigned long long key = s_all[rank * T + i];
  int pos[kSegCluster];
#pragma unroll
  for (int r = 0; r < kSegCluster; ++r) pos[r] = 0;
  for (int st = T >> 1; st > 0; st >>= 1) {
#pragma unroll
    for (int r = 0; r < kSegCluster; ++r) {
      const unsigned long long v = s_all[r * T + pos[r] + st - 1];
      pos[r] += ((r < rank) ? (v <= key) : (v < key)) ? st : 0;
    }
  }
  int total = i;
#pragma unroll
  for (int r = 0; r < kSegCluster; ++r) {
    const unsigned long long v = s_all[r * T + pos[r]];
    const int c = pos[r] + (((r < rank) ? (v <= key) : (v < key)) ? 1 : 0);
    total += (r == rank) ? 0 : c;
  }
  return total;
}

__global__ void __cluster_dims__(kSegCluster, 1, 1) __launch_bounds__(kSegSortThreads)
nms_order_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, const int64_t* __restrict__ groups,
                 const uint8_t* __restrict__ valid, const int32_t* __restrict__ seg_offsets, int offset_mode,
                 unsigned long long* __restrict__ skey, float4* __restrict__ sbox, uint32_t* __restrict__ runkey,
                 unsigned long long* __restrict__ keepbits, int32_t* __restrict__ seg_status) {
  pdl_trigger();                                              // the mask kernel may be scheduled behind this grid
  extern __shared__ __align__(16) unsigned long long s_all[];
  __shared__ uint32_t s_max;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), seg = blockIdx.x / kSegCluster, tid = threadIdx.x;
  const int s0 = seg_offsets[seg], n = seg_offsets[seg + 1] - s0;
  const int T = seg_tile(n);
  if (tid == 0) s_max = 0u;
  __syncthreads();
  bool bad = false;
  for (int t = tid; t < T; t += blockDim.x) {
    const int i = rank * T + t;
    unsigned long long ck = kKeyMax;
    if (i < n && (!valid || valid[s0 + i])) {
      long long g = (groups && !offset_mode) ? groups[s0 + i] : 0;   // offset mode: one run per segment, the groups act
      if (g < 0 || g > 65534) {                                      // through the coordinate shift only
        bad = true;
        g = 0;
      }
      ck = ((((unsigned long long)g << 32) | (unsigned long long)(~float_ordered(scores[s0 + i] + 0.f))) << 14) |
           (unsigned long long)i;                                                      // -0 sorts like +0
    }
    s_all[i] = ck;
  }
  if (offset_mode) {     // boxes.max() over the segment's valid boxes, TV ops/boxes.py:99 (every CTA of the cluster)
    uint32_t my_max = 0u;
    for (int i = tid; i < n; i += blockDim.x) {
      if (valid && !valid[s0 + i]) continue;
      const float4 b = ld_box(boxes, s0 + i);
      my_max = max(my_max, float_ordered(fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w))));
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) my_max = max(my_max, __shfl_xor_sync(0xffffffffu, my_max, d));
    if ((tid & 31) == 0 && my_max) atomicMax(&s_max, my_max);
  }
  bad = __syncthreads_or(bad);
  if (tid == 0) seg_status[seg * kSegCluster + rank] = bad ? 1 : 0;
  cluster_sort_tiles(cluster, s_all, T, rank);
  const float maxc = offset_mode ? float_from_ordered(s_max) : 0.f;
  for (int t = tid; t < T; t += blockDim.x) {
    const int j = cluster_rank(s_all, T, rank, t);
    if (j >= n) continue;                                 // padding behind the segment
    const unsigned long long ck = s_all[rank * T + t];
    const int pos = s0 + j;
    skey[pos] = ck;
    if (ck == kKeyMax) {
      runkey[pos] = dead_key(seg);
      sbox[pos] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    const int idx = s0 + (int)(ck & 0x3fffull);
    float4 b = ld_box(boxes, idx);
    if (offset_mode) {
      const float g = groups ? (float)groups[idx] : 0.f;           // idxs.to(boxes)
      const float off = __fmul_rn(g, __fadd_rn(maxc, 1.f));        // TV ops/boxes.py:100
      b.x = __fadd_rn(b.x, off); b.y = __fadd_rn(b.y, off);        // TV ops/boxes.py:101
      b.z = __fadd_rn(b.z, off); b.w = __fadd_rn(b.w, off);
    }
    sbox[pos] = b;
    runkey[pos] = ((uint32_t)seg << 16) | (uint32_t)((ck >> 46) & 0xffffull);
  }
  // the keep bitmap words that start inside this segment (the scan ORs into them)
  if (rank == 0)
    for (int w = (s0 + 63) / 64 + tid; w * 64 < s0 + n; w += blockDim.x) keepbits[w] = 0ull;
  cluster.sync();                                          // no CTA leaves while its shared memory may still be read
}

__global__ void __cluster_dims__(kSegCluster, 1, 1) __launch_bounds__(kSegSortThreads)
nms_emit_kernel(const unsigned long long* __restrict__ skey, const unsigned long long* __restrict__ keepbits,
                const int32_t* __restrict__ seg_offsets, const int32_t* __restrict__ seg_status, int n_seg,
                int out_stride, int64_t* __restrict__ keep_out, int32_t* __restrict__ keep_count,
                int32_t* __restrict__ status) {
  pdl_wait();                                                 // launched behind the mask kernel
  pdl_trigger();
  extern __shared__ __align__(16) unsigned long long s_all[];
  __shared__ int s_count;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), seg = blockIdx.x / kSegCluster, tid = threadIdx.x;
  const int s0 = seg_offsets[seg], n = seg_offsets[seg + 1] - s0;
  const int T = seg_tile(n);
  if (blockIdx.x == 0) {
    int bad = 0;
    for (int t = tid; t < n_seg * kSegCluster; t += blockDim.x) bad |= seg_status[t];
    bad = __syncthreads_or(bad);
    if (tid == 0) status[0] = bad ? 1 : 0;
  }
  if (tid == 0) s_count = 0;
  for (int t = tid; t < T; t += blockDim.x) {
    const int j = rank * T + t;
    unsigned long long k2 = kKeyMax;
    if (j < n) {
      const int pos = s0 + j;
      const unsigned long long ck = skey[pos];
      if (ck != kKeyMax && ((keepbits[pos >> 6] >> (pos & 63)) & 1ull)) k2 = ck & 0x00003fffffffffffull;   // drop the group
    }
    s_all[j] = k2;
  }
  cluster_sort_tiles(cluster, s_all, T, rank);
  if (tid < kSegCluster) {          // kept records of the segment: the live prefix of every sorted tile
    const unsigned long long* a = s_all + tid * T;
    int pos = 0;
    for (int st = T >> 1; st > 0; st >>= 1) pos += (a[pos + st - 1] != kKeyMax) ? st : 0;
    pos += (a[pos] != kKeyMax) ? 1 : 0;
    atomicAdd(&s_count, pos);
  }
  __syncthreads();
  const int n_out = min(s_count, out_stride);
  for (int t = tid; t < T; t += blockDim.x) {
    const unsigned long long k2 = s_all[rank * T + t];
    if (k2 == kKeyMax) continue;
    const int j = cluster_rank(s_all, T, rank, t);
    if (j < n_out) keep_out[(size_t)seg * out_stride + j] = (int64_t)(k2 & 0x3fffull);
  }
  if (rank == 0 && tid == 0) keep_count[seg] = n_out;
  cluster.sync();                                          // no CTA leaves while its shared memory may still be read
}

struct NmsBuffers {
  unsigned long long* keys;
  uint32_t* vals;
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
  b.sbox = ws.take<float4>(P);
  b.runkey = ws.take<uint32_t>(P);
  b.keepbits = ws.take<unsigned long long>(P / 64 + 1);
  b.diag_cols = ws.take<unsigned long long>(P);
  b.seg_max = ws.take<uint32_t>(n_seg > 0 ? n_seg : 1);
  b.seg_live = ws.take<int32_t>((size_t)(n_seg > 0 ? n_seg : 1) * 8);   // per segment and CTA: a group id was out of range
  b.mask = ws.take<unsigned long long>(nms_mask_rows(n_total > 0 ? n_total : 1) * nms_mask_row_words(max_seg_len));
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
  if (n_total == 0) {
    DGOD_CUDA(cudaMemsetAsync(keep_count, 0, (size_t)n_seg * sizeof(int32_t), st));
    DGOD_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    return DGOD_OK;
  }
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

  if (max_seg_len <= kSegSortMax) {
    // every segment fits one CTA's shared memory: order, mask, scan, emit — 4 launches, no memset
    const size_t smem = (size_t)seg_tile(max_seg_len) * kSegCluster * sizeof(unsigned long long);
    static size_t attr_smem = 0;
    if (smem > 48 * 1024 && smem > attr_smem) {
      DGOD_CUDA(cudaFuncSetAttribute(nms_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      DGOD_CUDA(cudaFuncSetAttribute(nms_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_smem = smem;
    }
    nms_order_kernel<<<n_seg * kSegCluster, kSegSortThreads, smem, st>>>(boxes, scores, groups, valid, seg_offsets, offset_mode, b.keys,
                                                          b.sbox, b.runkey, b.keepbits, b.seg_live);
    DGOD_LAUNCHED();
    int rc = launch_nms_mask(b.sbox, b.runkey, n_total, max_seg_len, thr, b.mask, b.diag_cols, st);
    if (rc) return rc;
    rc = launch_nms_scan(b.mask, b.diag_cols, b.runkey, nullptr, n_total, max_seg_len, b.keepbits, nullptr, nullptr, st);
    if (rc) return rc;
    DGOD_CUDA(launch_pdl(nms_emit_kernel, dim3(n_seg * kSegCluster), dim3(kSegSortThreads), smem, st, (const unsigned long long*)b.keys,
                         (const unsigned long long*)b.keepbits, seg_offsets, (const int32_t*)b.seg_live, n_seg, out_stride, keep_out, keep_count,
                         status));
    DGOD_LAUNCHED();
    return DGOD_OK;
  }

  // very long segments: global bitonic network
  DGOD_CUDA(cudaMemsetAsync(keep_count, 0, (size_t)n_seg * sizeof(int32_t), st));
  DGOD_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
  DGOD_CUDA(cudaMemsetAsync(b.keepbits, 0, (size_t)(P / 64 + 1) * sizeof(unsigned long long), st));
  if (offset_mode) DGOD_CUDA(cudaMemsetAsync(b.seg_max, 0, (size_t)n_seg * sizeof(uint32_t), st));
  nms_prepare_kernel<<<cdiv(P, 256), 256, 0, st>>>(boxes, scores, groups, valid, seg_offsets, n_seg,
                                                   n_total, P, offset_mode, b.keys, b.vals, b.seg_max, status);
  DGOD_LAUNCHED();
  int rc = bitonic_sort(b.keys, b.vals, P, st);
  if (rc) return rc;
  nms_gather_kernel<<<cdiv(P, 256), 256, 0, st>>>(boxes, groups, b.keys, b.vals, b.seg_max, P,
                                                  offset_mode, b.sbox, b.runkey);
  DGOD_LAUNCHED();
  // positions >= n_total are padding (kNoRun) and sort to the tail: only n_total positions matter
  rc = launch_nms_mask(b.sbox, b.runkey, n_total, max_seg_len, thr, b.mask, b.diag_cols, st);
  if (rc) return rc;
  rc = launch_nms_scan(b.mask, b.diag_cols, b.runkey, nullptr, n_total, max_seg_len, b.keepbits, nullptr, nullptr, st);
  if (rc) return rc;
  nms_rekey_kernel<<<cdiv(P, 256), 256, 0, st>>>(b.keys, b.keepbits, P);
  DGOD_LAUNCHED();
  rc = bitonic_sort(b.keys, b.vals, P, st);
  if (rc) return rc;
  nms_output_kernel<<<cdiv(P, 256), 256, 0, st>>>(b.keys, b.vals, seg_offsets, P, out_stride,
                                                  keep_out, keep_count);
  DGOD_LAUNCHED();
  return DGOD_OK;
}
