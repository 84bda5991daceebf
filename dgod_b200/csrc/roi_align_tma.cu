// MultiScaleRoIAlign forward / backward for channels_last features, built on the TMA engine
// (fasterrcnn.py:278,412-416 -> TV ops/poolers.py:147-227 ->
// torchvision::roi_align / _roi_align_backward).
//
// In NHWC a row of an RoI's footprint — pixels x0..x1 of feature row y, all C channels — is ONE
// contiguous span of (x1-x0+1)*C*s bytes.  Both kernels move whole footprint rows with the bulk
// async-copy engine instead of issuing per-tap loads / per-tap reductions from the SM lanes:
//
//   forward   cp.async.bulk global->shared (mbarrier complete_tx) streams the live footprint rows
//             through a multi-stage ring (up to 8 rows in flight per CTA); the pooled [C][49]
//             block leaves as one bulk store.
//   backward  each footprint row is assembled in shared memory and added to the gradient map by
//             ONE cp.reduce.async.bulk (SASS UBLKRED) — the L2 does the read-modify-write.  The
//             previous kernel issued 16 RED per output element and was bound by the SM's RED
//             issue rate (~1.3 cycles per lane).
//
// Arithmetic: bilinear pooling is separable.  With A_y[y][ph] = sum over the sampling rows of bin
// ph of the weight they put on feature row y (and the per-sample column taps likewise),
//       out[c][ph][pw]  = sum_y A_y[y][ph] * ( sum_{sx in pw} h*f[y][xlo][c] + l*f[y][xhi][c] ) / count
//       grad_f[y][x][c] = sum_pw A_x[x][pw] * ( sum_ph A_y[y][ph] * g[c][ph][pw] ) / count
// One thread owns one channel and keeps the 49 pooled values / gradients of that channel in
// registers; the tables are built once per RoI in shared memory and read as warp broadcasts.  A
// sampling axis has at most 2*PH*sr = 28 live rows (columns) however large the RoI is, so rows
// are visited through a compact list; RoIs whose column span exceeds 32 pixels switch from one
// span per row to one 2-pixel slot per sample ("slot mode").  Out-of-range samples (TV's skip
// rule) and the border clamp come from the same axis_tap() the exact kernels use.  The summation
// order differs from the CPU kernel's sample-by-sample order: results agree to fp32 rounding
// (tests: 1e-5 relative), not bitwise.
//
// Roofline: HBM.  Forward reads every touched feature line once (the image's maps stay in the
// 126 MB L2 while its RoIs are in flight) and writes K*C*49*s.  Backward reads K*C*49*s and
// writes every gradient line once: a persistent cooperative grid zero-fills image b+1 while it
// reduces image b's RoIs (per-image completion counters, no grid-wide barrier), so the
// read-modify-write traffic stays in L2 and DRAM sees each line once.  Measured on B200
// (tools/microbench/tma_bulk.cu): bulk reduce 5.8 TB/s into an L2-resident region, 3.0 TB/s into
// a 400 MB one (DRAM read-modify-write); bulk load up to 17 TB/s on L2 hits.
#include "roi_common.cuh"

namespace dgod {

constexpr int kP = 7;                  // pooled size handled by these kernels (PH = PW = 7)
constexpr int kNB = kP * kP;
constexpr int kMaxSamp = 14;           // samples per axis (kP * sampling_ratio, sr <= 2)
constexpr int kMaxLive = 2 * kMaxSamp; // live rows / columns per axis
constexpr int kSpanMax = 32;           // widest column span moved as one piece per row
constexpr int kStagesMax = 8;
constexpr int kFwdRing = 96 * 1024;    // forward: row ring (also the output staging block)
constexpr int kBwdRing = 64 * 1024;    // backward: gradient-block staging, then the row buffers

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
template <typename T> __device__ __forceinline__ void bulk_reduce_add(void* gmem_dst, const void* smem_src, unsigned bytes);
template <> __device__ __forceinline__ void bulk_reduce_add<float>(void* gmem_dst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;\n" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
template <> __device__ __forceinline__ void bulk_reduce_add<__nv_bfloat16>(void* gmem_dst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.noftz.bf16 [%0], [%1], %2;\n" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------- per-RoI tables (shared memory)
struct TapEntry { unsigned off_lo, off_hi; float h, l; };   // byte offsets inside a row buffer + weights

struct alignas(16) RoiSmem {
  alignas(16) float ay[kMaxLive][8];   // dense A_y of the live rows (list order)
  alignas(16) float ax[kSpanMax][8];   // backward, span mode: dense A_x of the span's columns
  alignas(16) TapEntry xs[kMaxSamp];   // column taps of every sample (row-buffer byte offsets)
  alignas(8) unsigned long long full[kStagesMax];
  alignas(8) unsigned long long gbar;  // backward: gradient block landed
  RoiGeom geo;
  short ylo[kMaxSamp], yhi[kMaxSamp], xlo[kMaxSamp], xhi[kMaxSamp];
  float yl[kMaxSamp], yh[kMaxSamp], xl[kMaxSamp], xh[kMaxSamp];   // both zero: sample skipped
  short rows[kMaxLive];                // live feature rows, ascending
  short slot_of[kMaxSamp];             // slot mode: compact slot index of a valid sample (-1 otherwise)
  int n_rows, x_first, x_last, n_slots, slot_mode;
  int next_k;
};

// Lane `lane` < n fills sample `lane` of one axis.
__device__ __forceinline__ void fill_axis_samples(short* lo, short* hi, float* l, float* h, int lane, int n, int sr,
                                                  float start, float bin, int size) {
  if (lane < n) {
    const AxisTap a = axis_tap(sample_coord(start, lane / sr, bin, lane % sr, sr), size);
    lo[lane] = (short)a.lo;
    hi[lane] = (short)a.hi;
    l[lane] = a.valid ? a.l : 0.f;
    h[lane] = a.valid ? a.h : 0.f;
  }
}

// weight the samples of bin p put on coordinate `coord`
__device__ __forceinline__ float axis_weight(const short* lo, const short* hi, const float* l, const float* h, int coord, int p,
                                             int sr) {
  float w = 0.f;
  for (int i = 0; i < sr; ++i) {
    const int q = p * sr + i;
    if (lo[q] == coord) w += h[q];
    if (hi[q] == coord) w += l[q];   // lo == hi at the clamped border: both weights land on the same pixel
  }
  return w;
}

// Warp 0: compact, ascending list of the rows that receive weight; warp 1: column extent, slot map.
// Call with all threads after the sample tables are visible; ends with the lists visible.
template <int SR>
__device__ __forceinline__ void build_lists(RoiSmem& t, int tid) {
  constexpr int NS = kP * SR;
  const int lane = tid & 31;
  if (tid < 32) {
    const int s = lane >> 1, is_hi = lane & 1;
    bool valid = false;
    int r = 0;
    if (s < NS) {
      const bool live = t.yh[s] != 0.f || t.yl[s] != 0.f;
      valid = live && (is_hi ? (t.yl[s] != 0.f && t.yhi[s] != t.ylo[s]) : true);
      r = is_hi ? t.yhi[s] : t.ylo[s];
    }
    const unsigned same = __match_any_sync(0xffffffffu, valid ? r : (0x10000 + lane));
    const bool first = valid && (__ffs(same) - 1 == lane);
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 2 * NS; ++j) {
      const int rj = __shfl_sync(0xffffffffu, r, j);
      const int fj = __shfl_sync(0xffffffffu, (int)first, j);
      rank += (fj && rj < r) ? 1 : 0;
    }
    if (first) t.rows[rank] = (short)r;
    const unsigned m = __ballot_sync(0xffffffffu, first);
    if (lane == 0) t.n_rows = __popc(m);
  } else if (tid < 64) {
    int lo = 0x7fffffff, hi = -1;
    bool valid = false;
    if (lane < NS) {
      valid = t.xh[lane] != 0.f || t.xl[lane] != 0.f;
      if (valid) { lo = t.xlo[lane]; hi = (t.xl[lane] != 0.f) ? t.xhi[lane] : t.xlo[lane]; }
    }
    const unsigned vm = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    if (lane < NS) t.slot_of[lane] = valid ? (short)__popc(vm & ((1u << lane) - 1u)) : (short)-1;
    if (lane == 0) {
      t.x_first = lo;
      t.x_last = hi;
      t.n_slots = __popc(vm);
      t.slot_mode = (hi - lo + 1) > kSpanMax;
    }
  }
  __syncthreads();
}

template <typename T> __device__ __forceinline__ float ld_elem(const unsigned char* p);
template <> __device__ __forceinline__ float ld_elem<float>(const unsigned char* p) { return *reinterpret_cast<const float*>(p); }
template <> __device__ __forceinline__ float ld_elem<__nv_bfloat16>(const unsigned char* p) {
  return __uint_as_float((unsigned)(*reinterpret_cast<const unsigned short*>(p)) << 16);
}

// ================================================================================================
// Forward
// ================================================================================================
template <typename T, int C, int SR>
__global__ void __launch_bounds__(C, (C == 256) ? 2 : 4)
msroi_fwd_tma_kernel(const RoiDev g, const float* __restrict__ rois, int n_rois, T* __restrict__ out) {
  constexpr int NS = kP * SR;
  constexpr int PIX = C * (int)sizeof(T);            // bytes per pixel
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ RoiSmem t;
  const int tid = threadIdx.x;
  const int k = blockIdx.x;

  if (tid == 0) {
    t.geo = roi_geometry(g, rois + (size_t)k * 5);
#pragma unroll
    for (int i = 0; i < kStagesMax; ++i) mbar_init(&t.full[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const RoiGeom r = t.geo;
  const bool usable = r.batch >= 0 && r.batch < g.B;
  if (usable) {
    if (tid < 32) fill_axis_samples(t.ylo, t.yhi, t.yl, t.yh, tid, NS, SR, r.start_h, r.bin_h, r.H);
    else if (tid < 64) fill_axis_samples(t.xlo, t.xhi, t.xl, t.xh, tid - 32, NS, SR, r.start_w, r.bin_w, r.W);
  }
  __syncthreads();

  float acc[kNB];
#pragma unroll
  for (int i = 0; i < kNB; ++i) acc[i] = 0.f;

  if (usable) {
    build_lists<SR>(t, tid);
    const int n_rows = t.n_rows;
    if (n_rows > 0 && t.x_first <= t.x_last) {
      const int x_first = t.x_first, slot_mode = t.slot_mode;
      const int row_px = slot_mode ? 2 * t.n_slots : (t.x_last - x_first + 1);
      const unsigned row_bytes = (unsigned)row_px * PIX;              // multiple of 128
      const int n_stage = min(kStagesMax, kFwdRing / (int)row_bytes);   // >= 3 (row_bytes <= 32 KB)
      // tables: A_y of the live rows, column taps as byte offsets into a row buffer
      for (int e = tid; e < kMaxLive * 8; e += C) {
        const int i = e >> 3, p = e & 7;
        t.ay[i][p] = (i < n_rows && p < kP) ? axis_weight(t.ylo, t.yhi, t.yl, t.yh, t.rows[i], p, SR) : 0.f;
      }
      if (tid < NS) {
        TapEntry e;
        const bool valid = t.xh[tid] != 0.f || t.xl[tid] != 0.f;
        e.h = t.xh[tid];
        e.l = t.xl[tid];
        if (!valid) {
          e.off_lo = e.off_hi = 0u;      // weights are zero; offset 0 is always a loaded pixel
        } else if (slot_mode) {
          e.off_lo = (unsigned)t.slot_of[tid] * 2u * PIX;
          e.off_hi = e.off_lo + ((t.xhi[tid] != t.xlo[tid]) ? PIX : 0u);
        } else {
          e.off_lo = (unsigned)(t.xlo[tid] - x_first) * PIX;
          e.off_hi = (unsigned)(t.xhi[tid] - x_first) * PIX;
        }
        t.xs[tid] = e;
      }
      __syncthreads();
      const T* __restrict__ img = reinterpret_cast<const T*>(g.feat[r.level]) + (size_t)r.batch * r.H * r.W * C;
      auto issue_row = [&](int i) {      // thread 0 only
        const int st = i % n_stage;
        const size_t y = (size_t)t.rows[i];
        unsigned char* dst = ring + (size_t)st * row_bytes;
        if (!slot_mode) {
          mbar_expect_tx(&t.full[st], row_bytes);
          bulk_load(dst, img + (y * r.W + x_first) * C, row_bytes, &t.full[st]);
        } else {
          unsigned total = 0;
          for (int s = 0; s < NS; ++s)
            if (t.slot_of[s] >= 0) total += (t.xhi[s] != t.xlo[s]) ? 2u * PIX : (unsigned)PIX;
          mbar_expect_tx(&t.full[st], total);
          for (int s = 0; s < NS; ++s)
            if (t.slot_of[s] >= 0)
              bulk_load(dst + (size_t)t.slot_of[s] * 2 * PIX, img + (y * r.W + t.xlo[s]) * C,
                        (t.xhi[s] != t.xlo[s]) ? 2u * PIX : (unsigned)PIX, &t.full[st]);
        }
      };
      if (tid == 0)
        for (int i = 0; i < min(n_stage, n_rows); ++i) issue_row(i);
      for (int i = 0; i < n_rows; ++i) {
        const int st = i % n_stage;
        mbar_wait(&t.full[st], (unsigned)(i / n_stage) & 1u);
        const unsigned char* __restrict__ row = ring + (size_t)st * row_bytes + tid * (int)sizeof(T);
        float rx[kP];
#pragma unroll
        for (int p = 0; p < kP; ++p) rx[p] = 0.f;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const uint4 e = *reinterpret_cast<const uint4*>(&t.xs[s]);
          const float v0 = ld_elem<T>(row + e.x), v1 = ld_elem<T>(row + e.y);
          rx[s / SR] = fmaf(__uint_as_float(e.w), v1, fmaf(__uint_as_float(e.z), v0, rx[s / SR]));
        }
        const float4 a0 = *reinterpret_cast<const float4*>(&t.ay[i][0]);
        const float4 a1 = *reinterpret_cast<const float4*>(&t.ay[i][4]);
#pragma unroll
        for (int pw = 0; pw < kP; ++pw) {
          acc[0 * kP + pw] = fmaf(a0.x, rx[pw], acc[0 * kP + pw]);
          acc[1 * kP + pw] = fmaf(a0.y, rx[pw], acc[1 * kP + pw]);
          acc[2 * kP + pw] = fmaf(a0.z, rx[pw], acc[2 * kP + pw]);
          acc[3 * kP + pw] = fmaf(a0.w, rx[pw], acc[3 * kP + pw]);
          acc[4 * kP + pw] = fmaf(a1.x, rx[pw], acc[4 * kP + pw]);
          acc[5 * kP + pw] = fmaf(a1.y, rx[pw], acc[5 * kP + pw]);
          acc[6 * kP + pw] = fmaf(a1.z, rx[pw], acc[6 * kP + pw]);
        }
        __syncthreads();                    // everyone is done with stage st: it may be refilled
        if (tid == 0 && i + n_stage < n_rows) issue_row(i + n_stage);
      }
    }
  }
  __syncthreads();
  // pooled block -> shared [C][49] (lane stride 49 elements: conflict-free) -> one bulk store
  const float inv = 1.f / r.count;          // count = sr*sr: a power of two
  T* so = reinterpret_cast<T*>(ring) + tid * kNB;
#pragma unroll
  for (int i = 0; i < kNB; ++i) so[i] = from_f32<T>(acc[i] * inv);
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    bulk_store(out + (size_t)k * C * kNB, ring, (unsigned)(C * kNB * sizeof(T)));
    bulk_commit();
    bulk_wait_read<0>();
  }
}

// ================================================================================================
// Backward
// ================================================================================================
// One RoI by one CTA (blockDim.x == C).  ring: gradient-block staging [C][49], then row buffers.
template <typename T, int C, int SR>
__device__ __forceinline__ void bwd_one_roi(const RoiDev& g, RoiSmem& t, unsigned char* ring, const T* __restrict__ grad_out,
                                            const float* __restrict__ rois, int k, unsigned gbar_parity) {
  constexpr int NS = kP * SR;
  constexpr int PIX = C * (int)sizeof(T);
  const int tid = threadIdx.x;
  // caller guarantees: no bulk operation of this CTA still reads `ring`, and a __syncthreads() since
  if (tid == 0) {
    mbar_expect_tx(&t.gbar, (unsigned)(C * kNB * sizeof(T)));
    bulk_load(ring, grad_out + (size_t)k * C * kNB, (unsigned)(C * kNB * sizeof(T)), &t.gbar);
    t.geo = roi_geometry(g, rois + (size_t)k * 5);
  }
  __syncthreads();
  const RoiGeom r = t.geo;
  const bool usable = r.batch >= 0 && r.batch < g.B;
  if (usable) {
    if (tid < 32) fill_axis_samples(t.ylo, t.yhi, t.yl, t.yh, tid, NS, SR, r.start_h, r.bin_h, r.H);
    else if (tid < 64) fill_axis_samples(t.xlo, t.xhi, t.xl, t.xh, tid - 32, NS, SR, r.start_w, r.bin_w, r.W);
  }
  __syncthreads();
  int n_rows = 0;
  if (usable) {
    build_lists<SR>(t, tid);
    n_rows = (t.x_first <= t.x_last) ? t.n_rows : 0;
  }
  const int x_first = t.x_first, slot_mode = t.slot_mode;
  const int row_px = n_rows ? (slot_mode ? 2 * t.n_slots : (t.x_last - x_first + 1)) : 1;
  const unsigned row_bytes = (unsigned)row_px * PIX;
  const int n_buf = min(4, kBwdRing / (int)row_bytes);        // 2..4 row buffers
  if (n_rows) {
    const float inv = 1.f / (float)(SR * SR);                 // the CPU backward divides by the raw grid product
    for (int e = tid; e < kMaxLive * 8; e += C) {
      const int i = e >> 3, p = e & 7;
      t.ay[i][p] = (i < n_rows && p < kP) ? axis_weight(t.ylo, t.yhi, t.yl, t.yh, t.rows[i], p, SR) * inv : 0.f;
    }
    if (!slot_mode) {
      for (int e = tid; e < kSpanMax * 8; e += C) {
        const int i = e >> 3, p = e & 7;
        t.ax[i][p] = (i < row_px && p < kP) ? axis_weight(t.xlo, t.xhi, t.xl, t.xh, x_first + i, p, SR) : 0.f;
      }
    }
  }
  // the gradient block has landed: 49 values of this thread's channel -> registers
  mbar_wait(&t.gbar, gbar_parity);
  float gr[kNB];
  {
    const T* sg = reinterpret_cast<const T*>(ring) + tid * kNB;
#pragma unroll
    for (int i = 0; i < kNB; ++i) gr[i] = to_f32<T>(sg[i]);
  }
  __syncthreads();                          // staging consumed, tables visible
  if (!n_rows) return;
  T* __restrict__ img = reinterpret_cast<T*>(g.gfeat[r.level]) + (size_t)r.batch * r.H * r.W * C;
  for (int i = 0; i < n_rows; ++i) {
    const float4 a0 = *reinterpret_cast<const float4*>(&t.ay[i][0]);
    const float4 a1 = *reinterpret_cast<const float4*>(&t.ay[i][4]);
    float tq[kP];
#pragma unroll
    for (int pw = 0; pw < kP; ++pw) {
      float s = a0.x * gr[0 * kP + pw];
      s = fmaf(a0.y, gr[1 * kP + pw], s);
      s = fmaf(a0.z, gr[2 * kP + pw], s);
      s = fmaf(a0.w, gr[3 * kP + pw], s);
      s = fmaf(a1.x, gr[4 * kP + pw], s);
      s = fmaf(a1.y, gr[5 * kP + pw], s);
      s = fmaf(a1.z, gr[6 * kP + pw], s);
      tq[pw] = s;
    }
    unsigned char* buf = ring + (size_t)(i % n_buf) * row_bytes;
    T* __restrict__ row = reinterpret_cast<T*>(buf) + tid;
    if (!slot_mode) {
#pragma unroll 4
      for (int x = 0; x < row_px; ++x) {
        const float4 w0 = *reinterpret_cast<const float4*>(&t.ax[x][0]);
        const float4 w1 = *reinterpret_cast<const float4*>(&t.ax[x][4]);
        float v = w0.x * tq[0];
        v = fmaf(w0.y, tq[1], v); v = fmaf(w0.z, tq[2], v); v = fmaf(w0.w, tq[3], v);
        v = fmaf(w1.x, tq[4], v); v = fmaf(w1.y, tq[5], v); v = fmaf(w1.z, tq[6], v);
        row[x * C] = from_f32<T>(v);
      }
    } else {
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const int slot = t.slot_of[s];
        if (slot >= 0) {                    // warp-uniform
          const bool two = t.xhi[s] != t.xlo[s];
          row[(slot * 2) * C] = from_f32<T>((two ? t.xh[s] : t.xh[s] + t.xl[s]) * tq[s / SR]);
          if (two) row[(slot * 2 + 1) * C] = from_f32<T>(t.xl[s] * tq[s / SR]);
        }
      }
    }
    fence_proxy_async_smem();               // generic-proxy writes -> visible to the bulk engine
    if (tid == 0) {                         // the buffer the NEXT row will write must have been read
      if (n_buf == 2) bulk_wait_read<0>();
      else if (n_buf == 3) bulk_wait_read<1>();
      else bulk_wait_read<2>();
    }
    __syncthreads();
    if (tid == 0) {
      const size_t y = (size_t)t.rows[i];
      if (!slot_mode) {
        bulk_reduce_add<T>(img + (y * r.W + x_first) * C, buf, row_bytes);
      } else {
        for (int s = 0; s < NS; ++s)
          if (t.slot_of[s] >= 0)
            bulk_reduce_add<T>(img + (y * r.W + t.xlo[s]) * C, buf + (size_t)t.slot_of[s] * 2 * PIX,
                               (t.xhi[s] != t.xlo[s]) ? 2u * PIX : (unsigned)PIX);
      }
      bulk_commit();
    }
  }
}

// One CTA per RoI; gradients must have been zero-filled.
template <typename T, int C, int SR>
__global__ void __launch_bounds__(C, (C == 256) ? 3 : 4)
msroi_bwd_tma_kernel(const RoiDev g, const T* __restrict__ grad_out, const float* __restrict__ rois, int n_rois) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ RoiSmem t;
  if (threadIdx.x == 0) {
    mbar_init(&t.gbar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  bwd_one_roi<T, C, SR>(g, t, ring, grad_out, rois, blockIdx.x, 0u);
  if (threadIdx.x == 0) bulk_wait_read<0>();
}

// Persistent cooperative variant.  Every CTA zero-fills its share of image b+1's gradient maps
// BEFORE it starts on image b's RoIs and publishes that in zero_done[b+1]; an RoI of image b is
// only started once zero_done[b] == gridDim.x.  All CTAs are co-resident (cooperative launch), so
// the wait cannot deadlock, and in steady state nobody waits: the zero-filled lines of an image
// (52.9 MB fp32 at 608x1024) are still dirty in L2 when the reductions arrive.
template <typename T, int C, int SR>
__global__ void __launch_bounds__(C, (C == 256) ? 3 : 4)
msroi_bwd_tma_persistent_kernel(const RoiDev g, const T* __restrict__ grad_out, const float* __restrict__ rois,
                                const int32_t* __restrict__ roi_img_offsets, int* __restrict__ counters) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ RoiSmem t;
  int* work = counters;                 // [B] next RoI of image b
  int* zero_done = counters + g.B;      // [B] CTAs that finished zero-filling image b
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&t.gbar, 1);
    mbar_fence_init();
  }
  auto zero_image = [&](int b) {
    for (int l = 0; l < g.n_levels; ++l) {
      const size_t n16 = (size_t)g.H[l] * g.W[l] * C * sizeof(T) / 16;
      uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<T*>(g.gfeat[l]) + (size_t)b * g.H[l] * g.W[l] * C);
      for (size_t i = (size_t)blockIdx.x * C + tid; i < n16; i += (size_t)gridDim.x * C) p[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("fence.proxy.async.global;\n" ::: "memory");   // generic-proxy zeros before the bulk engine's RMW
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicAdd(&zero_done[b], 1);
  };
  zero_image(0);
  unsigned n_done = 0;                  // RoIs processed by this CTA (parity of the staging barrier)
  for (int b = 0; b < g.B; ++b) {
    if (b + 1 < g.B) zero_image(b + 1);
    if (tid == 0) {
      while (atomicAdd(&zero_done[b], 0) < (int)gridDim.x) __nanosleep(200);
      __threadfence();
    }
    const int k0 = roi_img_offsets[b], k1 = roi_img_offsets[b + 1];
    while (true) {
      if (tid == 0) {
        bulk_wait_read<0>();            // row buffers of the previous RoI have been read
        t.next_k = k0 + atomicAdd(&work[b], 1);
      }
      __syncthreads();
      const int k = t.next_k;
      if (k >= k1) break;
      bwd_one_roi<T, C, SR>(g, t, ring, grad_out, rois, k, n_done & 1u);
      ++n_done;
    }
    __syncthreads();
  }
  if (tid == 0) bulk_wait_read<0>();
}

// ------------------------------------------------------------------------------------------------
static bool tma_shape_ok(const dgod_roi_config* cfg, const RoiDev& g) {
  if (!g.channels_last || g.PH != kP || g.PW != kP) return false;
  if (g.sr < 1 || g.sr > 2) return false;
  if (!(g.C == 256 || ((g.C == 128 || g.C == 64) && g.sr == 2))) return false;   // instantiated shapes
  for (int l = 0; l < g.n_levels; ++l)
    if (g.H[l] > 32000 || g.W[l] > 32000) return false;
  (void)cfg;
  return true;
}

template <typename T, int C, int SR>
static int launch_fwd_tma(const RoiDev& g, const float* rois, int n_rois, void* out, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    DGOD_CUDA(cudaFuncSetAttribute(msroi_fwd_tma_kernel<T, C, SR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdRing));
    attr = true;
  }
  msroi_fwd_tma_kernel<T, C, SR><<<n_rois, C, kFwdRing, st>>>(g, rois, n_rois, (T*)out);
  DGOD_LAUNCHED();
  return DGOD_OK;
}

template <typename T>
static int dispatch_fwd_tma(const RoiDev& g, const float* rois, int n_rois, void* out, cudaStream_t st) {
  if (g.C == 256) return g.sr == 2 ? launch_fwd_tma<T, 256, 2>(g, rois, n_rois, out, st) : launch_fwd_tma<T, 256, 1>(g, rois, n_rois, out, st);
  if (g.C == 128) return launch_fwd_tma<T, 128, 2>(g, rois, n_rois, out, st);
  return launch_fwd_tma<T, 64, 2>(g, rois, n_rois, out, st);
}

int msroi_fwd_tma(const dgod_roi_config* cfg, const RoiDev& g, const float* rois, int n_rois, void* out,
                  cudaStream_t st, int* handled) {
  *handled = 0;
  if (!tma_shape_ok(cfg, g) || ((uintptr_t)out & 15)) return DGOD_OK;
  for (int l = 0; l < g.n_levels; ++l)
    if ((uintptr_t)g.feat[l] & 15) return DGOD_OK;
  *handled = 1;
  return cfg->dtype == DGOD_F32 ? dispatch_fwd_tma<float>(g, rois, n_rois, out, st)
                                : dispatch_fwd_tma<__nv_bfloat16>(g, rois, n_rois, out, st);
}

template <typename T, int C, int SR>
static int launch_bwd_tma(const RoiDev& g, const void* grad_out, const float* rois, int n_rois,
                          const int32_t* roi_img_offsets, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  static bool attr = false;
  static int coop = 0, n_sm = 0, per_sm = 0;
  if (!attr) {
    int dev = 0;
    DGOD_CUDA(cudaGetDevice(&dev));
    DGOD_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    DGOD_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    DGOD_CUDA(cudaFuncSetAttribute(msroi_bwd_tma_kernel<T, C, SR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdRing));
    DGOD_CUDA(cudaFuncSetAttribute(msroi_bwd_tma_persistent_kernel<T, C, SR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdRing));
    DGOD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, msroi_bwd_tma_persistent_kernel<T, C, SR>, C, kBwdRing));
    attr = true;
  }
  const size_t need_ws = 2 * (size_t)g.B * sizeof(int);
  if (coop && per_sm >= 1 && roi_img_offsets && workspace && workspace_bytes >= need_ws) {
    int* counters = (int*)workspace;
    DGOD_CUDA(cudaMemsetAsync(counters, 0, need_ws, st));
    const T* go = (const T*)grad_out;
    RoiDev gg = g;
    void* args[] = {(void*)&gg, (void*)&go, (void*)&rois, (void*)&roi_img_offsets, (void*)&counters};
    DGOD_CUDA(cudaLaunchCooperativeKernel((const void*)msroi_bwd_tma_persistent_kernel<T, C, SR>, dim3(per_sm * n_sm), dim3(C), args,
                                          kBwdRing, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return DGOD_OK;
  }
  for (int l = 0; l < g.n_levels; ++l)
    DGOD_CUDA(cudaMemsetAsync(g.gfeat[l], 0, (size_t)g.B * g.C * g.H[l] * g.W[l] * sizeof(T), st));
  if (n_rois == 0) return DGOD_OK;
  msroi_bwd_tma_kernel<T, C, SR><<<n_rois, C, kBwdRing, st>>>(g, (const T*)grad_out, rois, n_rois);
  DGOD_LAUNCHED();
  return DGOD_OK;
}

template <typename T>
static int dispatch_bwd_tma(const RoiDev& g, const void* grad_out, const float* rois, int n_rois, const int32_t* offs,
                            void* ws, size_t wsb, cudaStream_t st) {
  if (g.C == 256)
    return g.sr == 2 ? launch_bwd_tma<T, 256, 2>(g, grad_out, rois, n_rois, offs, ws, wsb, st)
                     : launch_bwd_tma<T, 256, 1>(g, grad_out, rois, n_rois, offs, ws, wsb, st);
  if (g.C == 128) return launch_bwd_tma<T, 128, 2>(g, grad_out, rois, n_rois, offs, ws, wsb, st);
  return launch_bwd_tma<T, 64, 2>(g, grad_out, rois, n_rois, offs, ws, wsb, st);
}

int msroi_bwd_tma(const dgod_roi_config* cfg, const RoiDev& g, const void* grad_out, const float* rois, int n_rois,
                  const int32_t* roi_img_offsets, void* workspace, size_t workspace_bytes, cudaStream_t st,
                  int* handled) {
  *handled = 0;
  if (!tma_shape_ok(cfg, g) || ((uintptr_t)grad_out & 15)) return DGOD_OK;
  for (int l = 0; l < g.n_levels; ++l)
    if ((uintptr_t)g.gfeat[l] & 15) return DGOD_OK;
  *handled = 1;
  return cfg->dtype == DGOD_F32
             ? dispatch_bwd_tma<float>(g, grad_out, rois, n_rois, roi_img_offsets, workspace, workspace_bytes, st)
             : dispatch_bwd_tma<__nv_bfloat16>(g, grad_out, rois, n_rois, roi_img_offsets, workspace, workspace_bytes, st);
}

}  // namespace dgod
