// MultiScaleRoIAlign forward / backward for channels_last features, built on the TMA engine
// (fasterrcnn.py:278,412-416 -> TV ops/poolers.py:147-227 ->
// torchvision::roi_align / _roi_align_backward).
//
// In NHWC a row of an RoI's footprint — pixels x0..x1 of feature row y, all C channels — is ONE
// contiguous span of (x1-x0+1)*C*s bytes.  Both kernels move whole footprint rows with the bulk
// async-copy engine instead of issuing per-tap loads / per-tap reductions from the SM lanes.
//
// Three kernels:
//   msroi_plan_kernel   one warp per RoI: level mapping, sampling taps (the same axis_tap() the
//                       exact kernels use, i.e. TV's out-of-range skip rule and border clamp), the
//                       compact list of live feature rows, the separable weight tables.  The result
//                       is a 2.4 KB "plan" record per RoI in the caller's workspace, so the
//                       streaming kernels below do no per-RoI arithmetic at all.
//   msroi_fwd_tma       one CTA per RoI, warp-specialised.  Thread 0 bulk-loads the plan; the lanes of a
//                       producer warp stream the live footprint rows through a ring of up to 8 stages
//                       (cp.async.bulk global->shared, mbarrier complete_tx) — one lane per stage, or
//                       one lane per sample for the 1-2 pixel pieces of a wide RoI, because the bulk
//                       copies of ONE thread complete one at a time (~750 cycles each, measured).  C/2
//                       consumer threads own two adjacent channels each (packed fp32 FFMA2), keep the
//                       49 pooled values of both in registers and hand stages back through per-stage
//                       mbarriers — warps never wait for each other inside an RoI.  The pooled
//                       [C][49] block is staged in shared memory and leaves as ONE bulk store.
//   msroi_bwd_tma       persistent, cooperative launch.  A driver lane prefetches each RoI's plan and
//                       [C][49] gradient block; consumer warps assemble every live footprint row in
//                       shared memory and hand it over through an mbarrier; the driver adds it to the
//                       gradient map with ONE cp.reduce.async.bulk (SASS UBLKRED: the L2 does the
//                       read-modify-write) and publishes which row buffers have been read.  (The
//                       per-tap kernel issued 16 RED per output element and was bound by the SM's RED
//                       issue rate, ~1.3 cycles per lane.)  A third warp role zero-fills this CTA's
//                       share of the gradient maps image by image, a bounded distance ahead of the
//                       reductions; an RoI of image b waits for zero_done[b] == gridDim.x, so the maps
//                       need no memset and no grid-wide barrier.
//
// Arithmetic: bilinear pooling is separable.  With A_y[y][ph] = sum over the sampling rows of bin
// ph of the weight they put on feature row y (and the column taps likewise),
//       out[c][ph][pw]  = sum_y A_y[y][ph] * ( sum_{sx in pw} h*f[y][xlo][c] + l*f[y][xhi][c] ) / count
//       grad_f[y][x][c] = sum_pw A_x[x][pw] * ( sum_ph A_y[y][ph] * g[c][ph][pw] ) / count
// A sampling axis has at most 2*PH*sr = 28 live rows (columns) however large the RoI is; RoIs whose
// column span exceeds 28 pixels switch from one span per row to one 2-pixel slot per sample ("slot
// mode").  The summation order differs from the CPU kernel's sample-by-sample order: results agree
// to fp32 rounding (tests: 1e-5 relative), not bitwise.
//
// Roofline: HBM (SURVEY.md §8d byte model).  Measured on B200 (tools/microbench/): bulk reduce
// 5.8 TB/s into an L2-resident region, 3.0 TB/s into a 400 MB one (DRAM read-modify-write), engine
// ceiling 23 B/clk/SM; bulk load up to 21 TB/s on L2 hits with >= 4 issuing lanes per SM (6.3 TB/s
// with one); bulk store 7.3 TB/s.  profiles/README.md has the per-component timings of both kernels.
#include "roi_common.cuh"
#include "bulk.cuh"

namespace dgod {

constexpr int kP = 7;                  // pooled size handled by these kernels (PH = PW = 7)
constexpr int kNB = kP * kP;
constexpr int kMaxSamp = 14;           // samples per axis (kP * sampling_ratio, sr <= 2)
constexpr int kMaxLive = 2 * kMaxSamp; // live rows / columns per axis
constexpr int kSpanMax = 28;           // widest column span moved as one piece per row
constexpr int kPlanBufBwd = 2;
constexpr int kCounterBytes = 4096;    // head of the workspace: zero_done[batch] (backward)
constexpr int kSmemBudget = 112 * 1024;   // static + dynamic per CTA for 2 CTAs per SM

// ---------------------------------------------------------------- the per-RoI plan
struct alignas(16) ColTap { unsigned off_lo, off_hi; float h, l; };   // row-buffer byte offsets of the two taps, their weights

struct alignas(128) RoiPlan {
  // header (64 B)
  int n_rows;            // live feature rows (0: nothing to do — RoI unusable or entirely out of range)
  int row_px;            // pixels per row buffer: the span, or 2 * n_slots in slot mode
  int slot_mode, n_slots;
  int level, batch;
  int x_first;           // span mode: first column of the span
  float inv_count;       // 1 / (sr*sr)
  int pad[8];
  // lists (128 B)
  short rows[kMaxLive];                 // live feature rows, ascending
  short slot_x[kMaxSamp];               // slot mode: first pixel of sample s
  signed char slot_of[kMaxSamp + 2];    // slot mode: compact slot index of a valid sample, -1 otherwise
  unsigned char slot_two[kMaxSamp + 2]; // slot mode: the slot holds two pixels (xhi != xlo)
  unsigned char pad2[12];
  // tables
  ColTap xs[kMaxSamp + 2];              // 256 B   column taps of every sample
  float ay[kMaxLive][8];                // 896 B   A_y of the live rows, list order, unscaled
  // backward only
  float ax[kSpanMax + 4][8];            // 1024 B  span mode, dense A_x of the span's columns (+ zero rows)
};
static_assert(sizeof(RoiPlan) % 128 == 0, "plans are moved with bulk copies");
constexpr int kPlanFwdBytes = (int)offsetof(RoiPlan, ax);       // the forward kernel loads this prefix
constexpr int kPlanBwdBytes = (int)sizeof(RoiPlan);
static_assert(kPlanFwdBytes % 16 == 0, "bulk copy granularity");

// ---------------------------------------------------------------- plan kernel (one warp per RoI)
struct PlanScratch {
  RoiGeom geo;
  short ylo[kMaxSamp], yhi[kMaxSamp], xlo[kMaxSamp], xhi[kMaxSamp];
  float yl[kMaxSamp], yh[kMaxSamp], xl[kMaxSamp], xh[kMaxSamp];   // both zero: sample skipped
  short rows[kMaxLive];
  short slot_of[kMaxSamp];
  int n_rows, x_first, x_last, n_slots;
};

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

constexpr int kPlanWarps = 4;

__global__ void __launch_bounds__(kPlanWarps * 32)
msroi_plan_kernel(const RoiDev g, const float* __restrict__ rois, int n_rois, RoiPlan* __restrict__ plans, int pix_bytes,
                  int want_ax) {
  __shared__ PlanScratch scratch[kPlanWarps];
  pdl_trigger();                           // the streaming kernel may be scheduled behind this grid's CTAs (it waits for the plans)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int k = blockIdx.x * kPlanWarps + w;
  if (k >= n_rois) return;
  PlanScratch& t = scratch[w];
  RoiPlan& P = plans[k];
  const int sr = g.sr, ns = kP * sr;
  if (lane == 0) t.geo = roi_geometry(g, rois + (size_t)k * 5);
  __syncwarp();
  const RoiGeom r = t.geo;
  const bool usable = r.batch >= 0 && r.batch < g.B;
  int n_rows = 0;
  if (usable) {
    fill_axis_samples(t.ylo, t.yhi, t.yl, t.yh, lane, ns, sr, r.start_h, r.bin_h, r.H);
    fill_axis_samples(t.xlo, t.xhi, t.xl, t.xh, lane, ns, sr, r.start_w, r.bin_w, r.W);
    __syncwarp();
    {  // compact, ascending list of the feature rows that receive weight
      const int s = lane >> 1, is_hi = lane & 1;
      bool valid = false;
      int row = 0;
      if (s < ns) {
        const bool live = t.yh[s] != 0.f || t.yl[s] != 0.f;
        valid = live && (is_hi ? (t.yl[s] != 0.f && t.yhi[s] != t.ylo[s]) : true);
        row = is_hi ? t.yhi[s] : t.ylo[s];
      }
      const unsigned same = __match_any_sync(0xffffffffu, valid ? row : (0x10000 + lane));
      const bool first = valid && (__ffs(same) - 1 == lane);
      int rank = 0;
      for (int j = 0; j < 2 * ns; ++j) {
        const int rj = __shfl_sync(0xffffffffu, row, j);
        const int fj = __shfl_sync(0xffffffffu, (int)first, j);
        rank += (fj && rj < row) ? 1 : 0;
      }
      if (first) t.rows[rank] = (short)row;
      const unsigned m = __ballot_sync(0xffffffffu, first);
      if (lane == 0) t.n_rows = __popc(m);
    }
    {  // column extent of the weighted taps, compact slot index of every valid sample
      int lo = 0x7fffffff, hi = -1;
      bool valid = false;
      if (lane < ns) {
        valid = t.xh[lane] != 0.f || t.xl[lane] != 0.f;
        if (valid) { lo = t.xlo[lane]; hi = (t.xl[lane] != 0.f) ? t.xhi[lane] : t.xlo[lane]; }
      }
      const unsigned vm = __ballot_sync(0xffffffffu, valid);
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
      }
      if (lane < ns) t.slot_of[lane] = valid ? (short)__popc(vm & ((1u << lane) - 1u)) : (short)-1;
      if (lane == 0) { t.x_first = lo; t.x_last = hi; t.n_slots = __popc(vm); }
    }
    __syncwarp();
    n_rows = (t.x_first <= t.x_last) ? t.n_rows : 0;
  }
  const int x_first = n_rows ? t.x_first : 0;
  const int span = n_rows ? (t.x_last - x_first + 1) : 0;
  const int slot_mode = span > kSpanMax;
  const int n_slots = n_rows ? t.n_slots : 0;
  const int row_px = slot_mode ? 2 * n_slots : span;
  if (lane == 0) {
    P.n_rows = n_rows; P.row_px = row_px; P.slot_mode = slot_mode; P.n_slots = n_slots;
    P.level = r.level; P.batch = usable ? r.batch : 0; P.x_first = x_first;
    P.inv_count = 1.f / r.count;           // count = sr*sr: a power of two
  }
  if (!n_rows) return;
  if (lane < kMaxLive) P.rows[lane] = lane < n_rows ? t.rows[lane] : (short)0;
  if (lane < kMaxSamp) {
    const bool in = lane < ns;
    const bool valid = in && (t.xh[lane] != 0.f || t.xl[lane] != 0.f);
    const bool two = valid && t.xhi[lane] != t.xlo[lane];
    P.slot_x[lane] = valid ? t.xlo[lane] : (short)0;
    P.slot_of[lane] = (signed char)(valid ? t.slot_of[lane] : -1);
    P.slot_two[lane] = (unsigned char)two;
    ColTap e;
    e.h = in ? t.xh[lane] : 0.f;
    e.l = in ? t.xl[lane] : 0.f;
    if (!valid) {
      e.off_lo = e.off_hi = 0u;          // weights are zero; offset 0 is always a loaded pixel
    } else if (slot_mode) {
      e.off_lo = (unsigned)t.slot_of[lane] * 2u * pix_bytes;
      e.off_hi = e.off_lo + (two ? (unsigned)pix_bytes : 0u);
    } else {
      e.off_lo = (unsigned)(t.xlo[lane] - x_first) * pix_bytes;
      e.off_hi = (unsigned)(t.xhi[lane] - x_first) * pix_bytes;
    }
    P.xs[lane] = e;
  }
  for (int e = lane; e < kMaxLive * 8; e += 32) {
    const int i = e >> 3, p = e & 7;
    const float a = (i < n_rows && p < kP) ? axis_weight(t.ylo, t.yhi, t.yl, t.yh, t.rows[i], p, sr) : 0.f;
    P.ay[i][p] = a;
  }
  if (!want_ax) return;                // the forward kernel needs nothing below
  if (!slot_mode) {
    for (int e = lane; e < (kSpanMax + 4) * 8; e += 32) {
      const int i = e >> 3, p = e & 7;
      P.ax[i][p] = (i < span && p < kP) ? axis_weight(t.xlo, t.xhi, t.xl, t.xh, x_first + i, p, sr) : 0.f;
    }
  }
}

template <typename T> __device__ __forceinline__ float2 ld_pair(const unsigned char* p);
template <> __device__ __forceinline__ float2 ld_pair<float>(const unsigned char* p) { return *reinterpret_cast<const float2*>(p); }
template <> __device__ __forceinline__ float2 ld_pair<__nv_bfloat16>(const unsigned char* p) {
  const unsigned u = *reinterpret_cast<const unsigned*>(p);
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
template <typename T> __device__ __forceinline__ void st_pair(unsigned char* p, float2 v);
template <> __device__ __forceinline__ void st_pair<float>(unsigned char* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
template <> __device__ __forceinline__ void st_pair<__nv_bfloat16>(unsigned char* p, float2 v) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y);
}

// Shared-memory accesses of the hot loops go through 32-bit shared-window addresses computed once per
// row.  (With generic pointers into the dynamic shared array the compiler rematerialises the window
// base — S2UR SR_CgaCtaId + ULEA — in front of every predicated store; ncu showed that pair as the
// top stall of the backward row loop.)  The asm statements are volatile without a memory clobber:
// they keep their order among themselves and relative to the mbarrier / fence statements.
__device__ __forceinline__ float4 lds_f4(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
template <typename T> __device__ __forceinline__ float2 lds_pair(unsigned a);
template <> __device__ __forceinline__ float2 lds_pair<float>(unsigned a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
template <> __device__ __forceinline__ float2 lds_pair<__nv_bfloat16>(unsigned a) {
  unsigned u;
  asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(u) : "r"(a));
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
template <typename T> __device__ __forceinline__ void sts_pair(unsigned a, float2 v);
template <> __device__ __forceinline__ void sts_pair<float>(unsigned a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};\n" ::"r"(a), "f"(v.x), "f"(v.y));
}
template <> __device__ __forceinline__ void sts_pair<__nv_bfloat16>(unsigned a, float2 v) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
  asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(a), "r"(*reinterpret_cast<const unsigned*>(&h)));
}

// ================================================================================================
// Forward
// ================================================================================================
// One CTA per RoI (two resident per SM), warp-specialised: a producer lane streams the RoI's live
// rows through a ring of up to 8 stages; C/2 consumer threads own two adjacent channels each.
// (A persistent variant that packed the rows of consecutive RoIs into one ring measured slower on
// B200: with the pooled block staged separately the ring shrank to ~3 rows in flight per CTA, and the
// loads are latency-bound.)  Here the staging block aliases the ring once the rows are consumed.
constexpr int kStagesMax = 8;

struct alignas(16) FwdSync {
  alignas(8) unsigned long long plan_full;
  alignas(8) unsigned long long full[kStagesMax], empty[kStagesMax];
};

template <typename T, int C> struct FwdCfg {
  static constexpr int kStaging = C * kNB * (int)sizeof(T);                       // pooled block of one RoI
  static constexpr int kRing = ((kSmemBudget - kPlanFwdBytes - (int)sizeof(FwdSync) - 256) / 1024) * 1024;
  static constexpr int kSmem = kRing + kPlanFwdBytes;
  static_assert(kRing >= kStaging && kRing >= 3 * kSpanMax * C * (int)sizeof(T), "ring: staging alias and >= 3 widest rows");
};

template <typename T, int C, int SR>
__global__ void __launch_bounds__(C / 2 + 32, 2)
msroi_fwd_tma_kernel(const RoiDev g, const RoiPlan* __restrict__ plans, int n_rois, T* __restrict__ out) {
  constexpr int NS = kP * SR;
  constexpr int NC = C / 2;                          // consumer threads
  constexpr int PIX = C * (int)sizeof(T);            // bytes per pixel
  constexpr int RING = FwdCfg<T, C>::kRing;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* ring = smem;                        // row stages, then the pooled block [C][49]
  const RoiPlan& P = *reinterpret_cast<const RoiPlan*>(smem + RING);
  __shared__ FwdSync f;
  const int tid = threadIdx.x;
  const int k = blockIdx.x;

  if (tid == 0) {
    mbar_init(&f.plan_full, 1);
#pragma unroll
    for (int i = 0; i < kStagesMax; ++i) {
      mbar_init(&f.full[i], 1);
      mbar_init(&f.empty[i], NC / 32);
    }
    mbar_fence_init();
    pdl_wait();                            // launched behind the plan kernel (programmatic dependent launch)
    mbar_expect_tx(&f.plan_full, kPlanFwdBytes);
    bulk_load_hint(smem + RING, plans + k, kPlanFwdBytes, &f.plan_full, policy_evict_first());
  }
  __syncthreads();
  if (tid >= NC) {
    // Producer warp: the first row of every ring stage is requested from the plan's header in GLOBAL memory (a few
    // independent loads) while the plan's own bulk copy into shared memory is still in flight — the consumers' first
    // wait shrinks by the difference between a bulk-copy round trip + barrier wake-up and a plain load.  (Thread 0
    // executed pdl_wait() before the barrier above: the plan kernel's writes are visible to the whole grid.)
    const RoiPlan* __restrict__ gp = plans + k;
    const int4 h0 = __ldcg(reinterpret_cast<const int4*>(gp));          // n_rows, row_px, slot_mode, n_slots
    const int lane = tid - NC;
    if (h0.x > 0) {
      const int4 h1 = __ldcg(reinterpret_cast<const int4*>(gp) + 1);    // level, batch, x_first, inv_count
      const unsigned row_bytes = (unsigned)h0.y * PIX;
      const int n_stage = min(kStagesMax, RING / (int)row_bytes);
      const int W = g.W[h1.x];
      const T* __restrict__ img = reinterpret_cast<const T*>(g.feat[h1.x]) + (size_t)h1.y * g.H[h1.x] * W * C;
      const unsigned long long keep = policy_evict_last();
      if (!h0.z) {
        if (lane < n_stage && lane < h0.x) {
          const int row = (int)__ldcg(&gp->rows[lane]);
          mbar_expect_tx(&f.full[lane], row_bytes);
          bulk_load_hint(ring + (size_t)lane * row_bytes, img + ((size_t)row * W + h1.z) * C, row_bytes, &f.full[lane], keep);
        }
      } else {
        // slot mode: lane sx issues its 1-2 pixel piece of the first n_stage rows
        int so = -1, two = 0, sx0 = 0;
        if (lane < NS) {
          so = (int)__ldcg(&gp->slot_of[lane]);
          two = (int)__ldcg(&gp->slot_two[lane]);
          sx0 = (int)__ldcg(&gp->slot_x[lane]);
        }
        const int4 r03 = __ldcg(reinterpret_cast<const int4*>(gp->rows));          // rows 0..7 (kStagesMax)
        const unsigned vm = __ballot_sync(0xffffffffu, so >= 0), tm = __ballot_sync(0xffffffffu, so >= 0 && two);
        const unsigned slot_total = (unsigned)(__popc(vm) + __popc(tm)) * PIX;
        const int first = __ffs(vm) - 1;
        const int n_early = min(n_stage, h0.x);
        if (so >= 0) {
          const unsigned bytes = two ? 2u * PIX : (unsigned)PIX;
          const int rw[4] = {r03.x, r03.y, r03.z, r03.w};
          for (int i = 0; i < n_early; ++i) {
            const int row = (int)(short)((unsigned)rw[i >> 1] >> (16 * (i & 1)));
            if (lane == first) mbar_expect_tx(&f.full[i], slot_total);
            bulk_load_hint(ring + (size_t)i * row_bytes + (size_t)so * 2 * PIX, img + ((size_t)row * W + sx0) * C, bytes, &f.full[i], keep);
          }
        }
      }
    }
  }
  mbar_wait(&f.plan_full, 0u);
  const int n_rows = P.n_rows;
  const float inv = P.inv_count;
  float2 acc[kNB];
#pragma unroll
  for (int i = 0; i < kNB; ++i) acc[i] = make_float2(0.f, 0.f);

  if (n_rows) {
    const unsigned row_bytes = (unsigned)P.row_px * PIX;                // multiple of 128, <= 28 KB
    const int n_stage = min(kStagesMax, RING / (int)row_bytes);        // >= 3
    if (tid >= NC) {
      // ---------------- producer warp.  Bulk copies issued by one thread complete one at a time, ~750 cycles each
      // whatever their size (tools/microbench/bulk_load_issue.cu), while copies of different lanes overlap.  Span
      // mode: lane p owns ring stage p and streams rows p, p + n_stage, ...  Slot mode: lane sx owns sample sx and
      // issues its 1-2 pixel piece of every row.  Either way every lane waits on the phases of a barrier in order.
      const int lane = tid - NC;
      const int slot_mode = P.slot_mode, W = g.W[P.level];
      const T* __restrict__ img = reinterpret_cast<const T*>(g.feat[P.level]) + (size_t)P.batch * g.H[P.level] * W * C;
      const unsigned long long keep = policy_evict_last();      // the image's maps are re-read by its other RoIs
      if (!slot_mode) {
        if (lane < n_stage) {
          const T* __restrict__ span0 = img + (size_t)P.x_first * C;
          const int st = lane;
          for (int i = st + n_stage; i < n_rows; i += n_stage) {        // row st itself was requested above
            mbar_wait(&f.empty[st], (unsigned)(i / n_stage - 1) & 1u);
            mbar_expect_tx(&f.full[st], row_bytes);
            bulk_load_hint(ring + (size_t)st * row_bytes, span0 + (size_t)P.rows[i] * W * C, row_bytes, &f.full[st], keep);
          }
        }
      } else if (lane < NS && P.slot_of[lane] >= 0) {
        unsigned slot_total = 0;
        int first = -1;
        for (int sx = 0; sx < NS; ++sx)
          if (P.slot_of[sx] >= 0) {
            slot_total += P.slot_two[sx] ? 2u * PIX : (unsigned)PIX;
            if (first < 0) first = sx;
          }
        const unsigned bytes = P.slot_two[lane] ? 2u * PIX : (unsigned)PIX;
        const size_t dst_off = (size_t)P.slot_of[lane] * 2 * PIX;
        const T* __restrict__ src0 = img + (size_t)P.slot_x[lane] * C;
        // rows 0 .. n_early-1 were requested above
        const int n_early = min(n_stage, n_rows);
        int st = n_early == n_stage ? 0 : n_early;
        unsigned phase = n_early == n_stage ? 0u : 1u;   // toggles per lap: lap L >= 1 waits for the (L-1)-th release of its stage
        for (int i = n_early; i < n_rows; ++i) {
          mbar_wait(&f.empty[st], phase);
          // the expected byte count may be posted after some pieces have landed: the phase cannot complete before
          // this arrival
          if (lane == first) mbar_expect_tx(&f.full[st], slot_total);
          bulk_load_hint(ring + (size_t)st * row_bytes + dst_off, src0 + (size_t)P.rows[i] * W * C, bytes, &f.full[st], keep);
          if (++st == n_stage) { st = 0; phase ^= 1u; }
        }
      }
    } else if (tid < NC) {
      // ---------------- consumers: thread owns channels 2*tid, 2*tid+1
      const unsigned ring_s = smem_u32(ring) + (unsigned)tid * 2u * (unsigned)sizeof(T);
      const unsigned xs_s = smem_u32(smem + RING) + (unsigned)offsetof(RoiPlan, xs);
      const unsigned ay_s = smem_u32(smem + RING) + (unsigned)offsetof(RoiPlan, ay);
      int st = 0;
      unsigned phase = 0u;
      for (int i = 0; i < n_rows; ++i) {
        mbar_wait(&f.full[st], phase);
        const unsigned row = ring_s + (unsigned)st * row_bytes;
        float2 rx[kP];
#pragma unroll
        for (int p = 0; p < kP; ++p) rx[p] = make_float2(0.f, 0.f);
#pragma unroll
        for (int sx = 0; sx < NS; ++sx) {
          const float4 tap = lds_f4(xs_s + sx * 16u);        // off_lo, off_hi (bits), h, l
          const float2 v0 = lds_pair<T>(row + __float_as_uint(tap.x)), v1 = lds_pair<T>(row + __float_as_uint(tap.y));
          rx[sx / SR] = __ffma2_rn(make_float2(tap.w, tap.w), v1, __ffma2_rn(make_float2(tap.z, tap.z), v0, rx[sx / SR]));
        }
        const float4 a03 = lds_f4(ay_s + (unsigned)i * 32u), a46 = lds_f4(ay_s + (unsigned)i * 32u + 16u);
        const float ay[kP] = {a03.x, a03.y, a03.z, a03.w, a46.x, a46.y, a46.z};
#ifndef DGOD_FWD_DENSE_AY
        // a feature row carries weight for 1-2 of the 7 bin rows (all 7 only for RoIs under a few pixels): skip the others
#pragma unroll
        for (int ph = 0; ph < kP; ++ph) {
          if (ay[ph] != 0.f) {                                // warp-uniform (broadcast table read)
#pragma unroll
            for (int pw = 0; pw < kP; ++pw) acc[ph * kP + pw] = __ffma2_rn(make_float2(ay[ph], ay[ph]), rx[pw], acc[ph * kP + pw]);
          }
        }
#else
#pragma unroll
        for (int ph = 0; ph < kP; ++ph)
#pragma unroll
          for (int pw = 0; pw < kP; ++pw) acc[ph * kP + pw] = __ffma2_rn(make_float2(ay[ph], ay[ph]), rx[pw], acc[ph * kP + pw]);
#endif
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&f.empty[st]);       // this warp is done with stage st
        if (++st == n_stage) { st = 0; phase ^= 1u; }
      }
    }
  }
  __syncthreads();                          // every warp is done with the ring
  // pooled block -> shared [C][49] -> one bulk store
  if (tid < NC) {
    T* so = reinterpret_cast<T*>(ring) + (2 * tid) * kNB;
#pragma unroll
    for (int i = 0; i < kNB; ++i) {
      so[i] = from_f32<T>(acc[i].x * inv);
      so[kNB + i] = from_f32<T>(acc[i].y * inv);
    }
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    bulk_store_hint(out + (size_t)k * C * kNB, ring, (unsigned)(C * kNB * sizeof(T)), policy_evict_first());
    bulk_commit();
    bulk_wait_read<0>();
  }
}

// ================================================================================================
// Backward
// ================================================================================================
// Persistent, cooperative (all CTAs co-resident).  Per CTA: C/2 consumer threads (two adjacent
// channels each, the RoI's 49 gradient values of both in registers), one driver lane that streams
// plans / gradient blocks in and issues the row reductions, and one zero-fill warp.  Consumers and
// the driver are decoupled: a warp hands a finished row over through an mbarrier (row_full) and goes
// on with the next row as soon as the driver has published that the buffer it needs was read
// (rows_released) — there is no CTA-wide barrier in the row loop.
constexpr int kRowBars = 4;             // rows in flight per CTA (>= the largest number of row buffers)

struct alignas(16) BwdSync {
  alignas(8) unsigned long long plan_full[kPlanBufBwd], plan_empty[kPlanBufBwd];
  alignas(8) unsigned long long g_full, g_empty;
  alignas(8) unsigned long long row_full[kRowBars];
  volatile unsigned rows_released;      // rows 0 .. rows_released-1 of this CTA have been read by the bulk engine
  volatile int cur_img;                 // image the driver is working on (throttles the zero-fill warp)
};
constexpr int kZeroAhead = 2;           // the zero-fill warp runs at most this many images ahead (1: 10 % slower, 0: 75 %)

template <typename T, int C> struct BwdCfg {
  static constexpr int kStaging = C * kNB * (int)sizeof(T);                       // gradient block of one RoI
  static constexpr int kPlans = kPlanBufBwd * kPlanBwdBytes;
  static constexpr int kRowBytesMax = kSpanMax * C * (int)sizeof(T);
  // 2 CTAs per SM: (228 KB - 2 x 1 KB reserved) / 2 = 115 712 B of static + dynamic shared memory each
  static constexpr int kAvail = 115456 - kStaging - kPlans - (int)sizeof(BwdSync) - 320;
  static constexpr int kRing = (kAvail / 1024) * 1024 < kRowBars * kRowBytesMax ? (kAvail / 1024) * 1024 : kRowBars * kRowBytesMax;
  static constexpr int kSmem = kStaging + kPlans + kRing;
  static_assert(kRing >= kRowBytesMax, "the row ring must hold the widest row");
};

#ifdef DGOD_ROI_TIMING
#define TICK(n) do { t1 = clock64(); tacc[n] += t1 - t0; t0 = t1; } while (0)
#else
#define TICK(n) do {} while (0)
#endif

template <typename T, int C, int SR>
__global__ void __launch_bounds__(C / 2 + 64, 2)
msroi_bwd_tma_kernel(const RoiDev g, const RoiPlan* __restrict__ plans, const T* __restrict__ grad_out, int n_rois,
                     int* __restrict__ zero_done) {
  constexpr int NS = kP * SR;
  constexpr int NC = C / 2;             // consumer threads; then the driver warp and the zero-fill warp
  constexpr int PIX = C * (int)sizeof(T);
  constexpr int RING = BwdCfg<T, C>::kRing;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* staging = smem;                                          // [C][49] gradient block
  unsigned char* plan_buf = smem + BwdCfg<T, C>::kStaging;
  unsigned char* ring = plan_buf + BwdCfg<T, C>::kPlans;                  // row buffers
  __shared__ BwdSync f;
  const int tid = threadIdx.x;
  const int nj = (n_rois - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kPlanBufBwd; ++i) {
      mbar_init(&f.plan_full[i], 1);
      mbar_init(&f.plan_empty[i], NC / 32);
    }
    mbar_init(&f.g_full, 1);
    mbar_init(&f.g_empty, NC / 32);
#pragma unroll
    for (int i = 0; i < kRowBars; ++i) mbar_init(&f.row_full[i], NC / 32);
    mbar_fence_init();
    f.rows_released = 0u;
    f.cur_img = nj > 0 ? 0 : g.B;       // a CTA without RoIs never throttles its zero-fill warp
  }
  __syncthreads();

  if (tid == NC) {
    // ------------------------------------------------------------------ driver lane
    const unsigned long long stream = policy_evict_first();
    const unsigned long long keep = policy_evict_last();   // gradient maps of the images in flight stay in L2
    auto load_plan = [&](int j) {
      const int s = j % kPlanBufBwd;
      mbar_expect_tx(&f.plan_full[s], kPlanBwdBytes);
      bulk_load_hint(plan_buf + s * kPlanBwdBytes, plans + ((size_t)blockIdx.x + (size_t)j * gridDim.x), kPlanBwdBytes,
                     &f.plan_full[s], stream);
    };
    auto load_block = [&](int j) {
      mbar_expect_tx(&f.g_full, (unsigned)(C * kNB * sizeof(T)));
      bulk_load_hint(staging, grad_out + ((size_t)blockIdx.x + (size_t)j * gridDim.x) * C * kNB, (unsigned)(C * kNB * sizeof(T)),
                     &f.g_full, stream);
    };
    for (int j = 0; j < kPlanBufBwd && j < nj; ++j) load_plan(j);
    if (nj > 0) load_block(0);
    unsigned r = 0;                      // rows of this CTA, counted across RoIs
    int cur_img = -1;
    for (int j = 0; j < nj; ++j) {
      const int s = j % kPlanBufBwd;
      if (j + 1 < nj) {                  // consumers copy block j to registers first thing: refill the staging buffer
        mbar_wait(&f.g_empty, (unsigned)j & 1u);
        load_block(j + 1);
      }
      mbar_wait(&f.plan_full[s], (unsigned)(j / kPlanBufBwd) & 1u);
      const RoiPlan& P = *reinterpret_cast<const RoiPlan*>(plan_buf + s * kPlanBwdBytes);
      const int n_rows = P.n_rows;
      if (n_rows) {
        const int b = P.batch, slot_mode = P.slot_mode, W = g.W[P.level];
        const unsigned row_bytes = (unsigned)P.row_px * PIX;
        const int n_buf = min(kRowBars, RING / (int)row_bytes);
        if (b != cur_img) {              // the image's maps must be zero before the first reduction lands
          cur_img = b;
          f.cur_img = b;
          while (atomicAdd(&zero_done[b], 0) < (int)gridDim.x) __nanosleep(100);
          __threadfence();
        }
        T* __restrict__ img = reinterpret_cast<T*>(g.gfeat[P.level]) + (size_t)b * g.H[P.level] * W * C;
        for (int i = 0; i < n_rows; ++i, ++r) {
          mbar_wait(&f.row_full[r % kRowBars], (r / kRowBars) & 1u);
          unsigned char* buf = ring + (size_t)(i % n_buf) * row_bytes;
          const size_t yw = (size_t)P.rows[i] * W;
          if (!slot_mode) {
            bulk_reduce_add_hint<T>(img + (yw + P.x_first) * C, buf, row_bytes, keep);
          } else {
            for (int sx = 0; sx < NS; ++sx)
              if (P.slot_of[sx] >= 0)
                bulk_reduce_add_hint<T>(img + (yw + P.slot_x[sx]) * C, buf + (size_t)P.slot_of[sx] * 2 * PIX,
                                        P.slot_two[sx] ? 2u * PIX : (unsigned)PIX, keep);
          }
          bulk_commit();
          if (n_buf == 1) {              // a single buffer (widest rows only): the row just issued must be read
            bulk_wait_read<0>();
            f.rows_released = r + 1u;
          } else if (i + 1 < n_rows) {
            bulk_wait_read<1>();         // every row but the one just issued has been read
            f.rows_released = r;
          }
        }
        bulk_wait_read<0>();             // the next RoI lays its buffers out differently: drain
        f.rows_released = r;
      }
      if (j + kPlanBufBwd < nj) {        // recycle the plan slot once the consumers are done with it too
        mbar_wait(&f.plan_empty[s], (unsigned)(j / kPlanBufBwd) & 1u);
        load_plan(j + kPlanBufBwd);
      }
    }
    f.cur_img = g.B;                     // release the zero-fill warp for the remaining images
  } else if (tid >= NC + 32) {
    // ------------------------------------------------------------------ zero-fill warp
    // This CTA's share of every image's gradient maps, image by image, at most kZeroAhead images ahead
    // of the reductions; zero_done[b] counts the CTAs that finished image b.  All CTAs are co-resident
    // (cooperative launch), so a driver waiting for zero_done[b] == gridDim.x cannot deadlock.
    const int lane = tid - NC - 32;
    const unsigned long long keep = policy_evict_last();
    for (int b = 0; b < g.B; ++b) {
      while (f.cur_img < b - kZeroAhead) __nanosleep(200);
      for (int l = 0; l < g.n_levels; ++l) {
        const size_t n16 = (size_t)g.H[l] * g.W[l] * C * sizeof(T) / 16;
        uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<T*>(g.gfeat[l]) + (size_t)b * g.H[l] * g.W[l] * C);
        for (size_t i = (size_t)blockIdx.x * 32 + lane; i < n16; i += (size_t)gridDim.x * 32) st_zero16_hint(p + i, keep);
      }
      asm volatile("fence.proxy.async.global;\n" ::: "memory");   // generic-proxy zeros before the bulk engine's RMW
      __threadfence();
      __syncwarp();
      if (lane == 0) atomicAdd(&zero_done[b], 1);
    }
  } else if (tid < NC) {
    // ------------------------------------------------------------------ consumers: channels 2*tid, 2*tid+1
#ifdef DGOD_ROI_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t0 = clock64(), t1;
    const long long tstart = t0;
#endif
    unsigned r = 0, released = 0;
    const unsigned ring_s = smem_u32(ring), plan_s = smem_u32(plan_buf);
    for (int j = 0; j < nj; ++j) {
      const int s = j % kPlanBufBwd;
      mbar_wait(&f.plan_full[s], (unsigned)(j / kPlanBufBwd) & 1u);
      TICK(0);
      const RoiPlan& P = *reinterpret_cast<const RoiPlan*>(plan_buf + s * kPlanBwdBytes);
      const int n_rows = P.n_rows;
      // the gradient block: 49 values of this thread's two channels -> registers, pre-divided by count
      mbar_wait(&f.g_full, (unsigned)j & 1u);
      TICK(1);
      float2 gr[kNB];
      {
        const float inv = P.inv_count;
        const T* sg = reinterpret_cast<const T*>(staging) + (2 * tid) * kNB;
#pragma unroll
        for (int i = 0; i < kNB; ++i) gr[i] = make_float2(to_f32<T>(sg[i]) * inv, to_f32<T>(sg[kNB + i]) * inv);
      }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&f.g_empty);
      TICK(2);
      if (n_rows) {
        const int slot_mode = P.slot_mode, row_px = P.row_px;
        const unsigned row_bytes = (unsigned)row_px * PIX;
        const int n_buf = min(kRowBars, RING / (int)row_bytes);            // 1..4 row buffers of this RoI's row size
        const unsigned r0 = r;
        for (int i = 0; i < n_rows; ++i, ++r) {
          // ---- T: tq[pw] = sum_ph A_y[row][ph] * g[ph][pw]
          float2 tq[kP];
#pragma unroll
          for (int pw = 0; pw < kP; ++pw) tq[pw] = make_float2(0.f, 0.f);
          {
            const unsigned ayp = plan_s + (unsigned)s * kPlanBwdBytes + (unsigned)offsetof(RoiPlan, ay) + (unsigned)i * 32u;
            const float4 a03 = lds_f4(ayp), a46 = lds_f4(ayp + 16u);
            const float ay[kP] = {a03.x, a03.y, a03.z, a03.w, a46.x, a46.y, a46.z};
#pragma unroll
            for (int ph = 0; ph < kP; ++ph)
#pragma unroll
              for (int pw = 0; pw < kP; ++pw) tq[pw] = __ffma2_rn(make_float2(ay[ph], ay[ph]), gr[ph * kP + pw], tq[pw]);
          }
          TICK(3);
          // ---- the buffer this row goes to must have been read by the bulk engine
          {
            const unsigned need = i >= n_buf ? r - (unsigned)n_buf + 1u : r0;
            if (released < need) {
              while ((released = f.rows_released) < need) __nanosleep(40);
              __threadfence_block();
            }
          }
          TICK(4);
          const unsigned row = ring_s + (unsigned)(i % n_buf) * row_bytes + (unsigned)tid * 2u * (unsigned)sizeof(T);
          // ---- R: row[x] = sum_pw A_x[x][pw] * tq[pw]
          if (slot_mode) {
#pragma unroll
            for (int sx = 0; sx < NS; ++sx) {
              const int slot = P.slot_of[sx];
              if (slot >= 0) {                  // warp-uniform
                const float h = P.xs[sx].h, l = P.xs[sx].l;
                const bool two = P.slot_two[sx];
                const float w0 = two ? h : h + l;
                sts_pair<T>(row + (unsigned)(slot * 2) * PIX, __fmul2_rn(make_float2(w0, w0), tq[sx / SR]));
                if (two) sts_pair<T>(row + (unsigned)(slot * 2 + 1) * PIX, __fmul2_rn(make_float2(l, l), tq[sx / SR]));
              }
            }
          } else {
            // two pixels per step; the weights of the next two are in flight while these are computed and stored
            const unsigned axp = plan_s + (unsigned)s * kPlanBwdBytes + (unsigned)offsetof(RoiPlan, ax);
            float4 wa0 = lds_f4(axp), wb0 = lds_f4(axp + 16), wa1 = lds_f4(axp + 32), wb1 = lds_f4(axp + 48);
            for (int x = 0; x < row_px; x += 2) {
              const unsigned nx = axp + (unsigned)(x + 2) * 32u;            // rows >= span are zero (kSpanMax + 4 rows)
              const float4 na0 = lds_f4(nx), nb0 = lds_f4(nx + 16), na1 = lds_f4(nx + 32), nb1 = lds_f4(nx + 48);
              float2 va = __fmul2_rn(make_float2(wa0.x, wa0.x), tq[0]);
              float2 vb = __fmul2_rn(make_float2(wa0.y, wa0.y), tq[1]);
              float2 vc = __fmul2_rn(make_float2(wa1.x, wa1.x), tq[0]);
              float2 vd = __fmul2_rn(make_float2(wa1.y, wa1.y), tq[1]);
              va = __ffma2_rn(make_float2(wa0.z, wa0.z), tq[2], va);
              vb = __ffma2_rn(make_float2(wa0.w, wa0.w), tq[3], vb);
              vc = __ffma2_rn(make_float2(wa1.z, wa1.z), tq[2], vc);
              vd = __ffma2_rn(make_float2(wa1.w, wa1.w), tq[3], vd);
              va = __ffma2_rn(make_float2(wb0.x, wb0.x), tq[4], va);
              vb = __ffma2_rn(make_float2(wb0.y, wb0.y), tq[5], vb);
              vc = __ffma2_rn(make_float2(wb1.x, wb1.x), tq[4], vc);
              vd = __ffma2_rn(make_float2(wb1.y, wb1.y), tq[5], vd);
              va = __ffma2_rn(make_float2(wb0.z, wb0.z), tq[6], va);
              vc = __ffma2_rn(make_float2(wb1.z, wb1.z), tq[6], vc);
              sts_pair<T>(row + (unsigned)x * PIX, __fadd2_rn(va, vb));
              if (x + 1 < row_px) sts_pair<T>(row + (unsigned)(x + 1) * PIX, __fadd2_rn(vc, vd));
              wa0 = na0; wb0 = nb0; wa1 = na1; wb1 = nb1;
            }
          }
          TICK(5);
          fence_proxy_async_smem();             // generic-proxy writes -> visible to the bulk engine
          __syncwarp();
          if ((tid & 31) == 0) mbar_arrive(&f.row_full[r % kRowBars]);
          TICK(6);
        }
      }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&f.plan_empty[s]);
    }
#ifdef DGOD_ROI_TIMING
    if (tid == 0) {
      tacc[7] = clock64() - tstart;
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(zero_done) + 256;
      for (int q = 0; q < 8; ++q) atomicAdd(dst + q, (unsigned long long)tacc[q]);
    }
#endif
  }
}

// ------------------------------------------------------------------------------------------------
static bool tma_shape_ok(const RoiDev& g) {
  if (!g.channels_last || g.aligned || g.PH != kP || g.PW != kP) return false;   // DGOD: aligned=False (fasterrcnn.py:412-416)
  if (g.sr < 1 || g.sr > 2) return false;
  if (!(g.C == 256 || ((g.C == 128 || g.C == 64) && g.sr == 2))) return false;   // instantiated shapes
  if (g.B > kCounterBytes / (int)sizeof(int)) return false;
  for (int l = 0; l < g.n_levels; ++l)
    if (g.H[l] > 32000 || g.W[l] > 32000) return false;
  return true;
}

size_t msroi_tma_workspace(int n_rois) {
  return (size_t)kCounterBytes + (size_t)(n_rois > 0 ? n_rois : 1) * sizeof(RoiPlan);
}

static int launch_plan(const RoiDev& g, const float* rois, int n_rois, RoiPlan* plans, int pix_bytes, int want_ax, cudaStream_t st) {
  msroi_plan_kernel<<<cdiv(n_rois, kPlanWarps), kPlanWarps * 32, 0, st>>>(g, rois, n_rois, plans, pix_bytes, want_ax);
  DGOD_LAUNCHED();
  return DGOD_OK;
}

template <typename T, int C, int SR>
static int launch_fwd_tma(const RoiDev& g, const float* rois, int n_rois, void* out, RoiPlan* plans, cudaStream_t st) {
  constexpr int kSmem = FwdCfg<T, C>::kSmem;
  static bool attr = false;
  if (!attr) {
    DGOD_CUDA(cudaFuncSetAttribute(msroi_fwd_tma_kernel<T, C, SR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr = true;
  }
  int rc = launch_plan(g, rois, n_rois, plans, C * (int)sizeof(T), 0, st);
  if (rc) return rc;
  DGOD_CUDA(launch_pdl(msroi_fwd_tma_kernel<T, C, SR>, dim3(n_rois), dim3(C / 2 + 32), (size_t)kSmem, st, g, (const RoiPlan*)plans, n_rois,
                       (T*)out));
  DGOD_LAUNCHED();
  return DGOD_OK;
}

template <typename T>
static int dispatch_fwd_tma(const RoiDev& g, const float* rois, int n_rois, void* out, RoiPlan* plans, cudaStream_t st) {
  if (g.C == 256)
    return g.sr == 2 ? launch_fwd_tma<T, 256, 2>(g, rois, n_rois, out, plans, st) : launch_fwd_tma<T, 256, 1>(g, rois, n_rois, out, plans, st);
  if (g.C == 128) return launch_fwd_tma<T, 128, 2>(g, rois, n_rois, out, plans, st);
  return launch_fwd_tma<T, 64, 2>(g, rois, n_rois, out, plans, st);
}

int msroi_fwd_tma(const dgod_roi_config* cfg, const RoiDev& g, const float* rois, int n_rois, void* out, void* workspace,
                  size_t workspace_bytes, cudaStream_t st, int* handled) {
  *handled = 0;
  if (!tma_shape_ok(g) || ((uintptr_t)out & 15)) return DGOD_OK;
  if (!workspace || workspace_bytes < msroi_tma_workspace(n_rois) || ((uintptr_t)workspace & 127)) return DGOD_OK;
  for (int l = 0; l < g.n_levels; ++l)
    if ((uintptr_t)g.feat[l] & 15) return DGOD_OK;
  *handled = 1;
  RoiPlan* plans = reinterpret_cast<RoiPlan*>((char*)workspace + kCounterBytes);
  return cfg->dtype == DGOD_F32 ? dispatch_fwd_tma<float>(g, rois, n_rois, out, plans, st)
                                : dispatch_fwd_tma<__nv_bfloat16>(g, rois, n_rois, out, plans, st);
}

template <typename T, int C, int SR>
static int launch_bwd_tma(const RoiDev& g, const void* grad_out, const float* rois, int n_rois, void* workspace,
                          cudaStream_t st, int* handled) {
  constexpr int kSmem = BwdCfg<T, C>::kSmem;
  static bool init = false;
  static int coop = 0, n_cta = 0;
  if (!init) {
    int dev = 0, n_sm = 0, per_sm = 0;
    DGOD_CUDA(cudaGetDevice(&dev));
    DGOD_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    DGOD_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    DGOD_CUDA(cudaFuncSetAttribute(msroi_bwd_tma_kernel<T, C, SR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    DGOD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, msroi_bwd_tma_kernel<T, C, SR>, C / 2 + 64, kSmem));
    n_cta = n_sm * per_sm;
    init = true;
  }
  if (!coop || n_cta < 1) return DGOD_OK;                 // not handled: the caller falls back
  *handled = 1;
  int* zero_done = (int*)workspace;
  RoiPlan* plans = reinterpret_cast<RoiPlan*>((char*)workspace + kCounterBytes);
  DGOD_CUDA(cudaMemsetAsync(zero_done, 0, (size_t)g.B * sizeof(int), st));
  if (n_rois > 0) {
    int rc = launch_plan(g, rois, n_rois, plans, C * (int)sizeof(T), 1, st);
    if (rc) return rc;
  }
  // every CTA takes part in the zero fill (all of them are co-resident: cooperative launch)
  const T* go = (const T*)grad_out;
  const RoiPlan* cplans = plans;
  RoiDev gg = g;
  void* args[] = {(void*)&gg, (void*)&cplans, (void*)&go, (void*)&n_rois, (void*)&zero_done};
  DGOD_CUDA(cudaLaunchCooperativeKernel((const void*)msroi_bwd_tma_kernel<T, C, SR>, dim3(n_cta), dim3(C / 2 + 64), args, kSmem, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return DGOD_OK;
}

template <typename T>
static int dispatch_bwd_tma(const RoiDev& g, const void* grad_out, const float* rois, int n_rois, void* ws, cudaStream_t st,
                            int* handled) {
  if (g.C == 256)
    return g.sr == 2 ? launch_bwd_tma<T, 256, 2>(g, grad_out, rois, n_rois, ws, st, handled)
                     : launch_bwd_tma<T, 256, 1>(g, grad_out, rois, n_rois, ws, st, handled);
  if (g.C == 128) return launch_bwd_tma<T, 128, 2>(g, grad_out, rois, n_rois, ws, st, handled);
  return launch_bwd_tma<T, 64, 2>(g, grad_out, rois, n_rois, ws, st, handled);
}

int msroi_bwd_tma(const dgod_roi_config* cfg, const RoiDev& g, const void* grad_out, const float* rois, int n_rois,
                  void* workspace, size_t workspace_bytes, cudaStream_t st, int* handled) {
  *handled = 0;
  if (!tma_shape_ok(g) || ((uintptr_t)grad_out & 15)) return DGOD_OK;
  if (!workspace || workspace_bytes < msroi_tma_workspace(n_rois) || ((uintptr_t)workspace & 127)) return DGOD_OK;
  for (int l = 0; l < g.n_levels; ++l)
    if ((uintptr_t)g.gfeat[l] & 15) return DGOD_OK;
  return cfg->dtype == DGOD_F32 ? dispatch_bwd_tma<float>(g, grad_out, rois, n_rois, workspace, st, handled)
                                : dispatch_bwd_tma<__nv_bfloat16>(g, grad_out, rois, n_rois, workspace, st, handled);
}

}  // namespace dgod
