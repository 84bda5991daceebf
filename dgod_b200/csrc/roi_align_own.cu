// MultiScaleRoIAlign backward, owner-computes (torchvision::_roi_align_backward,
// the autograd of fasterrcnn.py:278 -> TV ops/poolers.py:147-227).
//
// The scatter formulation (one CTA per RoI adding its footprint to the gradient map, roi_align_tma.cu)
// moves every footprint through the L2's read-modify-write path: at B=8 x 512 RoIs that is 1.23 GB of
// reductions plus a 423 MB zero fill for a 423 MB map, and DRAM sees 2x the algorithmic bytes because
// zero-filled lines are evicted and re-read (profiles/r01_ncu_full_roi_align_tma.txt).  Here the map is
// the owner: a CTA owns a tile of 28 x 16 pixels x 64 channels of one image / level, keeps it in
// REGISTERS (warp w owns tile rows 2w, 2w+1; lane c owns channels c and c+32 of the slice: 2 x 16 x 2
// accumulators per thread), adds the contribution of every RoI whose footprint meets the tile and writes
// the tile exactly once, zeros included.  HBM traffic = grad_out read + gradient maps written once = the
// algorithmic bytes; no atomics, no memset, and the result is deterministic (RoIs are added in index
// order).
//
// Three kernels:
//   own_plan_kernel  one warp per RoI: level mapping, sampling taps (axis_tap(), i.e. TV's out-of-range
//                    skip and border clamp), the ascending lists of live feature rows / columns (<= 28
//                    each however large the RoI is) and the separable weight tables A_y[row][ph]/count,
//                    A_x[col][pw] -> a 1 920-byte plan; plus a 16-byte record (level, image, footprint box).
//   own_bin_kernel   one warp per tile: the RoIs (index order) whose footprint box meets the tile.
//   own_bwd_kernel   persistent, one CTA per SM.  A producer warp streams (plan, [64][49] gradient slice)
//                    pairs through a ring of stages with cp.async.bulk + mbarrier complete_tx, one lane per
//                    stage; 16 consumer warps are decoupled from each other (they own disjoint
//                    accumulators) and meet only at the per-stage empty barrier.  Per RoI and warp:
//                        T[pw]       = sum_ph A_y[row][ph] * g[ph][pw]        (rows of this warp, zero ph skipped)
//                        acc[row][x] += sum_pw A_x[x][pw] * T[pw]             (live columns of the tile)
//                    with packed FFMA2 on the lane's channel pair.
//
// Roofline: HBM.  Algorithmic bytes K*C*49*s + 20K + sum_l B*C*H_l*W_l*s.  On chip the
// gradient slices are re-read once per tile an RoI meets (about 3 at this tile size) from the L2.
#include <algorithm>
#include "roi_common.cuh"
#include "bulk.cuh"
#include "tmap.cuh"

#ifndef DGOD_OWN_TMAP
#define DGOD_OWN_TMAP 1   // 1: one elected lane issues tensor-map loads (UTMALDG); 0: one lane per stage issues 1-D bulk copies
#endif

namespace dgod {

namespace own {

constexpr int kP = 7;                  // pooled size (PH = PW = 7)
constexpr int kNB = kP * kP;
constexpr int kMaxSamp = 14;           // samples per axis (kP * sampling_ratio, sr <= 2)
constexpr int kMaxLive = 2 * kMaxSamp; // live rows / columns per axis
constexpr int kTileH = 28, kTileW = 16;   // 14 consumer warps + the producer warp = 15 warps: 128 registers per thread
constexpr int kCS = 64;                // channels per slice: lane c owns c and c + 32
constexpr int kWarps = kTileH / 2;     // consumer warps, two adjacent tile rows each
constexpr int kThreads = kWarps * 32 + 32;
constexpr int kCounterBytes = 8192;   // [0] pair cursor; +1024: per-CTA cycle counters of DGOD_OWN_TIMING builds

struct alignas(128) Plan {
  short n_rows, n_cols;                // 16-byte header, read by the consumers as one int4
  short y_first, y_last;               // extent of the live rows
  short x_first, x_last;               // extent of the live columns
  short span_mode, pad;                // 1: cols[] are the consecutive columns x_first .. x_last (<= 28 of them)
  short rows[kMaxLive];                // live feature rows, ascending (padding 0x7fff)
  short cols[kMaxLive];                // live feature columns, ascending (span mode: every column of the span)
  float ay[kMaxLive][8];               // A_y[row][ph] / count, list order
  float ax[kMaxLive][8];               // A_x[col][pw], list order (zero rows for columns without weight)
};
static_assert(sizeof(Plan) == 1920, "plans are moved with bulk copies");

struct alignas(16) Bin { int level, batch; short y0, y1, x0, x1; };   // level < 0: nothing to add
static_assert(sizeof(Bin) == 16, "read as one int4");

struct Tiles {
  int n_levels, n_slices, n_tiles, B;
  int lv[DGOD_MAX_LEVELS];             // processing order (coarsest level first: its tiles meet the most RoIs)
  int base[DGOD_MAX_LEVELS + 1];       // first tile id of the i-th processed level
  int ty[DGOD_MAX_LEVELS], tx[DGOD_MAX_LEVELS];
};

struct TileAt { int level, b, y0, x0; };
__device__ __forceinline__ TileAt tile_at(const Tiles& tg, int t) {
  int i = 0;
  while (i + 1 < tg.n_levels && t >= tg.base[i + 1]) ++i;
  const int local = t - tg.base[i], per = tg.ty[i] * tg.tx[i];
  const int b = local / per, r = local - b * per;
  TileAt a;
  a.level = tg.lv[i];
  a.b = b;
  a.y0 = (r / tg.tx[i]) * kTileH;
  a.x0 = (r % tg.tx[i]) * kTileW;
  return a;
}

// ---------------------------------------------------------------- plan kernel (one warp per RoI)
struct Scratch {
  RoiGeom geo;
  short lo[2][kMaxSamp], hi[2][kMaxSamp];     // axis 0 = y, 1 = x
  float l[2][kMaxSamp], h[2][kMaxSamp];       // both zero: sample skipped
  short list[2][kMaxLive];
  int n[2];
};

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
own_plan_kernel(const RoiDev g, const float* __restrict__ rois, int n_rois, Plan* __restrict__ plans, Bin* __restrict__ bins) {
  __shared__ Scratch scratch[kPlanWarps];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int k = blockIdx.x * kPlanWarps + w;
  if (k >= n_rois) return;
  Scratch& t = scratch[w];
  const int sr = g.sr, ns = kP * sr;
  if (lane == 0) t.geo = roi_geometry(g, rois + (size_t)k * 5);
  __syncwarp();
  const RoiGeom r = t.geo;
  const bool usable = r.batch >= 0 && r.batch < g.B;
  int n_rows = 0, n_cols = 0;
  if (usable) {
    {  // lanes 0..13: y samples, lanes 16..29: x samples
      const int ax = lane >> 4, s = lane & 15;
      if (s < ns) {
        const AxisTap a = ax ? axis_tap(sample_coord(r.start_w, s / sr, r.bin_w, s % sr, sr), r.W)
                             : axis_tap(sample_coord(r.start_h, s / sr, r.bin_h, s % sr, sr), r.H);
        t.lo[ax][s] = (short)a.lo;
        t.hi[ax][s] = (short)a.hi;
        t.l[ax][s] = a.valid ? a.l : 0.f;
        t.h[ax][s] = a.valid ? a.h : 0.f;
      }
    }
    __syncwarp();
#pragma unroll
    for (int ax = 0; ax < 2; ++ax) {   // ascending list of the coordinates that receive weight
      const int s = lane >> 1, is_hi = lane & 1;
      bool valid = false;
      int c = 0;
      if (s < ns) {
        const bool live = t.h[ax][s] != 0.f || t.l[ax][s] != 0.f;
        valid = live && (is_hi ? (t.l[ax][s] != 0.f && t.hi[ax][s] != t.lo[ax][s]) : true);
        c = is_hi ? t.hi[ax][s] : t.lo[ax][s];
      }
      const unsigned same = __match_any_sync(0xffffffffu, valid ? c : (0x10000 + lane));
      const bool first = valid && (__ffs(same) - 1 == lane);
      int rank = 0;
      for (int j = 0; j < 2 * ns; ++j) {
        const int cj = __shfl_sync(0xffffffffu, c, j);
        const int fj = __shfl_sync(0xffffffffu, (int)first, j);
        rank += (fj && cj < c) ? 1 : 0;
      }
      if (first) t.list[ax][rank] = (short)c;
      const unsigned m = __ballot_sync(0xffffffffu, first);
      if (lane == 0) t.n[ax] = __popc(m);
    }
    __syncwarp();
    n_rows = t.n[0];
    n_cols = t.n[1];
    if (n_rows == 0 || n_cols == 0) n_rows = n_cols = 0;
  }
  Plan& P = plans[k];
  const int x_first = n_rows ? t.list[1][0] : 0, x_last = n_rows ? t.list[1][n_cols - 1] : -1;
  const int span = x_last - x_first + 1;
  const int span_mode = n_rows && span <= kMaxLive;
  const int n_cols_out = span_mode ? span : n_cols;
  if (lane == 0) {
    P.n_rows = (short)n_rows; P.n_cols = (short)n_cols_out;
    P.y_first = n_rows ? t.list[0][0] : (short)0; P.y_last = n_rows ? t.list[0][n_rows - 1] : (short)-1;
    P.x_first = (short)x_first; P.x_last = (short)x_last;
    P.span_mode = (short)span_mode; P.pad = 0;
    Bin b;
    b.level = n_rows ? r.level : -1;
    b.batch = usable ? r.batch : -1;
    b.y0 = P.y_first; b.y1 = P.y_last; b.x0 = (short)x_first; b.x1 = (short)x_last;
    bins[k] = b;
  }
  if (!n_rows) return;
  if (lane < kMaxLive) {
    P.rows[lane] = lane < n_rows ? t.list[0][lane] : (short)0x7fff;
    P.cols[lane] = lane < n_cols_out ? (span_mode ? (short)(x_first + lane) : t.list[1][lane]) : (short)0x7fff;
  }
  const float inv = 1.f / r.count;             // count = sr*sr: a power of two, so scaling the table is exact
  for (int e = lane; e < kMaxLive * 8; e += 32) {
    const int i = e >> 3, p = e & 7;
    P.ay[i][p] = (i < n_rows && p < kP) ? axis_weight(t.lo[0], t.hi[0], t.l[0], t.h[0], t.list[0][i], p, sr) * inv : 0.f;
    const int col = span_mode ? x_first + i : (int)t.list[1][i < n_cols ? i : 0];
    P.ax[i][p] = (i < n_cols_out && p < kP) ? axis_weight(t.lo[1], t.hi[1], t.l[1], t.h[1], col, p, sr) : 0.f;
  }
}

// ---------------------------------------------------------------- bin kernel (one warp per tile)
constexpr int kBinWarps = 8;

__device__ __forceinline__ bool bin_hit(const Bin* __restrict__ bins, int k, int k_end, const TileAt& a) {
  if (k >= k_end) return false;
  const int4 v = __ldg(reinterpret_cast<const int4*>(bins + k));
  const int y0 = (short)(v.z & 0xffff), y1 = (short)(v.z >> 16), x0 = (short)(v.w & 0xffff), x1 = (short)(v.w >> 16);
  return v.x == a.level && v.y == a.b && y0 < a.y0 + kTileH && y1 >= a.y0 && x0 < a.x0 + kTileW && x1 >= a.x0;
}

__global__ void __launch_bounds__(kBinWarps * 32)
own_bin_kernel(const Tiles tg, const Bin* __restrict__ bins, int n_rois, const int32_t* __restrict__ roi_img_offsets,
               int2* __restrict__ tile_list, int* __restrict__ pair_k, int* __restrict__ cursor) {
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * kBinWarps + (threadIdx.x >> 5);
  if (t >= tg.n_tiles) return;
  const TileAt a = tile_at(tg, t);
  int k_begin = 0, k_end = n_rois;
  if (roi_img_offsets) { k_begin = roi_img_offsets[a.b]; k_end = roi_img_offsets[a.b + 1]; }
  int cnt = 0;
  for (int base = k_begin; base < k_end; base += 32) cnt += __popc(__ballot_sync(0xffffffffu, bin_hit(bins, base + lane, k_end, a)));
  int off = 0;
  if (lane == 0) {
    off = cnt ? atomicAdd(cursor, cnt) : 0;
    tile_list[t] = make_int2(off, cnt);
  }
  if (!cnt) return;
  off = __shfl_sync(0xffffffffu, off, 0);
  for (int base = k_begin; base < k_end; base += 32) {
    const bool hit = bin_hit(bins, base + lane, k_end, a);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (hit) pair_k[off + __popc(m & ((1u << lane) - 1u))] = base + lane;
    off += __popc(m);
  }
}

// ---------------------------------------------------------------- main kernel
template <typename T> struct Cfg {
  static constexpr int kGBytes = kCS * kNB * (int)sizeof(T);
  static constexpr int kStageBytes = (int)sizeof(Plan) + kGBytes;
  static constexpr int kStagesFit = (227 * 1024 - 1024) / kStageBytes;
  static constexpr int kStages = kStagesFit > 24 ? 24 : kStagesFit;
  static constexpr int kSmem = kStages * kStageBytes;
  // the tensor maps see grad_out / the plans as rows of 32-bit words
  static constexpr int kGRowWords = 196, kGBoxRows = kGBytes / (kGRowWords * 4);       // [64][49] slice = 16 (fp32) / 8 (bf16) rows
  static constexpr int kPlanRowWords = 240, kPlanBoxRows = (int)sizeof(Plan) / (kPlanRowWords * 4);
  static_assert(kGBoxRows * kGRowWords * 4 == kGBytes && kPlanBoxRows * kPlanRowWords * 4 == (int)sizeof(Plan), "box shapes");
  static_assert(kStageBytes % 128 == 0 && kStages >= 4 && kStages <= 32, "stage ring");
};

template <typename T> __device__ __forceinline__ float lds_as_f32(const T* p);
template <> __device__ __forceinline__ float lds_as_f32<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float lds_as_f32<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __uint_as_float((unsigned)(*reinterpret_cast<const unsigned short*>(p)) << 16);
}

// acc{0,1}[x] += sum_pw w[pw] * t{0,1}[pw] for the two rows of this warp (packed pairs of channels)
__device__ __forceinline__ void column_update(const float4 w0, const float4 w1, const float2 (&t0)[kP], const float2 (&t1)[kP],
                                              float2& a0, float2& a1) {
  a0 = __ffma2_rn(make_float2(w0.x, w0.x), t0[0], a0); a1 = __ffma2_rn(make_float2(w0.x, w0.x), t1[0], a1);
  a0 = __ffma2_rn(make_float2(w0.y, w0.y), t0[1], a0); a1 = __ffma2_rn(make_float2(w0.y, w0.y), t1[1], a1);
  a0 = __ffma2_rn(make_float2(w0.z, w0.z), t0[2], a0); a1 = __ffma2_rn(make_float2(w0.z, w0.z), t1[2], a1);
  a0 = __ffma2_rn(make_float2(w0.w, w0.w), t0[3], a0); a1 = __ffma2_rn(make_float2(w0.w, w0.w), t1[3], a1);
  a0 = __ffma2_rn(make_float2(w1.x, w1.x), t0[4], a0); a1 = __ffma2_rn(make_float2(w1.x, w1.x), t1[4], a1);
  a0 = __ffma2_rn(make_float2(w1.y, w1.y), t0[5], a0); a1 = __ffma2_rn(make_float2(w1.y, w1.y), t1[5], a1);
  a0 = __ffma2_rn(make_float2(w1.z, w1.z), t0[6], a0); a1 = __ffma2_rn(make_float2(w1.z, w1.z), t1[6], a1);
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
own_bwd_kernel(const RoiDev g, const Tiles tg, const __grid_constant__ CUtensorMap tm_plan, const __grid_constant__ CUtensorMap tm_g,
               const Plan* __restrict__ plans, const T* __restrict__ grad_out, const int2* __restrict__ tile_list,
               const int* __restrict__ pair_k, long long* __restrict__ timing) {
  constexpr int NS = Cfg<T>::kStages;
  constexpr int SB = Cfg<T>::kStageBytes;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ alignas(8) unsigned long long full[NS], empty[NS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_items = tg.n_tiles * tg.n_slices;
  const int C = g.C;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kWarps);
    }
    mbar_fence_init();
#if DGOD_OWN_TMAP
    tmap_prefetch(&tm_plan);
    tmap_prefetch(&tm_g);
#endif
  }
  __syncthreads();

  if (warp == kWarps) {
#if DGOD_OWN_TMAP
    // ------------------------------------------------------------------ producer: the warp walks the pair lists together
    // (32 RoI indices per coalesced read, the next item's first chunk prefetched); lane 0 issues the two tensor-map
    // loads of a stage in order, as far ahead as the ring allows.
    const int rows_per_roi = C / kCS * Cfg<T>::kGBoxRows;
    unsigned q = 0;
    int it = blockIdx.x;
    int2 lc = make_int2(0, 0);
    int kk = 0;
    if (it < n_items) {
      lc = __ldg(tile_list + it / tg.n_slices);
      kk = lane < lc.y ? __ldg(pair_k + lc.x + lane) : 0;
    }
    while (it < n_items) {
      const int sl = it % tg.n_slices;
      const int it_next = it + (int)gridDim.x;
      int2 lc_next = make_int2(0, 0);
      int kk_next = 0;
      if (it_next < n_items) {
        lc_next = __ldg(tile_list + it_next / tg.n_slices);
        kk_next = lane < lc_next.y ? __ldg(pair_k + lc_next.x + lane) : 0;
      }
      for (int p0 = 0; p0 < lc.y; p0 += 32) {
        if (p0) kk = p0 + lane < lc.y ? __ldg(pair_k + lc.x + p0 + lane) : 0;
        const int n = min(32, lc.y - p0);
        for (int i = 0; i < n; ++i, ++q) {
          const int k = __shfl_sync(0xffffffffu, kk, i);
          if (lane == 0) {
            const unsigned s = q % NS;
            unsigned char* st = smem + (size_t)s * SB;
            if (q >= (unsigned)NS) mbar_wait(&empty[s], ((q / NS) - 1u) & 1u);
            mbar_expect_tx(&full[s], (unsigned)SB);
            tmap_load_2d(st, &tm_plan, 0, k * Cfg<T>::kPlanBoxRows, &full[s]);
            tmap_load_2d(st + sizeof(Plan), &tm_g, 0, k * rows_per_roi + sl * Cfg<T>::kGBoxRows, &full[s]);
          }
        }
        __syncwarp();
      }
      it = it_next; lc = lc_next; kk = kk_next;
    }
#else
    // ------------------------------------------------------------------ producer: lane s owns ring stage s
    if (lane < NS) {
      unsigned char* st = smem + (size_t)lane * SB;
      unsigned q_base = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int t = it / tg.n_slices, sl = it - t * tg.n_slices;
        const int2 lc = __ldg(tile_list + t);
        const T* __restrict__ gsl = grad_out + (size_t)sl * kCS * kNB;
        int p = (int)((unsigned)(lane + NS - (int)(q_base % NS)) % NS);
        for (; p < lc.y; p += NS) {
          const unsigned q = q_base + (unsigned)p;
          if (q >= (unsigned)NS) mbar_wait(&empty[lane], ((q / NS) - 1u) & 1u);
          const int k = __ldg(pair_k + lc.x + p);
          mbar_expect_tx(&full[lane], (unsigned)SB);
          bulk_load(st, plans + k, (unsigned)sizeof(Plan), &full[lane]);
          bulk_load(st + sizeof(Plan), gsl + (size_t)k * C * kNB, (unsigned)Cfg<T>::kGBytes, &full[lane]);
        }
        q_base += (unsigned)lc.y;
      }
    }
#endif
    return;
  }

  // -------------------------------------------------------------------- consumers: tile rows 2*warp, 2*warp+1
  const int r0 = 2 * warp;
  unsigned q = 0;
#ifdef DGOD_OWN_TIMING
  const long long t_start = clock64();
  long long t_wait = 0, t_store = 0;
#endif
  for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
    const int t = it / tg.n_slices, sl = it - t * tg.n_slices;
    const TileAt a = tile_at(tg, t);
    const int2 lc = __ldg(tile_list + t);
    const int y_mine = a.y0 + r0;
    float2 acc0[kTileW], acc1[kTileW];
#pragma unroll
    for (int x = 0; x < kTileW; ++x) acc0[x] = acc1[x] = make_float2(0.f, 0.f);

    for (int p = 0; p < lc.y; ++p, ++q) {
      const unsigned s = q % NS;
#ifdef DGOD_OWN_TIMING
      const long long tw = clock64();
#endif
      mbar_wait(&full[s], (q / NS) & 1u);
#ifdef DGOD_OWN_TIMING
      t_wait += clock64() - tw;
#endif
      const unsigned char* st = smem + (size_t)s * SB;
      const Plan& P = *reinterpret_cast<const Plan*>(st);
      const int4 hdr = *reinterpret_cast<const int4*>(st);
      const int n_cols = hdr.x >> 16;
      const int y_first = (short)(hdr.y & 0xffff), y_last = hdr.y >> 16;
      const int x_first = (short)(hdr.z & 0xffff), x_last = hdr.z >> 16;
      const int span_mode = hdr.w & 0xffff;
      // quick reject (warp-uniform): this warp's two rows or the tile's columns are outside the footprint box
      if (y_mine + 1 >= y_first && y_mine <= y_last && x_last >= a.x0 && x_first < a.x0 + kTileW) {
        // tile-relative coordinate of live row `lane`
        const int rel_r = lane < kMaxLive ? (int)P.rows[lane] - a.y0 : 0x7fff;
        const unsigned m0 = __ballot_sync(0xffffffffu, rel_r == r0), m1 = __ballot_sync(0xffffffffu, rel_r == r0 + 1);
        if (m0 | m1) {
          float ay0[8], ay1[8];
          unsigned phm;                                    // bins whose samples put weight on either row (ballot: provably uniform)
          {
            const int i0 = m0 ? __ffs(m0) - 1 : 0, i1 = m1 ? __ffs(m1) - 1 : 0;
            const float4* pa = reinterpret_cast<const float4*>(P.ay[i0]);
            const float4* pb = reinterpret_cast<const float4*>(P.ay[i1]);
            const float e0 = P.ay[i0][lane & 7], e1 = P.ay[i1][lane & 7];
            phm = __ballot_sync(0xffffffffu, lane < kP && ((m0 && e0 != 0.f) || (m1 && e1 != 0.f)));
            const float4 u0 = pa[0], u1 = pa[1], v0 = pb[0], v1 = pb[1];
            const float z0 = m0 ? 1.f : 0.f, z1 = m1 ? 1.f : 0.f;
            ay0[0] = u0.x * z0; ay0[1] = u0.y * z0; ay0[2] = u0.z * z0; ay0[3] = u0.w * z0;
            ay0[4] = u1.x * z0; ay0[5] = u1.y * z0; ay0[6] = u1.z * z0;
            ay1[0] = v0.x * z1; ay1[1] = v0.y * z1; ay1[2] = v0.z * z1; ay1[3] = v0.w * z1;
            ay1[4] = v1.x * z1; ay1[5] = v1.y * z1; ay1[6] = v1.z * z1;
          }
          float2 t0[kP], t1[kP];
#pragma unroll
          for (int pw = 0; pw < kP; ++pw) t0[pw] = t1[pw] = make_float2(0.f, 0.f);
          const T* __restrict__ ga = reinterpret_cast<const T*>(st + sizeof(Plan)) + lane * kNB;
          const T* __restrict__ gb = ga + 32 * kNB;
#pragma unroll
          for (int ph = 0; ph < kP; ++ph) {
            if ((phm >> ph) & 1u) {                         // warp-uniform
#pragma unroll
              for (int pw = 0; pw < kP; ++pw) {
                const float2 gv = make_float2(lds_as_f32<T>(ga + ph * kP + pw), lds_as_f32<T>(gb + ph * kP + pw));
                t0[pw] = __ffma2_rn(make_float2(ay0[ph], ay0[ph]), gv, t0[pw]);
                t1[pw] = __ffma2_rn(make_float2(ay1[ph], ay1[ph]), gv, t1[pw]);
              }
            }
          }
          if (span_mode) {
            // the live columns are consecutive: table row of tile column x is x + shift, a fixed offset per pair
            const int shift = a.x0 - x_first;
            const int xa = max(0, -shift), xb = min(kTileW - 1, n_cols - 1 - shift);
            const unsigned cm = __ballot_sync(0xffffffffu, lane >= xa && lane <= xb);   // through a vote: provably uniform
            const float4* base = reinterpret_cast<const float4*>(&P.ax[0][0]) + 2 * shift;
#pragma unroll
            for (int x = 0; x < kTileW; ++x) {
              if ((cm >> x) & 1u) {                         // warp-uniform
                const float4 w0 = base[2 * x], w1 = base[2 * x + 1];
                column_update(w0, w1, t0, t1, acc0[x], acc1[x]);
              }
            }
          } else {
            const int rel_c = lane < kMaxLive ? (int)P.cols[lane] - a.x0 : 0x7fff;
            const unsigned cm = __reduce_or_sync(0xffffffffu, (rel_c >= 0 && rel_c < kTileW) ? (1u << rel_c) : 0u);
            const int j_first = __popc(__ballot_sync(0xffffffffu, rel_c < 0));
#pragma unroll
            for (int x = 0; x < kTileW; ++x) {
              if ((cm >> x) & 1u) {                         // warp-uniform
                const int j = j_first + __popc(cm & ((1u << x) - 1u));
                const float4* px = reinterpret_cast<const float4*>(P.ax[j]);
                column_update(px[0], px[1], t0, t1, acc0[x], acc1[x]);
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }

    // ---- the tile leaves once, zeros included
#ifdef DGOD_OWN_TIMING
    const long long ts = clock64();
#endif
    const int H = g.H[a.level], W = g.W[a.level];
    T* __restrict__ img = reinterpret_cast<T*>(g.gfeat[a.level]) + (size_t)a.b * H * W * C + (size_t)sl * kCS + lane;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int y = a.y0 + r0 + rr;
      if (y < H) {
        T* __restrict__ rowp = img + (size_t)y * W * C;
#pragma unroll
        for (int x = 0; x < kTileW; ++x) {
          const int X = a.x0 + x;
          if (X < W) {
            const float2 v = rr ? acc1[x] : acc0[x];
            rowp[(size_t)X * C] = from_f32<T>(v.x);
            rowp[(size_t)X * C + 32] = from_f32<T>(v.y);
          }
        }
      }
    }
#ifdef DGOD_OWN_TIMING
    t_store += clock64() - ts;
#endif
  }
#ifdef DGOD_OWN_TIMING
  if (lane == 0 && (warp == 0 || warp == kWarps - 1)) {      // two warps per CTA: total, waiting for a stage, storing, pairs
    long long* o = timing + ((size_t)blockIdx.x * 2 + (warp ? 1 : 0)) * 4;
    o[0] = clock64() - t_start; o[1] = t_wait; o[2] = t_store; o[3] = q;
  }
#endif
}

}  // namespace own

// ------------------------------------------------------------------------------------------------
static bool own_shape_ok(const RoiDev& g) {
  if (!g.channels_last || g.PH != own::kP || g.PW != own::kP) return false;
  if (g.sr < 1 || g.sr > 2) return false;
  if (g.C % own::kCS != 0) return false;
  for (int l = 0; l < g.n_levels; ++l)
    if (g.H[l] > 32000 || g.W[l] > 32000) return false;
  return true;
}

static own::Tiles own_tiles(const RoiDev& g) {
  own::Tiles tg;
  tg.n_levels = g.n_levels;
  tg.n_slices = g.C / own::kCS;
  tg.B = g.B;
  // coarsest level (smallest map) first
  int order[DGOD_MAX_LEVELS];
  for (int l = 0; l < g.n_levels; ++l) order[l] = l;
  for (int i = 1; i < g.n_levels; ++i)
    for (int j = i; j > 0 && (long long)g.H[order[j]] * g.W[order[j]] < (long long)g.H[order[j - 1]] * g.W[order[j - 1]]; --j) {
      const int tmp = order[j]; order[j] = order[j - 1]; order[j - 1] = tmp;
    }
  int total = 0;
  for (int i = 0; i < g.n_levels; ++i) {
    const int l = order[i];
    tg.lv[i] = l;
    tg.ty[i] = (g.H[l] + own::kTileH - 1) / own::kTileH;
    tg.tx[i] = (g.W[l] + own::kTileW - 1) / own::kTileW;
    tg.base[i] = total;
    total += g.B * tg.ty[i] * tg.tx[i];
  }
  for (int i = g.n_levels; i <= DGOD_MAX_LEVELS; ++i) tg.base[i] = total;
  for (int i = g.n_levels; i < DGOD_MAX_LEVELS; ++i) tg.lv[i] = tg.ty[i] = tg.tx[i] = 0;
  tg.n_tiles = total;
  return tg;
}

struct OwnLayout { size_t bins, tile_list, pair_k, plans, total; };

static OwnLayout own_layout(const own::Tiles& tg, int n_rois) {
  const size_t K = n_rois > 0 ? n_rois : 1;
  size_t per_roi = 1;                  // tiles one footprint box can meet: at most every tile of its level
  for (int i = 0; i < tg.n_levels; ++i) per_roi = std::max(per_roi, (size_t)tg.ty[i] * tg.tx[i]);
  OwnLayout L;
  size_t off = own::kCounterBytes;
  L.bins = off; off = align_up(off + K * sizeof(own::Bin), 256);
  L.tile_list = off; off = align_up(off + (size_t)tg.n_tiles * sizeof(int2), 256);
  L.pair_k = off; off = align_up(off + K * per_roi * sizeof(int), 256);
  L.plans = off; off = align_up(off + K * sizeof(own::Plan), 256);
  L.total = off;
  return L;
}

size_t msroi_own_workspace(const RoiDev& g, int n_rois) {
  if (!own_shape_ok(g)) return 0;
  return own_layout(own_tiles(g), n_rois).total;
}

template <typename T>
static int launch_own(const RoiDev& g, const own::Tiles& tg, const OwnLayout& L, const void* grad_out, const float* rois,
                      int n_rois, const int32_t* roi_img_offsets, char* ws, cudaStream_t st) {
  static bool init = false;
  static int n_sm = 0;
  if (!init) {
    int dev = 0;
    DGOD_CUDA(cudaGetDevice(&dev));
    DGOD_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    DGOD_CUDA(cudaFuncSetAttribute(own::own_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, own::Cfg<T>::kSmem));
    init = true;
  }
  int* cursor = reinterpret_cast<int*>(ws);
  own::Bin* bins = reinterpret_cast<own::Bin*>(ws + L.bins);
  int2* tile_list = reinterpret_cast<int2*>(ws + L.tile_list);
  int* pair_k = reinterpret_cast<int*>(ws + L.pair_k);
  own::Plan* plans = reinterpret_cast<own::Plan*>(ws + L.plans);
  DGOD_CUDA(cudaMemsetAsync(cursor, 0, own::kCounterBytes, st));
  if (n_rois > 0) {
    own::own_plan_kernel<<<cdiv(n_rois, own::kPlanWarps), own::kPlanWarps * 32, 0, st>>>(g, rois, n_rois, plans, bins);
    DGOD_LAUNCHED();
  }
  own::own_bin_kernel<<<cdiv(tg.n_tiles, own::kBinWarps), own::kBinWarps * 32, 0, st>>>(tg, bins, n_rois, roi_img_offsets, tile_list,
                                                                                      pair_k, cursor);
  DGOD_LAUNCHED();
  const int n_items = tg.n_tiles * tg.n_slices;
  const int grid = n_items < n_sm ? n_items : n_sm;
  CUtensorMap tm_plan, tm_g;
  memset(&tm_plan, 0, sizeof(tm_plan));
  memset(&tm_g, 0, sizeof(tm_g));
#if DGOD_OWN_TMAP
  {
    using CF = own::Cfg<T>;
    const unsigned long long K = n_rois > 0 ? n_rois : 1;
    int rc = encode_words_2d(&tm_plan, plans, K * CF::kPlanBoxRows, CF::kPlanRowWords, CF::kPlanBoxRows);
    if (rc) return rc;
    // grad_out may be null when there are no RoIs: the map is never dereferenced then, but must still encode
    rc = encode_words_2d(&tm_g, n_rois > 0 ? grad_out : (const void*)plans, K * (g.C / own::kCS) * CF::kGBoxRows, CF::kGRowWords,
                         CF::kGBoxRows);
    if (rc) return rc;
  }
#endif
  own::own_bwd_kernel<T><<<grid, own::kThreads, own::Cfg<T>::kSmem, st>>>(g, tg, tm_plan, tm_g, plans, (const T*)grad_out, tile_list,
                                                                       pair_k, reinterpret_cast<long long*>(ws + 1024));
  DGOD_LAUNCHED();
  return DGOD_OK;
}

int msroi_bwd_own(const dgod_roi_config* cfg, const RoiDev& g, const void* grad_out, const float* rois, int n_rois,
                  const int32_t* roi_img_offsets, void* workspace, size_t workspace_bytes, cudaStream_t st, int* handled) {
  *handled = 0;
  if (!own_shape_ok(g) || ((uintptr_t)grad_out & 15)) return DGOD_OK;
  if (!workspace || ((uintptr_t)workspace & 127)) return DGOD_OK;
  const own::Tiles tg = own_tiles(g);
  if (tg.n_tiles <= 0) return DGOD_OK;
  const OwnLayout L = own_layout(tg, n_rois);
  if (workspace_bytes < L.total) return DGOD_OK;
  *handled = 1;
  return cfg->dtype == DGOD_F32 ? launch_own<float>(g, tg, L, grad_out, rois, n_rois, roi_img_offsets, (char*)workspace, st)
                                : launch_own<__nv_bfloat16>(g, tg, L, grad_out, rois, n_rois, roi_img_offsets, (char*)workspace, st);
}

}  // namespace dgod
