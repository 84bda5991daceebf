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
//                    each however large the RoI is), the separable weight tables A_y[row][ph]/count and
//                    A_x[col][pw], and the BLOCKED form of A_x: absolute 4-column blocks, each with a base
//                    bin such that all its weights sit on 4 consecutive bins (3 of 4 RoIs have one) ->
//                    a 2 560-byte plan; plus a 16-byte record (level, image, footprint box).
//   own_bin_kernel   one warp per tile: the RoIs (index order) whose footprint box meets the tile.
//   own_bwd_kernel   persistent, one CTA per SM, 16 warps, work items (tile x 64-channel slice), coarse levels first:
//                    dealt round-robin (algo 4, the default: best L2 locality, 267 us in a single-GPU step) or claimed
//                    from a global counter (algo 5, kClaim: a CTA whose SM is still held by another stream's kernel —
//                    an NCCL all-reduce overlapping backward — just takes fewer items: 277 vs 342 us in a 2-GPU step,
//                    291 us in a single-GPU one).  Warp roles:
//                      producer  lane 0 streams (plan, [64][49] gradient slice) pairs through a ring of 15
//                                stages with two tensor-map loads per stage (cp.async.bulk.tensor, SASS
//                                UTMALDG; completion counted in bytes on the stage's `full` mbarrier);
//                      decoder   once per pair (not once per consumer warp) intersects the RoI's live
//                                rows / columns with the tile and publishes masks, list offsets and the
//                                tile's 4 block descriptors in shared memory (`ready` mbarrier); for a plan
//                                without a blocked form it tries to build one for this tile;
//                      14 consumers, decoupled from each other (they own disjoint accumulators), meet only
//                                at the per-stage `empty` barrier.  Per pair and warp:
//                                  T[pw]       = sum_ph A_y[row][ph] * g[ph][pw]   (zero bins skipped)
//                                  acc[row][x] += sum_pw A_x[x][pw] * T[pw]
//                                the second line per block of 4 columns x 4 bins, branch-free, packed FFMA2
//                                on the lane's channel pair (general form: 7 bins per live column).
//
// Roofline: HBM.  Algorithmic bytes K*C*49*s + 20K + sum_l B*C*H_l*W_l*s; measured DRAM
// traffic 0.98x of that (profiles/).  On chip the gradient slices are re-read once per tile an RoI meets
// (2.8 at this tile size) from the L2.  What bounds the kernel today is instruction issue at low warp-level parallelism
// (16 warps x 128 registers per SM; profiles/README.md has the ablation — of 293 us: T phase 37, block sweeps 101, stores
// 28, ring + decode + loop skeleton 157 — the ncu stall profile and the variants that were measured and not kept).
// Compile-time switches for those ablations: DGOD_OWN_SKIP_T / _SKIP_COLS / _SKIP_STORE, DGOD_OWN_TIMING.
#include <algorithm>
#include "roi_common.cuh"
#include "bulk.cuh"
#include "tmap.cuh"


namespace dgod {

namespace own {

constexpr int kP = 7;                  // pooled size (PH = PW = 7)
constexpr int kNB = kP * kP;
constexpr int kMaxSamp = 14;           // samples per axis (kP * sampling_ratio, sr <= 2)
constexpr int kMaxLive = 2 * kMaxSamp; // live rows / columns per axis
constexpr int kTileH = 28, kTileW = 16;
constexpr int kCS = 64;                // channels per slice: lane c owns c and c + 32
#define OWN_WAIT mbar_wait
#ifndef DGOD_OWN_RPW
#define DGOD_OWN_RPW 2
#endif
constexpr int kRPW = DGOD_OWN_RPW;     // tile rows per consumer warp: 1 -> 28 consumer warps of <= 64 registers, 2 -> 14 of <= 128
constexpr int kWarps = kTileH / kRPW;  // consumer warps; + the producer warp + the decoder warp
constexpr int kThreads = (kWarps + 2) * 32;
constexpr int kBlk = 4;                // the consumers sweep a tile row in blocks of 4 columns (absolute columns 4m .. 4m+3) ...
constexpr int kTaps = 4;               // ... whose weights sit on 4 consecutive bins (else the general 7-bin form runs)
constexpr int kMaxBlk = 8;             // blocks a span of <= 28 columns can meet
constexpr int kOwnTable = 0x40000000;  // PairInfo::w_off: the weights are in PairInfo::wtab, not in the plan
constexpr int kItemRing = 8;           // claims the producer may be ahead of the slowest warp of its CTA
constexpr int kCounterBytes = 73728;   // [0] pair cursor; +1024: per-CTA, per-warp cycle counters of DGOD_OWN_TIMING builds

struct alignas(128) Plan {
  short n_rows, n_cols;                // 16-byte header, read by the consumers as one int4
  short y_first, y_last;               // extent of the live rows
  short x_first, x_last;               // extent of the live columns
  short blocked;                       // 1: span mode (consecutive columns) and every block fits kTaps bins: wblk / blk are valid
  short pad0;
  short rows[kMaxLive];                // live feature rows, ascending (padding 0x7fff)
  short cols[kMaxLive];                // live feature columns, ascending (span mode: every column of the span)
  float ay[kMaxLive][8];               // A_y[row][ph] / count, list order
  float ax[kMaxLive][8];               // A_x[col][pw], list order (zero rows for columns without weight): the general form
  float4 wblk[kMaxBlk][kBlk];          // blocked form: weights of bins pb0 .. pb0+3 of absolute column 4*(x_first/4 + b) + c, zeros outside
  unsigned char blk[kMaxBlk];          // bits 0-1: pb0 of block b; bit 2: the block has a column with weight
  unsigned char aymask[kMaxLive];      // bit ph: A_y[row][ph] != 0
  unsigned char pad[92];
};
static_assert(sizeof(Plan) == 2560, "plans are moved with tensor-map loads");

struct alignas(16) Bin { int level, batch; short y0, y1, x0, x1; };   // level < 0: nothing to add
static_assert(sizeof(Bin) == 16, "read as one int4");

// What the decoder warp hands the consumer warps for one (RoI, tile) pair, so that the intersection of the RoI's live rows /
// columns with the tile is computed once per pair instead of once per consumer warp.
struct alignas(16) PairInfo {
  unsigned rowmask;                    // bit r: tile row r is live
  unsigned i_first;                    // list index of the first live row inside the tile
  unsigned blocks;                     // blocked form: byte B = block B of the tile (bits 0-1 base bin, bit 2 has weight); bit 31: general form
  int w_off;                           // blocked form: first wblk entry of the tile's block 0 (may be negative)
  unsigned colmask, j_first, pad0, pad1;   // general form: live tile columns, list index of the first one
  float4 wtab[kTileW];                 // blocked form of an RoI whose plan has none (w_off == kOwnTable): built per tile by the decoder
};
static_assert(sizeof(PairInfo) == 32 + 16 * kTileW, "16-byte reads");

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
own_plan_kernel(const RoiDev g, const float* __restrict__ rois, int n_rois, Plan* __restrict__ plans, Bin* __restrict__ bins,
                int* __restrict__ cursor) {
  __shared__ Scratch scratch[kPlanWarps];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (blockIdx.x == 0 && threadIdx.x == 0) { cursor[0] = 0; cursor[4] = 0; }   // pair cursor of the bin kernel, work counter of the main kernel
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
  // A_x, one lane per table row; in span mode also the blocked form: absolute 4-column blocks, each with a base bin pb0
  // such that every weight of its columns sits on bins pb0 .. pb0+3 (if one block needs more, the RoI keeps the general form)
  bool wide = false;
  if (n_rows) {
    float wv[8];
    const int col = span_mode ? x_first + lane : (int)t.list[1][lane < n_cols ? lane : 0];
    const bool in_list = lane < n_cols_out;
    unsigned nz = 0;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      wv[p] = (in_list && lane < kMaxLive && p < kP) ? axis_weight(t.lo[1], t.hi[1], t.l[1], t.h[1], col, p, sr) : 0.f;
      if (wv[p] != 0.f) nz |= 1u << p;
    }
    if (lane < kMaxLive) {
      *reinterpret_cast<float4*>(&P.ax[lane][0]) = make_float4(wv[0], wv[1], wv[2], wv[3]);
      *reinterpret_cast<float4*>(&P.ax[lane][4]) = make_float4(wv[4], wv[5], wv[6], wv[7]);
    }
    P.wblk[lane >> 2][lane & 3] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < kMaxBlk) P.blk[lane] = 0;
    __syncwarp();
    if (span_mode) {
      const int c = col & 3, bi = (col >> 2) - (x_first >> 2);
      int lo = nz ? __ffs(nz) - 1 : kP, hi = nz ? 31 - __clz(nz) : -1;
      int blo = kP, bhi = -1;
#pragma unroll
      for (int d = 0; d < kBlk; ++d) {                       // the (up to 4) lanes of this lane's block
        const int src = lane - c + d;
        const int olo = __shfl_sync(0xffffffffu, lo, src & 31), ohi = __shfl_sync(0xffffffffu, hi, src & 31);
        if (src >= 0 && src < n_cols_out) { blo = min(blo, olo); bhi = max(bhi, ohi); }
      }
      const int pb0 = min(blo, kP - kTaps);
      wide = in_list && bhi > pb0 + kTaps - 1;
      if (in_list) {
        float4 e;
        e.x = pb0 == 0 ? wv[0] : pb0 == 1 ? wv[1] : pb0 == 2 ? wv[2] : wv[3];
        e.y = pb0 == 0 ? wv[1] : pb0 == 1 ? wv[2] : pb0 == 2 ? wv[3] : wv[4];
        e.z = pb0 == 0 ? wv[2] : pb0 == 1 ? wv[3] : pb0 == 2 ? wv[4] : wv[5];
        e.w = pb0 == 0 ? wv[3] : pb0 == 1 ? wv[4] : pb0 == 2 ? wv[5] : wv[6];
        P.wblk[bi][c] = e;
        if (c == 0 || lane == 0) P.blk[bi] = (unsigned char)((pb0 & 3) | (bhi >= 0 ? 4 : 0));
      }
    }
  }
  const int blocked = span_mode && !__any_sync(0xffffffffu, wide);
  if (lane == 0) {
    P.n_rows = (short)n_rows; P.n_cols = (short)n_cols_out;
    P.y_first = n_rows ? t.list[0][0] : (short)0; P.y_last = n_rows ? t.list[0][n_rows - 1] : (short)-1;
    P.x_first = (short)x_first; P.x_last = (short)x_last;
    P.blocked = (short)blocked; P.pad0 = 0;
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
    const float inv = 1.f / r.count;           // count = sr*sr: a power of two, so scaling the table is exact
    unsigned m = 0;
    float wv[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      wv[p] = (lane < n_rows && p < kP) ? axis_weight(t.lo[0], t.hi[0], t.l[0], t.h[0], t.list[0][lane], p, sr) * inv : 0.f;
      if (wv[p] != 0.f) m |= 1u << p;
    }
    *reinterpret_cast<float4*>(&P.ay[lane][0]) = make_float4(wv[0], wv[1], wv[2], wv[3]);
    *reinterpret_cast<float4*>(&P.ay[lane][4]) = make_float4(wv[4], wv[5], wv[6], wv[7]);
    P.aymask[lane] = (unsigned char)m;
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
  // pass 1: count the hits, remembering those of the first 1024 RoIs of the image as one bit per chunk of 32 (the loads of
  // consecutive chunks are independent: unrolled, they are in flight together instead of one L2 round trip per chunk)
  int cnt = 0;
  unsigned bits = 0;
  {
    int c = 0;
#pragma unroll 8
    for (int base = k_begin; base < k_end; base += 32, ++c) {
      const bool hit = bin_hit(bins, base + lane, k_end, a);
      if (c < 32 && hit) bits |= 1u << c;
      cnt += __popc(__ballot_sync(0xffffffffu, hit));
    }
  }
  int off = 0;
  if (lane == 0) {
    off = cnt ? atomicAdd(cursor, cnt) : 0;
    tile_list[t] = make_int2(off, cnt);
  }
  if (!cnt) return;
  off = __shfl_sync(0xffffffffu, off, 0);
  {
    int c = 0;
#pragma unroll 4
    for (int base = k_begin; base < k_end; base += 32, ++c) {
      const bool hit = c < 32 ? ((bits >> c) & 1u) : bin_hit(bins, base + lane, k_end, a);
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) pair_k[off + __popc(m & ((1u << lane) - 1u))] = base + lane;
      off += __popc(m);
    }
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
  static constexpr int kPlanRowWords = 128, kPlanBoxRows = (int)sizeof(Plan) / (kPlanRowWords * 4);
  static_assert(kGBoxRows * kGRowWords * 4 == kGBytes && kPlanBoxRows * kPlanRowWords * 4 == (int)sizeof(Plan), "box shapes");
  static_assert(kStageBytes % 128 == 0 && kStages >= 4 && kStages <= 32, "stage ring");
};

template <typename T> __device__ __forceinline__ float lds_as_f32(const T* p);
template <> __device__ __forceinline__ float lds_as_f32<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float lds_as_f32<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __uint_as_float((unsigned)(*reinterpret_cast<const unsigned short*>(p)) << 16);
}

template <typename T> __device__ __forceinline__ float lds_g(unsigned a);        // gradient element at a 32-bit shared address
template <> __device__ __forceinline__ float lds_g<float>(unsigned a) { return lds_f32s(a); }
template <> __device__ __forceinline__ float lds_g<__nv_bfloat16>(unsigned a) { return lds_bf16s(a); }

// acc[r] += sum_pw w[pw] * t[r][pw]: all 7 bins (the general form)
__device__ __forceinline__ void column_dense(const float4 w0, const float4 w1, const float2 (&t)[kRPW][kP], float2 (&acc)[kRPW][kTileW],
                                             int x) {
  const float w[kP] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z};
#pragma unroll
  for (int pw = 0; pw < kP; ++pw)
#pragma unroll
    for (int r = 0; r < kRPW; ++r) acc[r][x] = __ffma2_rn(make_float2(w[pw], w[pw]), t[r][pw], acc[r][x]);
}

// One block of 4 tile columns whose weights sit on bins P0 .. P0+3 (the plan's wblk row): 4 broadcast reads, then
// 4 columns x kRPW rows x 4 multiply-adds with no branch and no dependence between columns.
template <int B, int P0>
__device__ __forceinline__ void column_block(unsigned wtab, const float2 (&t)[kRPW][kP], float2 (&acc)[kRPW][kTileW]) {
  float4 w[kBlk];
#pragma unroll
  for (int c = 0; c < kBlk; ++c) w[c] = lds_f4s(wtab + 16u * c);
#pragma unroll
  for (int c = 0; c < kBlk; ++c)
#pragma unroll
    for (int r = 0; r < kRPW; ++r) acc[r][B * kBlk + c] = __ffma2_rn(make_float2(w[c].x, w[c].x), t[r][P0], acc[r][B * kBlk + c]);
#pragma unroll
  for (int c = 0; c < kBlk; ++c)
#pragma unroll
    for (int r = 0; r < kRPW; ++r) acc[r][B * kBlk + c] = __ffma2_rn(make_float2(w[c].y, w[c].y), t[r][P0 + 1], acc[r][B * kBlk + c]);
#pragma unroll
  for (int c = 0; c < kBlk; ++c)
#pragma unroll
    for (int r = 0; r < kRPW; ++r) acc[r][B * kBlk + c] = __ffma2_rn(make_float2(w[c].z, w[c].z), t[r][P0 + 2], acc[r][B * kBlk + c]);
#pragma unroll
  for (int c = 0; c < kBlk; ++c)
#pragma unroll
    for (int r = 0; r < kRPW; ++r) acc[r][B * kBlk + c] = __ffma2_rn(make_float2(w[c].w, w[c].w), t[r][P0 + 3], acc[r][B * kBlk + c]);
}

// `blocks`: byte B describes the tile's block B (bits 0-1: base bin, bit 2: has weight); `w`: its 4 table entries
template <int B>
__device__ __forceinline__ void column_block_any(unsigned blocks, unsigned w, const float2 (&t)[kRPW][kP],
                                                 float2 (&acc)[kRPW][kTileW]) {
  if ((blocks >> (8 * B + 2)) & 1u) {
    const unsigned p0 = (blocks >> (8 * B)) & 3u;
    if (p0 == 0) column_block<B, 0>(w + 16u * (B * kBlk), t, acc);
    else if (p0 == 1) column_block<B, 1>(w + 16u * (B * kBlk), t, acc);
    else if (p0 == 2) column_block<B, 2>(w + 16u * (B * kBlk), t, acc);
    else column_block<B, 3>(w + 16u * (B * kBlk), t, acc);
  }
}

// kClaim: work items are claimed from a global counter (algo 5) instead of dealt round-robin (algo 4) — see the producer.
template <typename T, bool kClaim>
__global__ void __launch_bounds__(kThreads, 1)
own_bwd_kernel(const RoiDev g, const Tiles tg, const __grid_constant__ CUtensorMap tm_plan, const __grid_constant__ CUtensorMap tm_g,
               const int2* __restrict__ tile_list, const int* __restrict__ pair_k, int* __restrict__ work_counter,
               long long* __restrict__ timing) {
  constexpr int NS = Cfg<T>::kStages;
  constexpr int SB = Cfg<T>::kStageBytes;
  extern __shared__ __align__(128) unsigned char smem[];
  // item_*: the work items this CTA has claimed (item id, first pair, pairs), published by the producer for the other warps
  struct alignas(16) RingSync {
    PairInfo info[NS];
    int4 item[kItemRing];
    unsigned long long full[NS], ready[NS], empty[NS], item_full[kItemRing], item_empty[kItemRing];
  };
  __shared__ RingSync rs;
  PairInfo* const info = rs.info;
  unsigned long long* const full = rs.full;
  unsigned long long* const ready = rs.ready;
  unsigned long long* const empty = rs.empty;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_items = tg.n_tiles * tg.n_slices;

  const int C = g.C;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&ready[i], 1);
      mbar_init(&empty[i], kWarps);
    }
#pragma unroll
    for (int i = 0; i < kItemRing; ++i) {
      mbar_init(&rs.item_full[i], 1);
      mbar_init(&rs.item_empty[i], kWarps + 1);
    }
    mbar_fence_init();
    tmap_prefetch(&tm_plan);
    tmap_prefetch(&tm_g);
  }
  __syncthreads();
  // 32-bit shared-window addresses of the ring's bookkeeping, computed once and opaque to the compiler
  const unsigned rs_s = smem_opaque(&rs), smem_s = smem_opaque(smem);
  constexpr unsigned kFullOff = (unsigned)offsetof(RingSync, full), kReadyOff = (unsigned)offsetof(RingSync, ready),
                     kEmptyOff = (unsigned)offsetof(RingSync, empty), kItemOff = (unsigned)offsetof(RingSync, item),
                     kItemFullOff = (unsigned)offsetof(RingSync, item_full), kItemEmptyOff = (unsigned)offsetof(RingSync, item_empty);

  if (warp == kWarps) {
    // ------------------------------------------------------------------ producer.  Walks the pair lists of this CTA's
    // items (tile x channel slice, dealt round-robin: heavy levels first) with the whole warp — 32 RoI indices per coalesced
    // read, the next item's first chunk read early; lane 0 issues the two tensor-map loads of a stage in order, as far
    // ahead as the ring allows.
    const int rows_per_roi = C / kCS * Cfg<T>::kGBoxRows;
    unsigned s = 0, phase = 0, q = 0;
    if constexpr (kClaim) {
    // Work items are CLAIMED, not dealt: the first one is blockIdx.x, the following ones come from a global counter in index
    // order (coarse levels first).  A CTA that starts late — its SM was still busy with another stream's kernel, e.g. an NCCL
    // all-reduce overlapping this backward — or that drew heavier items simply takes fewer of them.  Three claims are in
    // flight: [2] its atomic is outstanding, [1] its pair-list header is loading, [0] ready to issue; every claim is
    // published in the item ring for the decoder and the consumers.
    auto draw = [&]() -> int { return lane == 0 ? atomicAdd(work_counter, 1) + (int)gridDim.x : 0; };
    auto settle = [&](int raw) -> int { const int v = __shfl_sync(0xffffffffu, raw, 0); return v < n_items ? v : -1; };
    auto header = [&](int it) -> int2 { return it >= 0 ? __ldg(tile_list + it / tg.n_slices) : make_int2(0, 0); };
    int it0 = (int)blockIdx.x < n_items ? (int)blockIdx.x : -1;
    int raw1 = draw();
    int2 lc0 = header(it0);
    int it1 = settle(raw1);
    int raw2 = it1 >= 0 ? draw() : n_items;
    int2 lc1 = header(it1);
    int kk0 = lane < lc0.y ? __ldg(pair_k + lc0.x + lane) : 0;
    unsigned n_pub = 0;
    while (true) {
      const int it2 = settle(raw2);
      raw2 = it2 >= 0 ? draw() : n_items;
      const int2 lc2 = header(it2);
      const int kk1 = lane < lc1.y ? __ldg(pair_k + lc1.x + lane) : 0;
      {
        const unsigned j = n_pub % kItemRing;
        if (lane == 0) {
          if (n_pub >= (unsigned)kItemRing) mbar_wait_s(rs_s + kItemEmptyOff + j * 8u, ((n_pub / kItemRing) & 1u) ^ 1u);
          sts_u4s(rs_s + kItemOff + j * 16u, make_uint4((unsigned)it0, (unsigned)lc0.x, (unsigned)lc0.y, 0u));
          mbar_arrive_s(rs_s + kItemFullOff + j * 8u);          // release: the entry is visible to whoever observes the phase
        }
        ++n_pub;
      }
      if (it0 < 0) break;
      const int sl = it0 % tg.n_slices;
      int kk = kk0;
      for (int p0 = 0; p0 < lc0.y; p0 += 32) {
        if (p0) kk = p0 + lane < lc0.y ? __ldg(pair_k + lc0.x + p0 + lane) : 0;
        const int n = min(32, lc0.y - p0);
        for (int i = 0; i < n; ++i, ++q) {
          const int k = __shfl_sync(0xffffffffu, kk, i);
          if (lane == 0) {
            unsigned char* st = smem + (size_t)s * SB;
            if (q >= (unsigned)NS) mbar_wait_s(rs_s + kEmptyOff + s * 8u, phase ^ 1u);
            mbar_expect_tx_s(rs_s + kFullOff + s * 8u, (unsigned)SB);
            tmap_load_2d(st, &tm_plan, 0, k * Cfg<T>::kPlanBoxRows, &full[s]);
            tmap_load_2d(st + sizeof(Plan), &tm_g, 0, k * rows_per_roi + sl * Cfg<T>::kGBoxRows, &full[s]);
          }
          if (++s == (unsigned)NS) { s = 0; phase ^= 1u; }
        }
        __syncwarp();
      }
      it0 = it1; lc0 = lc1; kk0 = kk1;
      it1 = it2; lc1 = lc2;
    }
    } else {
    int n_claimed = 0;
    auto claim = [&]() -> int {
      const int it = (int)blockIdx.x + n_claimed * (int)gridDim.x;
      ++n_claimed;
      return it < n_items ? it : -1;
    };
    int it = claim();
    int2 lc = make_int2(0, 0);
    int kk = 0;
    if (it >= 0) {
      lc = __ldg(tile_list + it / tg.n_slices);
      kk = lane < lc.y ? __ldg(pair_k + lc.x + lane) : 0;
    }
    while (it >= 0) {
      const int sl = it % tg.n_slices;
      const int it_next = claim();
      int2 lc_next = make_int2(0, 0);
      int kk_next = 0;
      if (it_next >= 0) {
        lc_next = __ldg(tile_list + it_next / tg.n_slices);
        kk_next = lane < lc_next.y ? __ldg(pair_k + lc_next.x + lane) : 0;
      }
      for (int p0 = 0; p0 < lc.y; p0 += 32) {
        if (p0) kk = p0 + lane < lc.y ? __ldg(pair_k + lc.x + p0 + lane) : 0;
        const int n = min(32, lc.y - p0);
        for (int i = 0; i < n; ++i, ++q) {
          const int k = __shfl_sync(0xffffffffu, kk, i);
          if (lane == 0) {
            unsigned char* st = smem + (size_t)s * SB;
            if (q >= (unsigned)NS) mbar_wait_s(rs_s + kEmptyOff + s * 8u, phase ^ 1u);
            mbar_expect_tx_s(rs_s + kFullOff + s * 8u, (unsigned)SB);
            tmap_load_2d(st, &tm_plan, 0, k * Cfg<T>::kPlanBoxRows, &full[s]);
            tmap_load_2d(st + sizeof(Plan), &tm_g, 0, k * rows_per_roi + sl * Cfg<T>::kGBoxRows, &full[s]);
          }
          if (++s == (unsigned)NS) { s = 0; phase ^= 1u; }
        }
        __syncwarp();
      }
      it = it_next; lc = lc_next; kk = kk_next;
    }
    }
    return;
  }

  if (warp == kWarps + 1) {
    // ------------------------------------------------------------------ decoder: once per pair, intersects the RoI's live
    // rows / columns with the tile and publishes the result in info[stage]
    unsigned s = 0, phase = 0;
    int2 lc_next = make_int2(0, 0);
    if constexpr (!kClaim) lc_next = (int)blockIdx.x < n_items ? __ldg(tile_list + blockIdx.x / tg.n_slices) : make_int2(0, 0);
    for (unsigned n_it = 0;; ++n_it) {
      int it;
      int2 lc;
      if constexpr (kClaim) {
        const unsigned jr = n_it % kItemRing;
        mbar_wait_s(rs_s + kItemFullOff + jr * 8u, (n_it / kItemRing) & 1u);
        const uint4 e = lds_u4(rs_s + kItemOff + jr * 16u);
        __syncwarp();
        if (lane == 0) mbar_arrive_s(rs_s + kItemEmptyOff + jr * 8u);
        if ((int)e.x < 0) break;
        it = (int)e.x;
        lc = make_int2((int)e.y, (int)e.z);
      } else {
        it = (int)blockIdx.x + (int)n_it * (int)gridDim.x;
        if (it >= n_items) break;
        lc = lc_next;
        if (it + (int)gridDim.x < n_items) lc_next = __ldg(tile_list + (it + gridDim.x) / tg.n_slices);   // next item's list, early
      }
      const TileAt a = tile_at(tg, it / tg.n_slices);
      for (int p = 0; p < lc.y; ++p) {
        mbar_wait_s(rs_s + kFullOff + s * 8u, phase);
        const unsigned st_s = smem_s + s * (unsigned)SB;
        const int rel_r = lane < kMaxLive ? lds_s16s(st_s + (unsigned)offsetof(Plan, rows) + 2u * lane) - a.y0 : 0x7fff;   // padding entries are 0x7fff
        const unsigned rowmask = __reduce_or_sync(0xffffffffu, (rel_r >= 0 && rel_r < kTileH) ? (1u << rel_r) : 0u);
        const unsigned i_first = __popc(__ballot_sync(0xffffffffu, rel_r < 0));
        const uint4 hdr_u = lds_u4(st_s);
        const int4 hdr = make_int4((int)hdr_u.x, (int)hdr_u.y, (int)hdr_u.z, (int)hdr_u.w);
        const int x_first = (short)(hdr.z & 0xffff);
        unsigned blocks = 0x80000000u, colmask = 0, j_first = 0;
        int w_off = 0;
        if (hdr.w & 0xffff) {
          const int b_first = (a.x0 >> 2) - (x_first >> 2);                       // plan block of the tile's block 0 (-3 .. 7)
          const unsigned long long all = lds_u64s(st_s + (unsigned)offsetof(Plan, blk));
          blocks = (b_first >= 0 ? (unsigned)(all >> (8 * b_first)) : (unsigned)(all << (8 * -b_first))) & 0x07070707u;
          w_off = b_first * kBlk;
          colmask = blocks;                                                         // only "any live column" matters below
        } else {
          // the plan has no blocked form (columns far apart, or a block of the span needs more than 4 bins): try again
          // for the 4 blocks of THIS tile from the A_x rows of its live columns; the general form is the last resort
          const int rel_c = lane < kMaxLive ? lds_s16s(st_s + (unsigned)offsetof(Plan, cols) + 2u * lane) - a.x0 : 0x7fff;
          colmask = __reduce_or_sync(0xffffffffu, (rel_c >= 0 && rel_c < kTileW) ? (1u << rel_c) : 0u);
          j_first = __popc(__ballot_sync(0xffffffffu, rel_c < 0));
          float row[8];
          {
            const bool live = lane < kTileW && ((colmask >> lane) & 1u);
            const int j = (int)j_first + __popc(colmask & ((1u << lane) - 1u));
            const unsigned pax = st_s + (unsigned)offsetof(Plan, ax) + (unsigned)j * 32u;
            const float4 a0 = live ? lds_f4s(pax) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 a1 = live ? lds_f4s(pax + 16u) : make_float4(0.f, 0.f, 0.f, 0.f);
            row[0] = a0.x; row[1] = a0.y; row[2] = a0.z; row[3] = a0.w; row[4] = a1.x; row[5] = a1.y; row[6] = a1.z; row[7] = 0.f;
          }
          unsigned nz = 0;
#pragma unroll
          for (int pw = 0; pw < kP; ++pw) nz |= row[pw] != 0.f ? 1u << pw : 0u;
          int lo = nz ? __ffs(nz) - 1 : kP, hi = nz ? 31 - __clz(nz) : -1;
          lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, 1)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, 1));
          lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, 2)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, 2));
          const int pb0 = min(lo, kP - kTaps);
          const unsigned wide = __ballot_sync(0xffffffffu, lane < kTileW && hi > pb0 + kTaps - 1);
          if (!wide) {
            float4 wv;
            wv.x = pb0 == 0 ? row[0] : pb0 == 1 ? row[1] : pb0 == 2 ? row[2] : row[3];
            wv.y = pb0 == 0 ? row[1] : pb0 == 1 ? row[2] : pb0 == 2 ? row[3] : row[4];
            wv.z = pb0 == 0 ? row[2] : pb0 == 1 ? row[3] : pb0 == 2 ? row[4] : row[5];
            wv.w = pb0 == 0 ? row[3] : pb0 == 1 ? row[4] : pb0 == 2 ? row[5] : row[6];
            if (lane < kTileW) sts_f4s(rs_s + s * (unsigned)sizeof(PairInfo) + (unsigned)offsetof(PairInfo, wtab) + 16u * lane, wv);
            blocks = 0;
#pragma unroll
            for (int b = 0; b < kTileW / kBlk; ++b) {
              const int pb = __shfl_sync(0xffffffffu, pb0, b * kBlk), bh = __shfl_sync(0xffffffffu, hi, b * kBlk);
              blocks |= ((unsigned)(pb & 3) | (bh >= 0 ? 4u : 0u)) << (8 * b);
            }
            w_off = kOwnTable;
            __syncwarp();
          }
        }
        if (lane == 0) {
          const unsigned is = rs_s + s * (unsigned)sizeof(PairInfo);
          sts_u4s(is, make_uint4(colmask ? rowmask : 0u, i_first, blocks, (unsigned)w_off));   // no live column in this tile: nothing to do for any warp
          sts_u4s(is + 16u, make_uint4(colmask, j_first, 0u, 0u));
          mbar_arrive_s(rs_s + kReadyOff + s * 8u);    // release: the stores above are visible to whoever observes the phase
        }
        __syncwarp();
        if (++s == (unsigned)NS) { s = 0; phase ^= 1u; }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers: tile rows kRPW*warp ..
  const int r0 = kRPW * warp;
  unsigned q = 0;
#ifdef DGOD_OWN_TIMING
  const long long t_start = clock64();
  long long t_wait = 0, t_store = 0;
#endif
  unsigned s = 0, phase = 0;                                // ring position of pair q
  int2 lc_next = make_int2(0, 0);
  if constexpr (!kClaim) lc_next = (int)blockIdx.x < n_items ? __ldg(tile_list + blockIdx.x / tg.n_slices) : make_int2(0, 0);
  for (unsigned n_it = 0;; ++n_it) {
    int it;
    int2 lc;
    if constexpr (kClaim) {
      const unsigned jr = n_it % kItemRing;
      mbar_wait_s(rs_s + kItemFullOff + jr * 8u, (n_it / kItemRing) & 1u);
      const uint4 e = lds_u4(rs_s + kItemOff + jr * 16u);
      __syncwarp();
      if (lane == 0) mbar_arrive_s(rs_s + kItemEmptyOff + jr * 8u);
      if ((int)e.x < 0) break;
      it = (int)e.x;
      lc = make_int2((int)e.y, (int)e.z);
    } else {
      it = (int)blockIdx.x + (int)n_it * (int)gridDim.x;
      if (it >= n_items) break;
      lc = lc_next;
      if (it + (int)gridDim.x < n_items) lc_next = __ldg(tile_list + (it + gridDim.x) / tg.n_slices);   // next item's list, early
    }
    const int t_id = it / tg.n_slices, sl = it - t_id * tg.n_slices;
    const TileAt a = tile_at(tg, t_id);
    float2 acc[kRPW][kTileW];
#pragma unroll
    for (int r = 0; r < kRPW; ++r)
#pragma unroll
      for (int x = 0; x < kTileW; ++x) acc[r][x] = make_float2(0.f, 0.f);

    for (int p = 0; p < lc.y; ++p, ++q) {
#ifdef DGOD_OWN_TIMING
      const long long tw = clock64();
#endif
      mbar_wait_s(rs_s + kReadyOff + s * 8u, phase);
#ifdef DGOD_OWN_TIMING
      t_wait += clock64() - tw;
#endif
      const uint4 ia = lds_u4(rs_s + s * (unsigned)sizeof(PairInfo));             // rowmask, i_first, blocks, w_off
      const unsigned rm = (ia.x >> r0) & ((1u << kRPW) - 1u);
      {
        if (rm) {
        const unsigned st_s = smem_s + s * (unsigned)SB;                       // this stage: plan, then the gradient slice
        int idx[kRPW];
        {
          const int i0 = (int)ia.y + __popc(ia.x & ((1u << r0) - 1u));          // list index of this warp's first live row
#pragma unroll
          for (int r = 0; r < kRPW; ++r) idx[r] = ((rm >> r) & 1u) ? i0 + __popc(rm & ((1u << r) - 1u)) : i0;
        }
        float ay[kRPW][8];
        unsigned phm = 0;
#pragma unroll
        for (int r = 0; r < kRPW; ++r) {
          const bool live = (rm >> r) & 1u;
          const float z = live ? 1.f : 0.f;                                     // a dead row reads row 0 of the table, times zero
          const unsigned pa = st_s + (unsigned)offsetof(Plan, ay) + (unsigned)idx[r] * 32u;
          const float4 u0 = lds_f4s(pa), u1 = lds_f4s(pa + 16u);
          ay[r][0] = u0.x * z; ay[r][1] = u0.y * z; ay[r][2] = u0.z * z; ay[r][3] = u0.w * z;
          ay[r][4] = u1.x * z; ay[r][5] = u1.y * z; ay[r][6] = u1.z * z;
          phm |= live ? lds_u8s(st_s + (unsigned)offsetof(Plan, aymask) + (unsigned)idx[r]) : 0u;
        }
        float2 t[kRPW][kP];
#pragma unroll
        for (int r = 0; r < kRPW; ++r)
#pragma unroll
          for (int pw = 0; pw < kP; ++pw) t[r][pw] = make_float2(0.f, 0.f);
        const unsigned ga = st_s + (unsigned)sizeof(Plan) + (unsigned)(lane * kNB) * (unsigned)sizeof(T);
        const unsigned gb = ga + 32u * kNB * (unsigned)sizeof(T);
#ifndef DGOD_OWN_SKIP_T
#pragma unroll
        for (int ph = 0; ph < kP; ++ph) {
          if ((phm >> ph) & 1u) {                           // warp-uniform
#pragma unroll
            for (int pw = 0; pw < kP; ++pw) {
              const float2 gv = make_float2(lds_g<T>(ga + (unsigned)((ph * kP + pw) * sizeof(T))), lds_g<T>(gb + (unsigned)((ph * kP + pw) * sizeof(T))));
#pragma unroll
              for (int r = 0; r < kRPW; ++r) t[r][pw] = __ffma2_rn(make_float2(ay[r][ph], ay[r][ph]), gv, t[r][pw]);
            }
          }
        }
#else
#pragma unroll
        for (int r = 0; r < kRPW; ++r)
#pragma unroll
          for (int pw = 0; pw < kP; ++pw) t[r][pw] = make_float2(ay[r][pw] + (float)phm, lds_g<T>(ga + (unsigned)(pw * sizeof(T))));
#endif
#ifdef DGOD_OWN_SKIP_COLS
#pragma unroll
        for (int r = 0; r < kRPW; ++r)
#pragma unroll
          for (int pw = 0; pw < kP; ++pw) acc[r][pw] = __fadd2_rn(acc[r][pw], t[r][pw]);
#else
        if (!(ia.z >> 31)) {
          // blocked form: the 4 absolute column blocks of this tile, branch-free inside a block
          const unsigned w = (int)ia.w == kOwnTable ? rs_s + s * (unsigned)sizeof(PairInfo) + (unsigned)offsetof(PairInfo, wtab)
                                                    : st_s + (unsigned)offsetof(Plan, wblk) + (unsigned)((int)ia.w * 16);   // absent blocks are never read
          column_block_any<0>(ia.z, w, t, acc);
          column_block_any<1>(ia.z, w, t, acc);
          column_block_any<2>(ia.z, w, t, acc);
          column_block_any<3>(ia.z, w, t, acc);
        } else {
          // general form (columns far apart, or more than 4 bins on a block): all 7 bins per live column
          const uint4 ib = lds_u4(rs_s + s * (unsigned)sizeof(PairInfo) + 16u);   // colmask, j_first
          const unsigned cm = ib.x;
          const int j_first = (int)ib.y;
#pragma unroll
          for (int x = 0; x < kTileW; ++x) {
            if ((cm >> x) & 1u) {                           // warp-uniform
              const int j = j_first + __popc(cm & ((1u << x) - 1u));
              const unsigned px = st_s + (unsigned)offsetof(Plan, ax) + (unsigned)j * 32u;
              column_dense(lds_f4s(px), lds_f4s(px + 16u), t, acc, x);
            }
          }
        }
#endif
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_s(rs_s + kEmptyOff + s * 8u);
      if (++s == (unsigned)NS) { s = 0; phase ^= 1u; }
    }

    // ---- the tile leaves once, zeros included
#ifdef DGOD_OWN_TIMING
    const long long ts = clock64();
#endif
    const int H = g.H[a.level], W = g.W[a.level];
    T* __restrict__ img = reinterpret_cast<T*>(g.gfeat[a.level]) + (size_t)a.b * H * W * C + (size_t)sl * kCS + lane;
#pragma unroll
    for (int r = 0; r < kRPW; ++r) {
      const int y = a.y0 + r0 + r;
#ifdef DGOD_OWN_SKIP_STORE
      if (y < H && acc[r][3].x == 123.456f) {
#else
      if (y < H) {
#endif
        T* __restrict__ rowp = img + ((size_t)y * W + a.x0) * C;
        if (a.x0 + kTileW <= W) {
#pragma unroll
          for (int x = 0; x < kTileW; ++x) {
            rowp[(size_t)x * C] = from_f32<T>(acc[r][x].x);
            rowp[(size_t)x * C + 32] = from_f32<T>(acc[r][x].y);
          }
        } else {
#pragma unroll
          for (int x = 0; x < kTileW; ++x) {
            if (a.x0 + x < W) {
              rowp[(size_t)x * C] = from_f32<T>(acc[r][x].x);
              rowp[(size_t)x * C + 32] = from_f32<T>(acc[r][x].y);
            }
          }
        }
      }
    }
#ifdef DGOD_OWN_TIMING
    t_store += clock64() - ts;
#endif
  }
#ifdef DGOD_OWN_TIMING
  if (lane == 0) {      // per CTA and consumer warp: total, waiting for a stage, storing, pairs
    long long* o = timing + ((size_t)blockIdx.x * kWarps + warp) * 4;
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

template <typename T, bool kClaim>
static int launch_own(const RoiDev& g, const own::Tiles& tg, const OwnLayout& L, const void* grad_out, const float* rois,
                      int n_rois, const int32_t* roi_img_offsets, char* ws, cudaStream_t st) {
  static bool init = false;
  static int n_sm = 0;
  if (!init) {
    int dev = 0;
    DGOD_CUDA(cudaGetDevice(&dev));
    DGOD_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    DGOD_CUDA(cudaFuncSetAttribute(own::own_bwd_kernel<T, kClaim>, cudaFuncAttributeMaxDynamicSharedMemorySize, own::Cfg<T>::kSmem));
    init = true;
  }
  int* cursor = reinterpret_cast<int*>(ws);
  own::Bin* bins = reinterpret_cast<own::Bin*>(ws + L.bins);
  int2* tile_list = reinterpret_cast<int2*>(ws + L.tile_list);
  int* pair_k = reinterpret_cast<int*>(ws + L.pair_k);
  own::Plan* plans = reinterpret_cast<own::Plan*>(ws + L.plans);
  if (n_rois > 0) {
    own::own_plan_kernel<<<cdiv(n_rois, own::kPlanWarps), own::kPlanWarps * 32, 0, st>>>(g, rois, n_rois, plans, bins, cursor);
    DGOD_LAUNCHED();
  } else {
    DGOD_CUDA(cudaMemsetAsync(cursor, 0, 8 * sizeof(int), st));
  }
  own::own_bin_kernel<<<cdiv(tg.n_tiles, own::kBinWarps), own::kBinWarps * 32, 0, st>>>(tg, bins, n_rois, roi_img_offsets, tile_list,
                                                                                      pair_k, cursor);
  DGOD_LAUNCHED();
  const int n_items = tg.n_tiles * tg.n_slices;
  const int grid = std::min(n_items, n_sm);
  CUtensorMap tm_plan, tm_g;
  memset(&tm_plan, 0, sizeof(tm_plan));
  memset(&tm_g, 0, sizeof(tm_g));
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
  own::own_bwd_kernel<T, kClaim><<<grid, own::kThreads, own::Cfg<T>::kSmem, st>>>(g, tg, tm_plan, tm_g, tile_list, pair_k, cursor + 4,
                                                                       reinterpret_cast<long long*>(ws + 1024));
  DGOD_LAUNCHED();
  return DGOD_OK;
}

int msroi_bwd_own(const dgod_roi_config* cfg, const RoiDev& g, const void* grad_out, const float* rois, int n_rois,
                  const int32_t* roi_img_offsets, void* workspace, size_t workspace_bytes, cudaStream_t st, int* handled, int claim) {
  *handled = 0;
  if (!own_shape_ok(g) || ((uintptr_t)grad_out & 15)) return DGOD_OK;
  if (!workspace || ((uintptr_t)workspace & 127)) return DGOD_OK;
  const own::Tiles tg = own_tiles(g);
  if (tg.n_tiles <= 0) return DGOD_OK;
  const OwnLayout L = own_layout(tg, n_rois);
  if (workspace_bytes < L.total) return DGOD_OK;
  *handled = 1;
  char* ws = (char*)workspace;
  if (cfg->dtype == DGOD_F32)
    return claim ? launch_own<float, true>(g, tg, L, grad_out, rois, n_rois, roi_img_offsets, ws, st)
                 : launch_own<float, false>(g, tg, L, grad_out, rois, n_rois, roi_img_offsets, ws, st);
  return claim ? launch_own<__nv_bfloat16, true>(g, tg, L, grad_out, rois, n_rois, roi_img_offsets, ws, st)
               : launch_own<__nv_bfloat16, false>(g, tg, L, grad_out, rois, n_rois, roi_img_offsets, ws, st);
}

}  // namespace dgod
