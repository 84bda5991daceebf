// MultiScaleRoIAlign forward / backward for channels_last features, built on the TMA engine
// (fasterrcnn.py:278,412-416 -> TV ops/poolers.py:147-227 ->
// torchvision::roi_align / _roi_align_backward).
//
// In NHWC a row of an RoI's footprint — pixels x0..x1 of feature row y, all C channels — is ONE
// contiguous span of (x1-x0+1)*C*s bytes.  Both kernels move whole footprint rows with the bulk
// async-copy engine instead of issuing per-tap loads / per-tap reductions from the SM lanes:
//
//   forward   cp.async.bulk global->shared (mbarrier complete_tx) streams the footprint rows
//             through a 2-stage ring; the pooled [C][PH*PW] block leaves as one bulk store.
//   backward  each footprint row is assembled in shared memory and added to the gradient map by
//             ONE cp.reduce.async.bulk (UBLKRED) — the L2 does the read-modify-write, the SM
//             issues one instruction per row instead of 16 RED per output element (the previous
//             kernel was bound by the SM's RED issue rate, ~1.3 cycles per lane).
//
// Arithmetic: bilinear pooling is separable.  With A_y[y][ph] = sum over the sampling rows of bin
// ph of the weight they put on feature row y (and A_x likewise),
//       out[c][ph][pw]  = sum_y A_y[y][ph] * ( sum_x A_x[x][pw] * f[y][x][c] ) / count
//       grad_f[y][x][c] = sum_pw A_x[x][pw] * ( sum_ph A_y[y][ph] * g[c][ph][pw] ) / count
// One thread owns one channel and keeps the PH*PW (= 49) pooled values / gradients of that channel
// in registers; the A tables are built once per RoI in shared memory and read as warp broadcasts.
// Out-of-range samples (TV's skip rule) and the border clamp are encoded in the tables by the same
// axis_tap() the exact kernels use.  The summation order differs from the CPU kernel's
// sample-by-sample order, i.e. results agree to fp32 rounding (tests: 1e-5 relative), not bitwise.
//
// Roofline: HBM.  Forward reads every touched feature line once (the image's maps stay in the
// 126 MB L2 while its RoIs are in flight) and writes K*C*49*s; backward reads K*C*49*s and writes
// every gradient line once: a cooperative persistent grid zero-fills image b, grid-syncs, then
// reduces image b's RoIs, so the read-modify-write traffic stays in L2 and DRAM sees one write.
#include "roi_common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace dgod {

constexpr int kTmaThreadsMax = 256;   // one thread per channel
constexpr int kFpCols = 32;           // footprint columns handled per pass (chunked beyond)
constexpr int kFpRows = 32;           // footprint rows per table build (chunked beyond)
constexpr int kP = 7;                 // pooled size handled by these kernels (PH = PW = 7)
constexpr int kPP = 8;                // table pitch (floats)
constexpr int kMaxSamp = 16;          // samples per axis (PH * sampling_ratio)

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
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }

// ---------------------------------------------------------------- per-RoI tables
struct AxisSamples {       // the PH*sr sampling coordinates of one axis, reduced to taps
  short lo[kMaxSamp], hi[kMaxSamp];
  float l[kMaxSamp], h[kMaxSamp];   // both zero when the sample is skipped (out of range)
  int first, last;                  // footprint extent over the valid samples (first > last: empty)
};

struct RoiTables {
  RoiGeom geo;
  AxisSamples sy, sx;
  float ay[kFpRows][kPP];   // A_y for rows  y0c .. y0c + kFpRows - 1 of the current chunk
  float ax[kFpCols][kPP];   // A_x for columns of the current chunk
  unsigned row_live;        // bit r: row r of the chunk receives any weight
};

// Threads 0..n-1 of a warp fill one axis; extent is reduced with shuffles.
__device__ __forceinline__ void fill_samples(AxisSamples& s, int lane, int n, int sr, float start, float bin, int size) {
  int lo = 0x7fffffff, hi = -1;
  if (lane < n) {
    const AxisTap a = axis_tap(sample_coord(start, lane / sr, bin, lane % sr, sr), size);
    s.lo[lane] = (short)a.lo;
    s.hi[lane] = (short)a.hi;
    s.l[lane] = a.valid ? a.l : 0.f;
    s.h[lane] = a.valid ? a.h : 0.f;
    if (a.valid) { lo = a.lo; hi = a.hi; }
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
  }
  if (lane == 0) { s.first = lo; s.last = hi; }
}

// A[r][p] for footprint coordinate base + r: thread (r, p), r < nrc, p < kPP.
__device__ __forceinline__ float axis_weight(const AxisSamples& s, int coord, int p, int sr) {
  float w = 0.f;
  if (p < kP) {
    for (int i = 0; i < sr; ++i) {
      const int q = p * sr + i;
      if (s.lo[q] == coord) w += s.h[q];
      if (s.hi[q] == coord) w += s.l[q];   // lo == hi at the clamped border: both weights land here
    }
  }
  return w;
}

// ================================================================================================
// Forward
// ================================================================================================
template <typename T>
__global__ void __launch_bounds__(kTmaThreadsMax, 2)
msroi_fwd_tma_kernel(const RoiDev g, const float* __restrict__ rois, int n_rois, T* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ RoiTables tb;
  __shared__ __align__(8) unsigned long long bar[2];
  const int C = g.C, tid = threadIdx.x, sr = g.sr;
  const int k = blockIdx.x;
  const size_t stage_bytes = (size_t)kFpCols * C * sizeof(T);
  T* ring[2] = {reinterpret_cast<T*>(smem_raw), reinterpret_cast<T*>(smem_raw + stage_bytes)};
  T* s_out = reinterpret_cast<T*>(smem_raw);   // [C][49], aliases the ring after the row loop

  if (tid == 0) {
    tb.geo = roi_geometry(g, rois + (size_t)k * 5);
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const RoiGeom r = tb.geo;
  const bool usable = r.batch >= 0 && r.batch < g.B;
  if (usable) {
    if (tid < 32) fill_samples(tb.sy, tid, kP * sr, sr, r.start_h, r.bin_h, r.H);
    else if (tid < 64) fill_samples(tb.sx, tid - 32, kP * sr, sr, r.start_w, r.bin_w, r.W);
  }
  __syncthreads();

  float acc[kP * kP];
#pragma unroll
  for (int i = 0; i < kP * kP; ++i) acc[i] = 0.f;

  if (usable && tb.sy.first <= tb.sy.last && tb.sx.first <= tb.sx.last) {
    const int y_first = tb.sy.first, y_last = tb.sy.last, x_first = tb.sx.first, x_last = tb.sx.last;
    const T* __restrict__ img = reinterpret_cast<const T*>(g.feat[r.level]) + (size_t)r.batch * r.H * r.W * C;
    unsigned phase[2] = {0u, 0u};
    for (int yc = y_first; yc <= y_last; yc += kFpRows) {
      const int nrows = min(kFpRows, y_last - yc + 1);
      for (int xc = x_first; xc <= x_last; xc += kFpCols) {
        const int ncols = min(kFpCols, x_last - xc + 1);
        __syncthreads();                      // previous chunk done with the tables and the ring
        for (int e = tid; e < kFpRows * kPP; e += blockDim.x) {   // 32 rows x 8 table columns
          const int er = e >> 3, ep = e & 7;
          tb.ay[er][ep] = er < nrows ? axis_weight(tb.sy, yc + er, ep, sr) : 0.f;
          tb.ax[er][ep] = er < ncols ? axis_weight(tb.sx, xc + er, ep, sr) : 0.f;
        }
        __syncthreads();
        if (tid < 32) {
          bool live = false;
          if (tid < nrows) {
#pragma unroll
            for (int p = 0; p < kP; ++p) live |= tb.ay[tid][p] != 0.f;
          }
          const unsigned m = __ballot_sync(0xffffffffu, live);
          if (tid == 0) tb.row_live = m;
        }
        __syncthreads();
        unsigned live = tb.row_live;
        const unsigned row_bytes = (unsigned)ncols * C * sizeof(T);
        // software pipeline over the live rows: row i+1 streams in while row i is consumed
        int cur = live ? __ffs(live) - 1 : -1;
        int stage = 0;
        if (cur >= 0 && tid == 0) {
          mbar_expect_tx(&bar[0], row_bytes);
          bulk_load(ring[0], img + ((size_t)(yc + cur) * r.W + xc) * C, row_bytes, &bar[0]);
        }
        while (cur >= 0) {
          live &= live - 1;
          const int nxt = live ? __ffs(live) - 1 : -1;
          if (nxt >= 0 && tid == 0) {
            mbar_expect_tx(&bar[stage ^ 1], row_bytes);
            bulk_load(ring[stage ^ 1], img + ((size_t)(yc + nxt) * r.W + xc) * C, row_bytes, &bar[stage ^ 1]);
          }
          mbar_wait(&bar[stage], phase[stage]);
          phase[stage] ^= 1u;
          const T* __restrict__ row = ring[stage] + tid;
          float rx[kP];
#pragma unroll
          for (int p = 0; p < kP; ++p) rx[p] = 0.f;
#pragma unroll 4
          for (int x = 0; x < ncols; ++x) {
            const float v = to_f32<T>(row[(size_t)x * C]);
            const float4 w0 = *reinterpret_cast<const float4*>(&tb.ax[x][0]);
            const float4 w1 = *reinterpret_cast<const float4*>(&tb.ax[x][4]);
            rx[0] = fmaf(w0.x, v, rx[0]); rx[1] = fmaf(w0.y, v, rx[1]); rx[2] = fmaf(w0.z, v, rx[2]);
            rx[3] = fmaf(w0.w, v, rx[3]); rx[4] = fmaf(w1.x, v, rx[4]); rx[5] = fmaf(w1.y, v, rx[5]);
            rx[6] = fmaf(w1.z, v, rx[6]);
          }
          const float4 a0 = *reinterpret_cast<const float4*>(&tb.ay[cur][0]);
          const float4 a1 = *reinterpret_cast<const float4*>(&tb.ay[cur][4]);
          const float a[kP] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z};
#pragma unroll
          for (int ph = 0; ph < kP; ++ph)
#pragma unroll
            for (int pw = 0; pw < kP; ++pw) acc[ph * kP + pw] = fmaf(a[ph], rx[pw], acc[ph * kP + pw]);
          __syncthreads();                    // everyone is done with ring[stage]: it may be refilled
          cur = nxt;
          stage ^= 1;
        }
      }
    }
  }
  __syncthreads();
  // pooled block -> shared [C][49] (lane stride 49 words: conflict-free) -> one bulk store
  const float inv = 1.f / r.count;           // count = sr*sr, a power of two for sr in {1, 2, 4}
  T* so = s_out + (size_t)tid * (kP * kP);
#pragma unroll
  for (int i = 0; i < kP * kP; ++i) so[i] = from_f32<T>(acc[i] * inv);
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    bulk_store(out + (size_t)k * C * (kP * kP), s_out, (unsigned)(C * kP * kP * sizeof(T)));
    bulk_commit();
    bulk_wait_read_all();
  }
}

// ================================================================================================
// Backward
// ================================================================================================
// Work of one RoI by one CTA (blockDim.x == C).  smem_raw: gradient block staging [C][49] which is
// then reused as the two row buffers.
template <typename T>
__device__ __forceinline__ void bwd_one_roi(const RoiDev& g, RoiTables& tb, unsigned char* smem_raw, const T* __restrict__ grad_out,
                                            const float* __restrict__ rois, int k) {
  const int C = g.C, tid = threadIdx.x, sr = g.sr;
  __syncthreads();                            // previous RoI of this CTA is completely done with smem
  if (tid == 0) tb.geo = roi_geometry(g, rois + (size_t)k * 5);
  __syncthreads();
  const RoiGeom r = tb.geo;
  if (r.batch < 0 || r.batch >= g.B) return;
  {
    // gradient block of this RoI, contiguous [C][49]: coalesced 16-byte copies into shared memory
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(grad_out + (size_t)k * C * (kP * kP));
    uint4* dst = reinterpret_cast<uint4*>(smem_raw);
    const int n16 = C * kP * kP * (int)sizeof(T) / 16;
    for (int i = tid; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  if (tid < 32) fill_samples(tb.sy, tid, kP * sr, sr, r.start_h, r.bin_h, r.H);
  else if (tid < 64) fill_samples(tb.sx, tid - 32, kP * sr, sr, r.start_w, r.bin_w, r.W);
  __syncthreads();
  if (tb.sy.first > tb.sy.last || tb.sx.first > tb.sx.last) return;
  float gr[kP * kP];
  {
    const float inv = 1.f / (float)(sr * sr);   // the CPU backward divides by the raw grid product
    const T* sg = reinterpret_cast<const T*>(smem_raw) + (size_t)tid * (kP * kP);
#pragma unroll
    for (int i = 0; i < kP * kP; ++i) gr[i] = to_f32<T>(sg[i]) * inv;
  }
  const int y_first = tb.sy.first, y_last = tb.sy.last, x_first = tb.sx.first, x_last = tb.sx.last;
  const size_t stage_bytes = (size_t)kFpCols * C * sizeof(T);
  T* ring[2] = {reinterpret_cast<T*>(smem_raw), reinterpret_cast<T*>(smem_raw + stage_bytes)};
  T* __restrict__ img = reinterpret_cast<T*>(g.gfeat[r.level]) + (size_t)r.batch * r.H * r.W * C;
  for (int yc = y_first; yc <= y_last; yc += kFpRows) {
    const int nrows = min(kFpRows, y_last - yc + 1);
    for (int xc = x_first; xc <= x_last; xc += kFpCols) {
      const int ncols = min(kFpCols, x_last - xc + 1);
      if (tid == 0) bulk_wait_read_all();     // row buffers of the previous chunk have been read
      __syncthreads();                        // ... and everyone is done with the staging / tables
      for (int e = tid; e < kFpRows * kPP; e += blockDim.x) {
        const int er = e >> 3, ep = e & 7;
        tb.ay[er][ep] = er < nrows ? axis_weight(tb.sy, yc + er, ep, sr) : 0.f;
        tb.ax[er][ep] = er < ncols ? axis_weight(tb.sx, xc + er, ep, sr) : 0.f;
      }
      __syncthreads();
      if (tid < 32) {
        bool live = false;
        if (tid < nrows) {
#pragma unroll
          for (int p = 0; p < kP; ++p) live |= tb.ay[tid][p] != 0.f;
        }
        const unsigned m = __ballot_sync(0xffffffffu, live);
        if (tid == 0) tb.row_live = m;
      }
      __syncthreads();
      unsigned live = tb.row_live;
      const unsigned row_bytes = (unsigned)ncols * C * sizeof(T);
      int stage = 0;
      while (live) {
        const int cur = __ffs(live) - 1;
        live &= live - 1;
        const float4 a0 = *reinterpret_cast<const float4*>(&tb.ay[cur][0]);
        const float4 a1 = *reinterpret_cast<const float4*>(&tb.ay[cur][4]);
        const float a[kP] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z};
        float t[kP];
#pragma unroll
        for (int pw = 0; pw < kP; ++pw) {
          float s = 0.f;
#pragma unroll
          for (int ph = 0; ph < kP; ++ph) s = fmaf(a[ph], gr[ph * kP + pw], s);
          t[pw] = s;
        }
        T* __restrict__ row = ring[stage] + tid;
#pragma unroll 4
        for (int x = 0; x < ncols; ++x) {
          const float4 w0 = *reinterpret_cast<const float4*>(&tb.ax[x][0]);
          const float4 w1 = *reinterpret_cast<const float4*>(&tb.ax[x][4]);
          float v = w0.x * t[0];
          v = fmaf(w0.y, t[1], v); v = fmaf(w0.z, t[2], v); v = fmaf(w0.w, t[3], v);
          v = fmaf(w1.x, t[4], v); v = fmaf(w1.y, t[5], v); v = fmaf(w1.z, t[6], v);
          row[(size_t)x * C] = from_f32<T>(v);
        }
        fence_proxy_async_smem();             // generic-proxy writes -> visible to the bulk engine
        if (tid == 0) bulk_wait_read_all();   // the other buffer (row issued one step ago) has been read
        __syncthreads();
        if (tid == 0) {
          bulk_reduce_add<T>(img + ((size_t)(yc + cur) * r.W + xc) * C, ring[stage], row_bytes);
          bulk_commit();
        }
        stage ^= 1;
      }
    }
  }
  if (tid == 0) bulk_wait_read_all();
}

// One CTA per RoI; gradients must have been zero-filled.
template <typename T>
__global__ void __launch_bounds__(kTmaThreadsMax, 3)
msroi_bwd_tma_kernel(const RoiDev g, const T* __restrict__ grad_out, const float* __restrict__ rois, int n_rois) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ RoiTables tb;
  bwd_one_roi<T>(g, tb, smem_raw, grad_out, rois, blockIdx.x);
}

// Persistent cooperative variant: per image b, zero-fill the image's gradient maps, grid-sync,
// reduce the image's RoIs (dynamic work counter).  The zero-filled lines of one image (52.9 MB
// fp32 at 608x1024) are still dirty in L2 when the reductions arrive, so DRAM sees each line once.
template <typename T>
__global__ void __launch_bounds__(kTmaThreadsMax, 3)
msroi_bwd_tma_persistent_kernel(const RoiDev g, const T* __restrict__ grad_out, const float* __restrict__ rois,
                                const int32_t* __restrict__ roi_img_offsets, int* __restrict__ counters) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ RoiTables tb;
  __shared__ int s_k;
  cg::grid_group grid = cg::this_grid();
  const int C = g.C;
  for (int b = 0; b < g.B; ++b) {
    for (int l = 0; l < g.n_levels; ++l) {
      const size_t n16 = (size_t)g.H[l] * g.W[l] * C * sizeof(T) / 16;
      uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<T*>(g.gfeat[l]) + (size_t)b * g.H[l] * g.W[l] * C);
      for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("fence.proxy.async.global;\n" ::: "memory");   // generic-proxy zeros before the bulk engine's RMW
    __threadfence();
    grid.sync();
    const int k0 = roi_img_offsets[b], k1 = roi_img_offsets[b + 1];
    while (true) {
      __syncthreads();
      if (threadIdx.x == 0) s_k = k0 + atomicAdd(&counters[b], 1);
      __syncthreads();
      const int k = s_k;
      if (k >= k1) break;
      bwd_one_roi<T>(g, tb, smem_raw, grad_out, rois, k);
    }
  }
}

// ------------------------------------------------------------------------------------------------
static bool tma_shape_ok(const dgod_roi_config* cfg, const RoiDev& g) {
  const int esz = cfg->dtype == DGOD_F32 ? 4 : 2;
  if (!g.channels_last || g.PH != kP || g.PW != kP) return false;
  if (g.sr < 1 || g.sr > 2) return false;
  if (g.C % 32 != 0 || g.C < 64 || g.C > kTmaThreadsMax || (g.C * esz) % 16 != 0) return false;
  for (int l = 0; l < g.n_levels; ++l)
    if (g.H[l] > 32000 || g.W[l] > 32000) return false;
  return true;
}

static size_t tma_smem_bytes(const RoiDev& g, int esz) {
  const size_t ring = 2 * (size_t)kFpCols * g.C * esz;
  const size_t block = (size_t)g.C * kP * kP * esz;
  return ring > block ? ring : block;
}

int msroi_fwd_tma(const dgod_roi_config* cfg, const RoiDev& g, const float* rois, int n_rois, void* out,
                  cudaStream_t st, int* handled) {
  *handled = 0;
  if (!tma_shape_ok(cfg, g) || ((uintptr_t)out & 15)) return DGOD_OK;
  for (int l = 0; l < g.n_levels; ++l)
    if ((uintptr_t)g.feat[l] & 15) return DGOD_OK;
  const int esz = cfg->dtype == DGOD_F32 ? 4 : 2;
  const size_t smem = tma_smem_bytes(g, esz);
  static size_t attr[2] = {0, 0};
  if (cfg->dtype == DGOD_F32) {
    if (smem > attr[0]) {
      DGOD_CUDA(cudaFuncSetAttribute(msroi_fwd_tma_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr[0] = smem;
    }
    msroi_fwd_tma_kernel<float><<<n_rois, g.C, smem, st>>>(g, rois, n_rois, (float*)out);
  } else {
    if (smem > attr[1]) {
      DGOD_CUDA(cudaFuncSetAttribute(msroi_fwd_tma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr[1] = smem;
    }
    msroi_fwd_tma_kernel<__nv_bfloat16><<<n_rois, g.C, smem, st>>>(g, rois, n_rois, (__nv_bfloat16*)out);
  }
  DGOD_LAUNCHED();
  *handled = 1;
  return DGOD_OK;
}

template <typename T>
static int launch_bwd_tma(const RoiDev& g, const void* grad_out, const float* rois, int n_rois,
                          const int32_t* roi_img_offsets, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const size_t smem = tma_smem_bytes(g, (int)sizeof(T));
  static size_t attr = 0, attr_p = 0;
  static int coop = -1, n_sm = 0;
  if (coop < 0) {
    int dev = 0;
    DGOD_CUDA(cudaGetDevice(&dev));
    DGOD_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    DGOD_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  const size_t need_ws = (size_t)g.B * sizeof(int);
  const bool persistent = coop && roi_img_offsets && workspace && workspace_bytes >= need_ws && g.B > 1;
  if (persistent) {
    if (smem > attr_p) {
      DGOD_CUDA(cudaFuncSetAttribute(msroi_bwd_tma_persistent_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_p = smem;
    }
    static int per_sm = 0, per_sm_c = 0;
    if (per_sm_c != g.C) {
      DGOD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, msroi_bwd_tma_persistent_kernel<T>, g.C, smem));
      per_sm_c = g.C;
    }
    if (per_sm >= 1) {
      int* counters = (int*)workspace;
      DGOD_CUDA(cudaMemsetAsync(counters, 0, need_ws, st));
      const T* go = (const T*)grad_out;
      RoiDev gg = g;
      void* args[] = {(void*)&gg, (void*)&go, (void*)&rois, (void*)&roi_img_offsets, (void*)&counters};
      DGOD_CUDA(cudaLaunchCooperativeKernel((const void*)msroi_bwd_tma_persistent_kernel<T>, dim3(per_sm * n_sm), dim3(g.C), args, smem, st));
      g_launches.fetch_add(1, std::memory_order_relaxed);
      return DGOD_OK;
    }
  }
  for (int l = 0; l < g.n_levels; ++l)
    DGOD_CUDA(cudaMemsetAsync(g.gfeat[l], 0, (size_t)g.B * g.C * g.H[l] * g.W[l] * sizeof(T), st));
  if (n_rois == 0) return DGOD_OK;
  if (smem > attr) {
    DGOD_CUDA(cudaFuncSetAttribute(msroi_bwd_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  msroi_bwd_tma_kernel<T><<<n_rois, g.C, smem, st>>>(g, (const T*)grad_out, rois, n_rois);
  DGOD_LAUNCHED();
  return DGOD_OK;
}

int msroi_bwd_tma(const dgod_roi_config* cfg, const RoiDev& g, const void* grad_out, const float* rois, int n_rois,
                  const int32_t* roi_img_offsets, void* workspace, size_t workspace_bytes, cudaStream_t st,
                  int* handled) {
  *handled = 0;
  if (!tma_shape_ok(cfg, g) || ((uintptr_t)grad_out & 15)) return DGOD_OK;
  for (int l = 0; l < g.n_levels; ++l)
    if ((uintptr_t)g.gfeat[l] & 15) return DGOD_OK;
  *handled = 1;
  if (cfg->dtype == DGOD_F32)
    return launch_bwd_tma<float>(g, grad_out, rois, n_rois, roi_img_offsets, workspace, workspace_bytes, st);
  return launch_bwd_tma<__nv_bfloat16>(g, grad_out, rois, n_rois, roi_img_offsets, workspace, workspace_bytes, st);
}

}  // namespace dgod
