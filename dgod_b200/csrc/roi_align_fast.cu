// Tuned MultiScaleRoIAlign kernels for the configuration DGOD uses (fasterrcnn.py:412-416:
// 7x7 bins, sampling_ratio 2, aligned=False, 4 FPN levels, C = 256).
//
// Layout in HBM: features are read either as NHWC (channels_last, what cuDNN's tensor-core
// convolutions produce natively) or NCHW; the output is always [K][C][PH*PW] contiguous so
// that the box head's flatten is a view.
//
// Forward, NHWC (msroi_fwd_nhwc_kernel): one CTA per RoI, lanes = channels.  A tap is then
//   one contiguous 128..256-byte read per warp straight from global/L1 — no staging — and the
//   16 taps of a bin share warp-uniform offsets and weights, which are tabulated once per RoI in
//   shared memory (separable bilinear weights: 14 row samples x 14 column samples).  The RoI's
//   [C][49] result is assembled in shared memory and leaves as contiguous 128-bit stores.
//   Roofline: HBM — output K*C*49*s written once, every touched feature line read once (the
//   image's maps stay L2-resident while its RoIs are processed); the on-chip limit is L1
//   bandwidth (16 taps * s bytes per output element).
// Backward (msroi_bwd_tile_kernel): gather formulation, no atomics, deterministic.  A CTA owns a
//   16x16-pixel tile of one image/level and a slab of channels, accumulates in shared memory the
//   contributions of every RoI whose footprint meets the tile (found by scanning the image's
//   RoIs), and writes the tile exactly once — zeros included, so no memset and no
//   read-modify-write traffic: grad_out read + grad_in written once = the algorithmic bytes.
#include "roi_common.cuh"

namespace dgod {

constexpr int kMaxS = 16;  // samples per axis handled by the tuned kernels (PH*sr <= 16)

struct AxisTab {       // per-RoI sample tables (one axis)
  int lo[kMaxS], hi[kMaxS];     // element offsets (already multiplied by the pixel pitch)
  float l[kMaxS], h[kMaxS];     // weights; both zero when the sample is out of range
};

__device__ __forceinline__ void fill_axis(AxisTab& t, int i, float start, float bin, int grid, int size,
                                          int pitch) {
  const AxisTap a = axis_tap(sample_coord(start, i / grid, bin, i % grid, grid), size);
  t.lo[i] = a.lo * pitch;
  t.hi[i] = a.hi * pitch;
  t.l[i] = a.valid ? a.l : 0.f;
  t.h[i] = a.valid ? a.h : 0.f;
}

template <typename T, int VEC> struct VecLoad;
template <> struct VecLoad<float, 1> {
  static __device__ __forceinline__ void ld(const float* p, float* v) { v[0] = __ldg(p); }
};
template <> struct VecLoad<float, 2> {
  static __device__ __forceinline__ void ld(const float* p, float* v) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p)); v[0] = t.x; v[1] = t.y;
  }
};
template <> struct VecLoad<float, 4> {
  static __device__ __forceinline__ void ld(const float* p, float* v) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <> struct VecLoad<__nv_bfloat16, 2> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* v) {
    const float2 f = __bfloat1622float2(__ldg(reinterpret_cast<const __nv_bfloat162*>(p))); v[0] = f.x; v[1] = f.y;
  }
};
template <> struct VecLoad<__nv_bfloat16, 4> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* v) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
};

// ------------------------------------------------------------------------------------------------
// Forward, NHWC.  blockDim.x * VEC == C, sampling ratio SR in {1,2} (compile time), PH*SR <= 16.
//
// ncu showed the first version issue-bound (66 % issue-active, ~230 instructions per output
// element: per-thread address and weight arithmetic that is identical for every channel).  Here
// that work is done ONCE per RoI: a shared-memory table holds, for each bin, the 4*SR*SR taps as
// (element offset, weight product) pairs; the per-channel loop is then one 8-byte table read, one
// 128-bit feature load and 2 fp32 instructions per channel and tap.  The taps of bin b+1 are
// loaded while bin b is being reduced (explicit double buffer in registers).
template <typename T, int VEC, int SR>
__global__ void __launch_bounds__(256)
msroi_fwd_nhwc_kernel(const RoiDev g, const float* __restrict__ rois, int n_rois, T* __restrict__ out) {
  constexpr int NTAP = 4 * SR * SR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ AxisTab ty, tx;
  __shared__ RoiGeom s_geo;
  const int k = blockIdx.x;
  if (threadIdx.x == 0) s_geo = roi_geometry(g, rois + (size_t)k * 5);
  __syncthreads();
  const RoiGeom r = s_geo;
  const int PH = g.PH, PW = g.PW, C = g.C;
  const int nbin = PH * PW;
  int2* s_tab = reinterpret_cast<int2*>(smem_raw);                                   // [nbin][NTAP]
  T* s_out = reinterpret_cast<T*>(smem_raw + (size_t)nbin * NTAP * sizeof(int2));     // [C][nbin]
  T* __restrict__ o = out + (size_t)k * C * nbin;
  const bool usable = r.batch >= 0 && r.batch < g.B;
  if (usable) {
    const int ny = PH * SR, nx = PW * SR;
    if (threadIdx.x < ny) fill_axis(ty, threadIdx.x, r.start_h, r.bin_h, SR, r.H, r.W * C);
    else if (threadIdx.x >= 32 && threadIdx.x < 32 + nx) fill_axis(tx, threadIdx.x - 32, r.start_w, r.bin_w, SR, r.W, C);
  }
  __syncthreads();
  if (usable) {
    for (int e = threadIdx.x; e < nbin * NTAP; e += blockDim.x) {
      const int bin = e / NTAP, tap = e - bin * NTAP;
      const int ph = bin / PW, pw = bin - ph * PW;
      const int smp = tap >> 2, j = tap & 3;           // sample (iy, ix) in the CPU kernel's order, tap 1..4
      const int sy = ph * SR + smp / SR, sx = pw * SR + smp % SR;
      const int yo = (j & 2) ? ty.hi[sy] : ty.lo[sy];
      const int xo = (j & 1) ? tx.hi[sx] : tx.lo[sx];
      const float wy = (j & 2) ? ty.l[sy] : ty.h[sy];
      const float wx = (j & 1) ? tx.l[sx] : tx.h[sx];
      s_tab[e] = make_int2(yo + xo, __float_as_int(__fmul_rn(wy, wx)));
    }
  }
  __syncthreads();
  const int c0 = threadIdx.x * VEC;
  if (usable) {
    const T* __restrict__ img = reinterpret_cast<const T*>(g.feat[r.level]) + (size_t)r.batch * r.H * r.W * C + c0;
    const float count = r.count;
    const bool pow2 = (SR & (SR - 1)) == 0;            // then acc / count == acc * (1/count) exactly
    const float inv_count = 1.f / count;
    float va[NTAP][VEC], vb[NTAP][VEC], wa[NTAP], wb[NTAP];
    auto load_bin = [&](int bin, float (&v)[NTAP][VEC], float (&w)[NTAP]) {
#pragma unroll
      for (int t = 0; t < NTAP; ++t) {
        const int2 e = s_tab[bin * NTAP + t];
        w[t] = __int_as_float(e.y);
        VecLoad<T, VEC>::ld(img + e.x, v[t]);
      }
    };
    auto reduce_bin = [&](int bin, const float (&v)[NTAP][VEC], const float (&w)[NTAP]) {
      float acc[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
#pragma unroll
      for (int smp = 0; smp < SR * SR; ++smp) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          // the CPU kernel's order: ((w1 v1 + w2 v2) + w3 v3) + w4 v4, accumulated sample by sample
          const float sum = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w[4 * smp], v[4 * smp][e]),
                                                          __fmul_rn(w[4 * smp + 1], v[4 * smp + 1][e])),
                                                __fmul_rn(w[4 * smp + 2], v[4 * smp + 2][e])),
                                      __fmul_rn(w[4 * smp + 3], v[4 * smp + 3][e]));
          acc[e] = __fadd_rn(acc[e], sum);
        }
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e)
        s_out[(c0 + e) * nbin + bin] = from_f32<T>(pow2 ? __fmul_rn(acc[e], inv_count) : __fdiv_rn(acc[e], count));
    };
    load_bin(0, va, wa);
    int bin = 0;
    for (; bin + 2 < nbin; bin += 2) {
      load_bin(bin + 1, vb, wb);
      reduce_bin(bin, va, wa);
      load_bin(bin + 2, va, wa);
      reduce_bin(bin + 1, vb, wb);
    }
    if (bin + 1 < nbin) {        // two bins left: bin (in a) and bin + 1
      load_bin(bin + 1, vb, wb);
      reduce_bin(bin, va, wa);
      reduce_bin(bin + 1, vb, wb);
    } else {
      reduce_bin(bin, va, wa);
    }
  } else {
    for (int e = 0; e < VEC; ++e)
      for (int b = 0; b < nbin; ++b) s_out[(c0 + e) * nbin + b] = from_f32<T>(0.f);
  }
  __syncthreads();
  // contiguous [C][PH*PW] block -> global, 128-bit stores
  const int n16 = C * nbin * (int)sizeof(T) / 16;
  const uint4* __restrict__ src = reinterpret_cast<const uint4*>(s_out);
  uint4* __restrict__ dst = reinterpret_cast<uint4*>(o);
  for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------
// Backward, NHWC, vector reductions (algo 1 on channels_last gradients).  Mirror image of the
// forward: one CTA per RoI, lanes = channels, the same per-bin (offset, weight) table.  The RoI's
// [C][49] gradient block is staged through shared memory (contiguous 128-bit reads), then every
// tap is ONE 16-byte `red.global.add.v4.f32` per lane — 512 contiguous bytes per warp, so the L2
// atomic units see full sectors instead of the 4-byte scatter of the reference kernel.
// grad_in must be zero-filled first (done by the launcher).  Accumulation order across RoIs is
// not deterministic (like the reference CUDA kernel); the tile-gather kernel below is.
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int SR>
__global__ void __launch_bounds__(256)
msroi_bwd_red_nhwc_kernel(const RoiDev g, const float* __restrict__ grad_out, const float* __restrict__ rois,
                          int n_rois) {
  constexpr int NTAP = 4 * SR * SR;
  constexpr int VEC = 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ AxisTab ty, tx;
  __shared__ RoiGeom s_geo;
  const int k = blockIdx.x;
  if (threadIdx.x == 0) s_geo = roi_geometry(g, rois + (size_t)k * 5);
  __syncthreads();
  const RoiGeom r = s_geo;
  if (r.batch < 0 || r.batch >= g.B) return;
  const int PH = g.PH, PW = g.PW, C = g.C;
  const int nbin = PH * PW;
  int2* s_tab = reinterpret_cast<int2*>(smem_raw);                                        // [nbin][NTAP]
  float* s_g = reinterpret_cast<float*>(smem_raw + (size_t)nbin * NTAP * sizeof(int2));   // [C][nbin]
  const int ny = PH * SR, nx = PW * SR;
  if (threadIdx.x < ny) fill_axis(ty, threadIdx.x, r.start_h, r.bin_h, SR, r.H, r.W * C);
  else if (threadIdx.x >= 32 && threadIdx.x < 32 + nx) fill_axis(tx, threadIdx.x - 32, r.start_w, r.bin_w, SR, r.W, C);
  {
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(grad_out + (size_t)k * C * nbin);
    uint4* dst = reinterpret_cast<uint4*>(s_g);
    for (int i = threadIdx.x; i < C * nbin / 4; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nbin * NTAP; e += blockDim.x) {
    const int bin = e / NTAP, tap = e - bin * NTAP;
    const int ph = bin / PW, pw = bin - ph * PW;
    const int smp = tap >> 2, j = tap & 3;
    const int sy = ph * SR + smp / SR, sx = pw * SR + smp % SR;
    const int yo = (j & 2) ? ty.hi[sy] : ty.lo[sy];
    const int xo = (j & 1) ? tx.hi[sx] : tx.lo[sx];
    const float wy = (j & 2) ? ty.l[sy] : ty.h[sy];
    const float wx = (j & 1) ? tx.l[sx] : tx.h[sx];
    s_tab[e] = make_int2(yo + xo, __float_as_int(__fmul_rn(wy, wx)));
  }
  __syncthreads();
  const int c0 = threadIdx.x * VEC;
  float* __restrict__ img = reinterpret_cast<float*>(g.gfeat[r.level]) + (size_t)r.batch * r.H * r.W * C + c0;
  const float inv_cnt = 1.f / (float)(SR * SR);   // SR*SR is a power of two: (g*w)/count == (g*w)*inv exactly
  for (int bin = 0; bin < nbin; ++bin) {
    float gv[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) gv[v] = s_g[(c0 + v) * nbin + bin];
#pragma unroll
    for (int t = 0; t < NTAP; ++t) {
      const int2 e = s_tab[bin * NTAP + t];
      const float w = __int_as_float(e.y);
      if (w != 0.f)   // out-of-range samples (and exactly-zero weights) add nothing
        red_add_v4(img + e.x, __fmul_rn(__fmul_rn(gv[0], w), inv_cnt), __fmul_rn(__fmul_rn(gv[1], w), inv_cnt),
                   __fmul_rn(__fmul_rn(gv[2], w), inv_cnt), __fmul_rn(__fmul_rn(gv[3], w), inv_cnt));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Backward, tile gather.
constexpr int kTile = 16;          // tile is kTile x kTile pixels
constexpr int kBwdThreads = 256;   // 8 warps; warp w owns pixels w, w+8, ... of the tile
constexpr int kMaxList = 1024;     // RoIs listed per pass over the image's RoIs

// Per-RoI record written once by msroi_prep_kernel and read by every tile CTA that meets the RoI.
struct RoiPrep {
  int level, batch;                 // batch < 0: RoI unusable (bad batch index)
  short fy0, fy1, fx0, fx1;         // footprint (pixels touched by any tap), inclusive
  short ylo[kMaxS], yhi[kMaxS], xlo[kMaxS], xhi[kMaxS];
  float yl[kMaxS], yh[kMaxS], xl[kMaxS], xh[kMaxS];   // zero weights mark out-of-range samples
};
static_assert(sizeof(RoiPrep) % 16 == 0, "RoiPrep is copied in 16-byte chunks");

__global__ void __launch_bounds__(128)
msroi_prep_kernel(const RoiDev g, const float* __restrict__ rois, int n_rois, RoiPrep* __restrict__ prep) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= n_rois) return;
  const RoiGeom r = roi_geometry(g, rois + (size_t)k * 5);
  const int sr = g.sr, ny = g.PH * sr, nx = g.PW * sr;
  RoiPrep& p = prep[k];
  int lo = 0x7fff, hi = -1;
  const bool is_x = lane >= 16;
  const int i = lane & 15;
  if (i < (is_x ? nx : ny)) {
    const AxisTap a = is_x ? axis_tap(sample_coord(r.start_w, i / sr, r.bin_w, i % sr, sr), r.W)
                           : axis_tap(sample_coord(r.start_h, i / sr, r.bin_h, i % sr, sr), r.H);
    (is_x ? p.xlo : p.ylo)[i] = (short)a.lo;
    (is_x ? p.xhi : p.yhi)[i] = (short)a.hi;
    (is_x ? p.xl : p.yl)[i] = a.valid ? a.l : 0.f;
    (is_x ? p.xh : p.yh)[i] = a.valid ? a.h : 0.f;
    if (a.valid) { lo = a.lo; hi = a.hi; }
  }
#pragma unroll
  for (int d = 8; d >= 1; d >>= 1) {   // min / max inside each half-warp
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
  }
  const int xlo = __shfl_sync(0xffffffffu, lo, 16), xhi = __shfl_sync(0xffffffffu, hi, 16);
  if (lane == 0) {
    p.level = r.level;
    p.batch = (r.batch >= 0 && r.batch < g.B) ? r.batch : -1;
    p.fy0 = (short)lo; p.fy1 = (short)hi; p.fx0 = (short)xlo; p.fx1 = (short)xhi;   // empty when hi < lo
  }
}

struct BwdCsr {                    // (bin index, weight) pairs per tile row / column for one RoI
  unsigned char row_ptr[kTile + 1], col_ptr[kTile + 1];
  unsigned char row_bin[2 * kMaxS], col_bin[2 * kMaxS];
  float row_w[2 * kMaxS], col_w[2 * kMaxS];
};

struct BwdTiles { int first[DGOD_MAX_LEVELS + 1]; int tiles_x[DGOD_MAX_LEVELS]; };

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

// CTA = (tile of one level, image, channel slab of CB = 32*VEC channels).  Shared memory:
// fp32 accumulator tile [256 pixels][CB+1], two gradient blocks [CB][nbin] (raw layout: lanes =
// channels read it conflict-free because nbin = 49 is odd) and two RoiPrep records (double buffer).
template <typename T, int VEC, bool NHWC>
__global__ void __launch_bounds__(kBwdThreads)
msroi_bwd_tile_kernel(const RoiDev g, const BwdTiles tiles, const T* __restrict__ grad_out,
                      const RoiPrep* __restrict__ prep, int n_rois,
                      const int32_t* __restrict__ roi_img_offsets) {
  constexpr int CB = 32 * VEC;
  constexpr int PITCH = CB + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int PH = g.PH, PW = g.PW, C = g.C, sr = g.sr, nbin = PH * PW;
  const int gblock = (CB * nbin * (int)sizeof(T) + 15) & ~15;          // bytes of one gradient block
  RoiPrep* s_prep = reinterpret_cast<RoiPrep*>(smem_raw + 2 * gblock);   // [2]
  float* s_acc = reinterpret_cast<float*>(smem_raw + 2 * gblock + 2 * sizeof(RoiPrep));
  __shared__ int s_list[kMaxList];
  __shared__ int s_nlist;
  __shared__ int s_warp_cnt[kBwdThreads / 32];

  int level = 0;
  while (level + 1 < g.n_levels && (int)blockIdx.x >= tiles.first[level + 1]) ++level;
  const int tile = blockIdx.x - tiles.first[level];
  const int H = g.H[level], W = g.W[level];
  const int b = blockIdx.y, slab = blockIdx.z;
  const int ty0 = (tile / tiles.tiles_x[level]) * kTile, tx0 = (tile % tiles.tiles_x[level]) * kTile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cbase = slab * CB;

  for (int i = threadIdx.x; i < kTile * kTile * PITCH; i += blockDim.x) s_acc[i] = 0.f;

  int k_begin = 0, k_end = n_rois;
  if (roi_img_offsets) { k_begin = roi_img_offsets[b]; k_end = roi_img_offsets[b + 1]; }

  auto prefetch = [&](int k, int buf) {
    const char* src = reinterpret_cast<const char*>(grad_out + ((size_t)k * C + cbase) * nbin);
    char* dst = reinterpret_cast<char*>(smem_raw) + buf * gblock;
    const int n16 = CB * nbin * (int)sizeof(T) / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) cp_async16(dst + 16 * i, src + 16 * i);
    const char* psrc = reinterpret_cast<const char*>(prep + k);
    char* pdst = reinterpret_cast<char*>(s_prep + buf);
    for (int i = threadIdx.x; i < (int)sizeof(RoiPrep) / 16; i += blockDim.x) cp_async16(pdst + 16 * i, psrc + 16 * i);
    cp_async_commit();
  };

  for (int chunk0 = k_begin; chunk0 < k_end; chunk0 += kMaxList) {
    // ---- 1. list the RoIs of this chunk whose footprint meets the tile (ordered by index) ----
    __syncthreads();
    if (threadIdx.x == 0) s_nlist = 0;
    __syncthreads();
    const int chunk1 = min(chunk0 + kMaxList, k_end);
    for (int base = chunk0; base < chunk1; base += blockDim.x) {
      const int k = base + threadIdx.x;
      bool hit = false;
      if (k < chunk1) {
        const int4 hdr = __ldg(reinterpret_cast<const int4*>(prep + k));   // level, batch, (fy0,fy1), (fx0,fx1)
        const int fy0 = (short)(hdr.z & 0xffff), fy1 = (short)(hdr.z >> 16);
        const int fx0 = (short)(hdr.w & 0xffff), fx1 = (short)(hdr.w >> 16);
        hit = hdr.x == level && hdr.y == b && fy0 < ty0 + kTile && fy1 >= ty0 && fx0 < tx0 + kTile && fx1 >= tx0;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) s_warp_cnt[warp] = __popc(bal);
      __syncthreads();
      int off = s_nlist;
      for (int w = 0; w < warp; ++w) off += s_warp_cnt[w];
      if (hit) s_list[off + __popc(bal & ((1u << lane) - 1u))] = k;
      __syncthreads();
      if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < kBwdThreads / 32; ++w) tot += s_warp_cnt[w];
        s_nlist += tot;
      }
      __syncthreads();
    }
    const int nlist = s_nlist;
    if (nlist == 0) continue;

    // ---- 2. accumulate every listed RoI; gradient block + tables of RoI i+1 stream in meanwhile ----
    prefetch(s_list[0], 0);
    for (int li = 0; li < nlist; ++li) {
      const int buf = li & 1;
      cp_async_wait_all();
      __syncthreads();                       // block li landed; everyone is done with block li-1
      if (li + 1 < nlist) prefetch(s_list[li + 1], buf ^ 1);
      const RoiPrep& rp = s_prep[buf];
      // Every warp works on its own: lane e < 2*n samples holds tap e of each axis
      // (sample e/2, low tap for even e, high tap for odd e): coordinate, weight.  A sample adds
      // h*g at its low tap and l*g at its high tap, both when they coincide at the border.
      const int e_s = lane >> 1, e_hi = lane & 1;
      int ycoord = -1, xcoord = -1;
      float yw = 0.f, xw = 0.f;
      if (e_s < PH * sr) {
        const bool live = rp.yh[e_s] != 0.f || rp.yl[e_s] != 0.f;
        if (live) { ycoord = e_hi ? rp.yhi[e_s] : rp.ylo[e_s]; yw = e_hi ? rp.yl[e_s] : rp.yh[e_s]; }
      }
      if (e_s < PW * sr) {
        const bool live = rp.xh[e_s] != 0.f || rp.xl[e_s] != 0.f;
        if (live) { xcoord = e_hi ? rp.xhi[e_s] : rp.xlo[e_s]; xw = e_hi ? rp.xl[e_s] : rp.xh[e_s]; }
      }
      // footprint of this RoI inside the tile
      const int ry0 = max((int)rp.fy0, ty0), ry1 = min((int)rp.fy1, ty0 + kTile - 1);
      const int rx0 = max((int)rp.fx0, tx0), rx1 = min((int)rp.fx1, tx0 + kTile - 1);
      const int ncols = rx1 - rx0 + 1, npix = (ry1 - ry0 + 1) * ncols;
      const float cnt = (float)(sr * sr);       // the CPU backward divides by the raw grid product
      const bool pow2 = (sr & (sr - 1)) == 0;   // then g*w/count == (g*w) * (1/count) exactly
      const float inv_cnt = 1.f / cnt;
      const T* __restrict__ gb = reinterpret_cast<const T*>(smem_raw + buf * gblock) + lane * nbin;
      for (int i = warp; i < npix; i += kBwdThreads / 32) {
        const int yy = i / ncols, xx = i - yy * ncols;
        const int y = ry0 + yy, x = rx0 + xx;
        unsigned rows = __ballot_sync(0xffffffffu, ycoord == y);
        const unsigned cols = __ballot_sync(0xffffffffu, xcoord == x);
        if (rows == 0u || cols == 0u) continue;
        float acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
        while (rows) {
          const int a = __ffs(rows) - 1;
          rows &= rows - 1;
          const float wy = __shfl_sync(0xffffffffu, yw, a);
          const int rowbin = ((a >> 1) / sr) * PW;
          unsigned cc = cols;
          while (cc) {
            const int q = __ffs(cc) - 1;
            cc &= cc - 1;
            const float w = __fmul_rn(wy, __shfl_sync(0xffffffffu, xw, q));
            const T* gp = gb + rowbin + (q >> 1) / sr;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              const float gw = __fmul_rn(to_f32<T>(gp[32 * v * nbin]), w);   // g*w/count as in the CPU kernel
              acc[v] = __fadd_rn(acc[v], pow2 ? __fmul_rn(gw, inv_cnt) : __fdiv_rn(gw, cnt));
            }
          }
        }
        float* ap = s_acc + ((y - ty0) * kTile + (x - tx0)) * PITCH + lane;
#pragma unroll
        for (int v = 0; v < VEC; ++v) ap[32 * v] += acc[v];
      }
    }
  }
  __syncthreads();
  // ---- 3. write the tile once (zeros included: no memset, no read-modify-write) ----
  T* __restrict__ gin = reinterpret_cast<T*>(g.gfeat[level]) + (size_t)b * C * H * W;
  if (NHWC) {
    for (int p = warp; p < kTile * kTile; p += kBwdThreads / 32) {
      const int y = ty0 + p / kTile, x = tx0 + p % kTile;
      if (y >= H || x >= W) continue;
      T* dst = gin + ((size_t)y * W + x) * C + cbase + lane;
      const float* ap = s_acc + p * PITCH + lane;
#pragma unroll
      for (int v = 0; v < VEC; ++v) dst[32 * v] = from_f32<T>(ap[32 * v]);
    }
  } else {
    // NCHW: lanes run along x inside a tile row (16 contiguous pixels = 64 B per channel row)
    for (int i = threadIdx.x; i < CB * kTile * kTile; i += blockDim.x) {
      const int xx = i % kTile, yy = (i / kTile) % kTile, c = i / (kTile * kTile);
      const int y = ty0 + yy, x = tx0 + xx;
      if (y >= H || x >= W) continue;
      gin[((size_t)(cbase + c) * H + y) * W + x] = from_f32<T>(s_acc[(yy * kTile + xx) * PITCH + c]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
static bool fast_shape_ok(const dgod_roi_config* cfg) {
  return cfg->sampling_ratio >= 1 && cfg->sampling_ratio <= 2 && cfg->pooled_h * cfg->sampling_ratio <= kMaxS &&
         cfg->pooled_w * cfg->sampling_ratio <= kMaxS;
}

template <typename T, int VEC, int SR>
static int launch_fwd_nhwc_sr(const RoiDev& g, const float* rois, int n_rois, void* out, cudaStream_t st) {
  const size_t smem = (size_t)g.C * g.PH * g.PW * sizeof(T) + (size_t)g.PH * g.PW * 4 * SR * SR * sizeof(int2);
  static size_t attr_smem = 0;
  if (smem > 48 * 1024 && smem > attr_smem) {
    DGOD_CUDA(cudaFuncSetAttribute(msroi_fwd_nhwc_kernel<T, VEC, SR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  msroi_fwd_nhwc_kernel<T, VEC, SR><<<n_rois, g.C / VEC, smem, st>>>(g, rois, n_rois, (T*)out);
  DGOD_LAUNCHED();
  return DGOD_OK;
}

template <typename T, int VEC>
static int launch_fwd_nhwc(const RoiDev& g, const float* rois, int n_rois, void* out, cudaStream_t st) {
  return g.sr == 2 ? launch_fwd_nhwc_sr<T, VEC, 2>(g, rois, n_rois, out, st)
                   : launch_fwd_nhwc_sr<T, VEC, 1>(g, rois, n_rois, out, st);
}

int msroi_fwd_fast(const dgod_roi_config* cfg, const RoiDev& g, const float* rois, int n_rois, void* out,
                   cudaStream_t st, int* handled) {
  *handled = 0;
  if (!fast_shape_ok(cfg) || !g.channels_last) return DGOD_OK;
  const size_t esz = cfg->dtype == DGOD_F32 ? 4 : 2;
  const size_t block_bytes = (size_t)g.C * g.PH * g.PW * esz;
  if (block_bytes % 16 != 0 || block_bytes > 200 * 1024 || ((uintptr_t)out & 15)) return DGOD_OK;
  for (int l = 0; l < g.n_levels; ++l)
    if ((uintptr_t)g.feat[l] & 15) return DGOD_OK;
  int rc = DGOD_OK;
  if (cfg->dtype == DGOD_F32) {
    if (g.C % 128 == 0 && g.C / 4 <= 256 && g.C / 4 >= 32 + kMaxS) rc = launch_fwd_nhwc<float, 4>(g, rois, n_rois, out, st);
    else if (g.C % 64 == 0 && g.C / 2 <= 256 && g.C / 2 >= 32 + kMaxS) rc = launch_fwd_nhwc<float, 2>(g, rois, n_rois, out, st);
    else return DGOD_OK;
  } else {
    if (g.C % 128 == 0 && g.C / 4 <= 256 && g.C / 4 >= 32 + kMaxS) rc = launch_fwd_nhwc<__nv_bfloat16, 4>(g, rois, n_rois, out, st);
    else if (g.C % 64 == 0 && g.C / 2 <= 256 && g.C / 2 >= 32 + kMaxS) rc = launch_fwd_nhwc<__nv_bfloat16, 2>(g, rois, n_rois, out, st);
    else return DGOD_OK;
  }
  *handled = 1;
  return rc;
}

int msroi_bwd_red(const dgod_roi_config* cfg, const RoiDev& g, const void* grad_out, const float* rois,
                  int n_rois, cudaStream_t st, int* handled) {
  *handled = 0;
  if (!g.channels_last || cfg->dtype != DGOD_F32 || !fast_shape_ok(cfg) || g.C % 4 != 0 || g.C / 4 > 256 ||
      g.C / 4 < 32 + kMaxS || ((size_t)g.C * g.PH * g.PW) % 4 != 0 || ((uintptr_t)grad_out & 15))
    return DGOD_OK;
  for (int l = 0; l < g.n_levels; ++l)
    if ((uintptr_t)g.gfeat[l] & 15) return DGOD_OK;
  const int sr = g.sr;
  const size_t smem = (size_t)g.C * g.PH * g.PW * sizeof(float) + (size_t)g.PH * g.PW * 4 * sr * sr * sizeof(int2);
  static size_t attr_smem[3] = {0, 0, 0};
  if (smem > 48 * 1024 && smem > attr_smem[sr]) {
    if (sr == 2) DGOD_CUDA(cudaFuncSetAttribute(msroi_bwd_red_nhwc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else DGOD_CUDA(cudaFuncSetAttribute(msroi_bwd_red_nhwc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem[sr] = smem;
  }
  if (sr == 2) msroi_bwd_red_nhwc_kernel<2><<<n_rois, g.C / 4, smem, st>>>(g, (const float*)grad_out, rois, n_rois);
  else msroi_bwd_red_nhwc_kernel<1><<<n_rois, g.C / 4, smem, st>>>(g, (const float*)grad_out, rois, n_rois);
  DGOD_LAUNCHED();
  *handled = 1;
  return DGOD_OK;
}

size_t msroi_bwd_workspace(int n_rois) { return align_up((size_t)(n_rois > 0 ? n_rois : 1) * sizeof(RoiPrep), 256); }

template <typename T, int VEC, bool NHWC>
static int launch_bwd_tile(const RoiDev& g, const void* grad_out, const float* rois, int n_rois,
                           const int32_t* offs, RoiPrep* prep, cudaStream_t st) {
  constexpr int CB = 32 * VEC;
  const int nbin = g.PH * g.PW;
  const size_t gblock = ((size_t)CB * nbin * sizeof(T) + 15) & ~(size_t)15;
  const size_t smem = 2 * gblock + 2 * sizeof(RoiPrep) + (size_t)kTile * kTile * (CB + 1) * sizeof(float);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    DGOD_CUDA(cudaFuncSetAttribute(msroi_bwd_tile_kernel<T, VEC, NHWC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  msroi_prep_kernel<<<cdiv(n_rois, 4), 128, 0, st>>>(g, rois, n_rois, prep);
  DGOD_LAUNCHED();
  BwdTiles tiles;
  int total = 0;
  for (int l = 0; l < g.n_levels; ++l) {
    const int tx = (g.W[l] + kTile - 1) / kTile, ty = (g.H[l] + kTile - 1) / kTile;
    tiles.first[l] = total; tiles.tiles_x[l] = tx;
    total += tx * ty;
  }
  for (int l = g.n_levels; l <= DGOD_MAX_LEVELS; ++l) tiles.first[l] = total;
  dim3 grid(total, g.B, g.C / CB);
  msroi_bwd_tile_kernel<T, VEC, NHWC><<<grid, kBwdThreads, smem, st>>>(g, tiles, (const T*)grad_out, prep, n_rois, offs);
  DGOD_LAUNCHED();
  return DGOD_OK;
}

int msroi_bwd_fast(const dgod_roi_config* cfg, const RoiDev& g, const void* grad_out, const float* rois,
                   int n_rois, const int32_t* roi_img_offsets, void* workspace, size_t workspace_bytes,
                   cudaStream_t st, int* handled) {
  *handled = 0;
  const size_t esz = cfg->dtype == DGOD_F32 ? 4 : 2;
  if (!fast_shape_ok(cfg) || g.C % 64 != 0 || g.B > 65535 || g.C / 64 > 65535) return DGOD_OK;
  if (((size_t)64 * g.PH * g.PW * esz) % 16 != 0 || ((uintptr_t)grad_out & 15)) return DGOD_OK;
  for (int l = 0; l < g.n_levels; ++l)
    if (g.H[l] > 32000 || g.W[l] > 32000) return DGOD_OK;
  if (!workspace || workspace_bytes < msroi_bwd_workspace(n_rois)) {
    set_error("roi_align backward: workspace too small (%zu < %zu)", workspace_bytes, msroi_bwd_workspace(n_rois));
    return DGOD_ERR_WORKSPACE;
  }
  RoiPrep* prep = (RoiPrep*)workspace;
  int rc;
  const bool wide = g.C % 128 == 0;
  if (cfg->dtype == DGOD_F32) {
    if (wide) rc = g.channels_last ? launch_bwd_tile<float, 4, true>(g, grad_out, rois, n_rois, roi_img_offsets, prep, st)
                                   : launch_bwd_tile<float, 4, false>(g, grad_out, rois, n_rois, roi_img_offsets, prep, st);
    else rc = g.channels_last ? launch_bwd_tile<float, 2, true>(g, grad_out, rois, n_rois, roi_img_offsets, prep, st)
                              : launch_bwd_tile<float, 2, false>(g, grad_out, rois, n_rois, roi_img_offsets, prep, st);
  } else {
    if (wide) rc = g.channels_last ? launch_bwd_tile<__nv_bfloat16, 4, true>(g, grad_out, rois, n_rois, roi_img_offsets, prep, st)
                                   : launch_bwd_tile<__nv_bfloat16, 4, false>(g, grad_out, rois, n_rois, roi_img_offsets, prep, st);
    else rc = g.channels_last ? launch_bwd_tile<__nv_bfloat16, 2, true>(g, grad_out, rois, n_rois, roi_img_offsets, prep, st)
                              : launch_bwd_tile<__nv_bfloat16, 2, false>(g, grad_out, rois, n_rois, roi_img_offsets, prep, st);
  }
  *handled = 1;
  return rc;
}

}  // namespace dgod
