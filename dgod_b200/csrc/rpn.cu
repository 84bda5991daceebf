// RPN proposal stage: per-level top-k on the raw objectness maps, analytic anchors, decode of
// the survivors only, clip / small-box / score filters, per-level NMS, score-ordered merge.
// 
//
// Reference behaviour restated (TV = torchvision 0.26.0; called from fasterrcnn.py:166-182):
//   anchors            TV models/detection/anchor_utils.py:58-74,84-133
//   permute/concat     TV models/detection/rpn.py:81-110  -> candidate order (level, y, x, a)
//   decode             TV models/detection/_utils.py:186-224 (weights 1,1,1,1; clamp log(1000/16))
//   per-level top-k    TV models/detection/rpn.py:231-240,263-272
//   filters + NMS      TV models/detection/rpn.py:276-297 (clip, remove_small 1e-3, score >= 0,
//                      batched_nms per level, keep[:post_nms_top_n])
//
// The reference decodes all A anchors per image (A = 155k..268k) and materialises anchors,
// permuted logits and permuted deltas; here a candidate is touched only if it survives the
// per-level top-k, so the compulsory traffic is 4 B per anchor (objectness) + 16 B per survivor.
// Where torch.topk leaves the order of *equal logits* implementation-defined, this kernel uses
// ascending anchor index ("ties").
#include "nms_core.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace dgod {

constexpr int kTopkThreads = 512;    // 2 CTAs per SM (1024 threads: one, i.e. ~15 resident clusters and three waves at B=8 x 5 levels)
constexpr int kTopkCluster = 8;        // CTAs (one cluster) sharing the objectness of one (level, image)
constexpr int kSelUnroll = 8;          // independent loads in flight per thread in the selection passes

struct RpnDev {
  const float* obj[DGOD_MAX_LEVELS];
  const float* del[DGOD_MAX_LEVELS];
  int n_levels, A, n_img, k_tot, flat;
  int H[DGOD_MAX_LEVELS], W[DGOD_MAX_LEVELS], sh[DGOD_MAX_LEVELS], sw[DGOD_MAX_LEVELS];
  int n_l[DGOD_MAX_LEVELS];     // anchors per level
  int a_off[DGOD_MAX_LEVELS];   // anchor offset of the level inside an image (flat layout)
  int k_l[DGOD_MAX_LEVELS];     // min(pre_nms_top_n, n_l)
  int c_off[DGOD_MAX_LEVELS];   // candidate offset of the level inside an image
  int a_total;
  float min_size, score_thresh, xform_clip;
  float cell[DGOD_MAX_LEVELS][DGOD_MAX_CELL_ANCHORS][4];
};

// ---- cluster-wide radix select: threshold of the k smallest 32-bit keys -----------------------
// The kTopkCluster CTAs of a cluster each own the slice [m0, m1) of the level's elements.  Per
// 8-bit digit: local histogram in shared memory, cluster barrier, rank 0 adds the eight histograms
// through distributed shared memory and picks the digit, second cluster barrier, every CTA reads
// the decision from rank 0.  keyfn(m, key) -> bool participates.  On return (all threads of all
// CTAs): T = k-th smallest key, n_lt = #keys < T, n_eq = #keys == T.
struct SelShared {
  uint32_t hist[256];
  uint32_t total[256];   // rank 0 only
  uint32_t bc[4];        // rank 0 only: chosen digit, #keys below it, #keys in it
};

template <typename KeyFn>
__device__ void radix_select(KeyFn keyfn, int m0, int m1, int k, SelShared* sh, cg::cluster_group& cluster,
                             uint32_t& T, int& n_lt, int& n_eq) {
  uint32_t prefix = 0u, pmask = 0u;
  int remaining = k, lt_total = 0, eq = 0;
  const int lane = threadIdx.x & 31;
  const unsigned rank = cluster.block_rank();
  SelShared* sh0 = cluster.map_shared_rank(sh, 0);
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh->hist[i] = 0u;
    __syncthreads();
    for (int base = m0; base < m1; base += kSelUnroll * blockDim.x) {
      // issue the loads of kSelUnroll elements before the first histogram update (the update's
      // warp-wide match would otherwise serialise one global-memory latency per element)
      uint32_t keys[kSelUnroll];
      bool parts[kSelUnroll];
#pragma unroll
      for (int u = 0; u < kSelUnroll; ++u) {
        const int m = base + u * blockDim.x + threadIdx.x;
        keys[u] = 0u;
        parts[u] = m < m1 && keyfn(m, keys[u]);
      }
#pragma unroll
      for (int u = 0; u < kSelUnroll; ++u) {
        const bool part = parts[u] && ((keys[u] & pmask) == prefix);
        const uint32_t digit = part ? ((keys[u] >> shift) & 255u) : 256u;
        const uint32_t peers = __match_any_sync(0xffffffffu, digit);
        if (part && lane == __ffs(peers) - 1) atomicAdd(&sh->hist[digit], (uint32_t)__popc(peers));
      }
    }
    cluster.sync();                               // all local histograms complete
    if (rank == 0) {
      if (threadIdx.x < 256) {
        uint32_t tot = 0u;
#pragma unroll
        for (int rk = 0; rk < kTopkCluster; ++rk) tot += cluster.map_shared_rank(sh, rk)->hist[threadIdx.x];
        sh->total[threadIdx.x] = tot;
      }
      __syncthreads();
      if (threadIdx.x < 32) {
        uint32_t loc[8], sum = 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) { loc[i] = sh->total[lane * 8 + i]; sum += loc[i]; }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += o;
        }
        const uint32_t excl = incl - sum;
        const uint32_t hit = __ballot_sync(0xffffffffu, incl >= (uint32_t)remaining);
        if (lane == __ffs(hit) - 1) {
          uint32_t cum = excl;
          int d = 0;
          for (; d < 8; ++d) {
            if (cum + loc[d] >= (uint32_t)remaining) break;
            cum += loc[d];
          }
          sh->bc[0] = (uint32_t)(lane * 8 + d);
          sh->bc[1] = cum;        // keys below the chosen digit within the current prefix
          sh->bc[2] = loc[d];
        }
      }
    }
    cluster.sync();                               // decision published; rank 0 is done with remote reads
    const uint32_t digit = sh0->bc[0];
    lt_total += (int)sh0->bc[1];
    remaining -= (int)sh0->bc[1];
    eq = (int)sh0->bc[2];
    prefix |= digit << shift;
    pmask |= 255u << shift;
    cluster.sync();                               // everyone has read bc before rank 0 rewrites it
  }
  T = prefix; n_lt = lt_total; n_eq = eq;
}

// warp-aggregated append to a list in the CTA's own shared memory
__device__ __forceinline__ void append_selected(bool sel, unsigned long long rec,
                                                unsigned long long* s_sel, int* s_count) {
  const uint32_t bal = __ballot_sync(0xffffffffu, sel);
  if (bal == 0u) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == __ffs(bal) - 1) base = atomicAdd(s_count, __popc(bal));
  base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
  if (sel) s_sel[base + __popc(bal & ((1u << lane) - 1u))] = rec;
}

__device__ __forceinline__ float sigmoid_rn(float x) {
  // correctly rounded (up to double error) sigmoid; torch's CPU sigmoid is within 1 ulp of it
  return (float)(1.0 / (1.0 + exp(-(double)x)));
}

// One 8-CTA cluster per (level, image): top-k of the level's objectness, sorted, decoded, filtered.
//   * every CTA owns a contiguous slice of the level and keeps its keys (~ordered logit) in shared memory after the
//     first pass, so the 4-5 passes of the radix select and the selection pass read DRAM / L2 once;
//   * the survivors stay in the CTA that found them: it sorts its own list (bitonic, shared memory), and after one
//     cluster barrier ranks every record against the seven other sorted lists by binary search through distributed
//     shared memory — the rank is the candidate's position in the level's descending-score order;
//   * every CTA decodes, clips and filters its own survivors.
__global__ void __cluster_dims__(kTopkCluster, 1, 1) __launch_bounds__(kTopkThreads)
rpn_topk_decode_kernel(const RpnDev g, const float* __restrict__ proposals_flat,
                       const float* __restrict__ objectness_flat,
                       const float* __restrict__ image_sizes, int kp_max, int cache_keys, float4* __restrict__ sbox,
                       float* __restrict__ cscore, uint8_t* __restrict__ alive,
                       uint32_t* __restrict__ runkey) {
  pdl_trigger();                                              // the mask kernel may be scheduled behind this grid
  extern __shared__ unsigned long long s_sel[];  // kp_max records: ~ordered(logit):32 | r:32, then the key cache
  uint32_t* s_keys = reinterpret_cast<uint32_t*>(s_sel + kp_max);
  __shared__ SelShared s_sh;
  __shared__ int s_count;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  // x = (image, cluster rank), y = level: clusters are scheduled level by level, the finest (largest) level of every image
  // first.  An SM holds one CTA of this kernel, so ~15 clusters are resident at a time (ncu: launch__cluster_max_active):
  // in (level, image)-fastest order every one of the three waves contained a 116 k-anchor level; now the first wave
  // holds the large levels and the small ones fill the slots they free.
  const int l = blockIdx.y, b = blockIdx.x / kTopkCluster;
  const int n = g.n_l[l], k = g.k_l[l], A = g.A;
  const int HW = g.H[l] * g.W[l];
  const float* __restrict__ obj =
      g.flat ? objectness_flat + (size_t)b * g.a_total + g.a_off[l] : g.obj[l] + (size_t)b * n;
  // memory index m -> reference index r (position in torchvision's (y, x, a) flattening)
  auto ref_index = [&](int m) -> uint32_t {
    if (g.flat) return (uint32_t)m;
    const int a = m / HW, rem = m - a * HW;
    return (uint32_t)(rem * A + a);
  };
  if (threadIdx.x == 0) s_count = 0;
  // this CTA's slice of the level
  const int slice = (n + kTopkCluster - 1) / kTopkCluster;
  const int m0 = min((int)rank * slice, n), m1 = min(m0 + slice, n);
  const bool cached = slice <= cache_keys;
  if (cached) {
    for (int base = m0; base < m1; base += kSelUnroll * blockDim.x) {
      float v[kSelUnroll];
#pragma unroll
      for (int u = 0; u < kSelUnroll; ++u) {
        const int m = base + u * blockDim.x + threadIdx.x;
        v[u] = m < m1 ? __ldg(obj + m) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kSelUnroll; ++u) {
        const int m = base + u * blockDim.x + threadIdx.x;
        if (m < m1) s_keys[m - m0] = ~float_ordered(v[u] + 0.f);
      }
    }
  }
  auto key_logit = [&](int m, uint32_t& key) -> bool {
    key = cached ? s_keys[m - m0] : ~float_ordered(__ldg(obj + m) + 0.f);  // smallest key = largest logit
    return true;
  };
  __syncthreads();

  if (n <= k) {
    for (int base = m0; base < m1; base += blockDim.x) {
      const int m = base + threadIdx.x;
      uint32_t key = 0u;
      const bool sel = m < m1 && key_logit(m, key);
      append_selected(sel, ((unsigned long long)key << 32) | ref_index(m < m1 ? m : 0), s_sel, &s_count);
    }
  } else {
    uint32_t T; int n_lt, n_eq;
    radix_select(key_logit, m0, m1, k, &s_sh, cluster, T, n_lt, n_eq);
    const int need = k - n_lt;  // how many of the n_eq logits equal to the threshold are taken
    uint32_t T2 = 0xffffffffu;
    if (need < n_eq) {
      // more ties than slots: take those with the smallest reference index
      auto key_tie = [&](int m, uint32_t& key2) -> bool {
        uint32_t k1; key_logit(m, k1);
        key2 = ref_index(m);
        return k1 == T;
      };
      int a_, b_;
      radix_select(key_tie, m0, m1, need, &s_sh, cluster, T2, a_, b_);
    }
    for (int base = m0; base < m1; base += kSelUnroll * blockDim.x) {
      uint32_t keys[kSelUnroll];
#pragma unroll
      for (int u = 0; u < kSelUnroll; ++u) {
        const int m = base + u * blockDim.x + threadIdx.x;
        keys[u] = 0xffffffffu;
        if (m < m1) key_logit(m, keys[u]);
      }
#pragma unroll
      for (int u = 0; u < kSelUnroll; ++u) {
        const int m = base + u * blockDim.x + threadIdx.x;
        uint32_t r = 0u;
        bool sel = false;
        if (m < m1) {
          r = ref_index(m);
          sel = keys[u] < T || (keys[u] == T && r <= T2);
        }
        append_selected(sel, ((unsigned long long)keys[u] << 32) | r, s_sel, &s_count);
      }
    }
  }
  __syncthreads();
  // own survivors, ascending: logit descending, ties by reference index
  const int mine = s_count;
  int tile = 2;
  while (tile < mine) tile <<= 1;
  for (int i = mine + threadIdx.x; i < tile; i += blockDim.x) s_sel[i] = ~0ull;
  __syncthreads();
  for (int kk = 2; kk <= tile; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int q = threadIdx.x; q < (tile >> 1); q += blockDim.x) {
        const int i = 2 * q - (q & (j - 1)), o = i + j;
        const bool asc = (i & kk) == 0;
        const unsigned long long x = s_sel[i], y = s_sel[o];
        if ((x > y) == asc) { s_sel[i] = y; s_sel[o] = x; }
      }
      __syncthreads();
    }
  }
  cluster.sync();          // every CTA's list is sorted and its count final

  const unsigned long long* lists[kTopkCluster];
  int counts[kTopkCluster];
#pragma unroll
  for (int r = 0; r < kTopkCluster; ++r) {
    lists[r] = cluster.map_shared_rank(s_sel, r);
    counts[r] = *cluster.map_shared_rank(&s_count, r);
  }
  const float img_h = image_sizes[2 * b], img_w = image_sizes[2 * b + 1];
  for (int t = threadIdx.x; t < mine; t += blockDim.x) {
    const unsigned long long rec = s_sel[t];
    // records are unique (the reference index is): position = #records of the level that sort before this one
    int lo[kTopkCluster], len[kTopkCluster];
#pragma unroll
    for (int r = 0; r < kTopkCluster; ++r) { lo[r] = 0; len[r] = (r == (int)rank) ? 0 : counts[r]; }
    for (int it = kp_max; it > 0; it >>= 1) {        // log2(kp_max) + 1 halvings, the eight searches in lockstep
#pragma unroll
      for (int r = 0; r < kTopkCluster; ++r) {
        if (len[r] > 0) {
          const int half = len[r] >> 1;
          const bool less = lists[r][lo[r] + half] < rec;
          lo[r] = less ? lo[r] + half + 1 : lo[r];
          len[r] = less ? len[r] - half - 1 : half;
        }
      }
    }
    int j = t;
#pragma unroll
    for (int r = 0; r < kTopkCluster; ++r) j += lo[r];

    const uint32_t r = (uint32_t)rec;
    const float logit = float_from_ordered(~(uint32_t)(rec >> 32));
    float x1, y1, x2, y2;
    if (g.flat) {
      const float4 p = ld_box(proposals_flat, (size_t)b * g.a_total + g.a_off[l] + r);
      x1 = p.x; y1 = p.y; x2 = p.z; y2 = p.w;
    } else {
      const int a = r % A, rem = r / A;
      const int y = rem / g.W[l], x = rem - y * g.W[l];
      const float* d = g.del[l] + ((size_t)b * A * 4 + a * 4) * HW + rem;
      const float dx = __ldg(d), dy = __ldg(d + HW);
      float dw = __ldg(d + 2 * HW), dh = __ldg(d + 3 * HW);
      const float sx = (float)(x * g.sw[l]), sy = (float)(y * g.sh[l]);
      const float ax1 = __fadd_rn(sx, g.cell[l][a][0]), ay1 = __fadd_rn(sy, g.cell[l][a][1]);
      const float ax2 = __fadd_rn(sx, g.cell[l][a][2]), ay2 = __fadd_rn(sy, g.cell[l][a][3]);
      const float w = __fsub_rn(ax2, ax1), h = __fsub_rn(ay2, ay1);
      const float cx = __fadd_rn(ax1, __fmul_rn(0.5f, w)), cy = __fadd_rn(ay1, __fmul_rn(0.5f, h));
      dw = fminf(dw, g.xform_clip);
      dh = fminf(dh, g.xform_clip);
      const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
      const float pw = __fmul_rn((float)exp((double)dw), w), ph = __fmul_rn((float)exp((double)dh), h);
      const float hw = __fmul_rn(0.5f, pw), hh = __fmul_rn(0.5f, ph);
      x1 = __fsub_rn(pcx, hw); y1 = __fsub_rn(pcy, hh);
      x2 = __fadd_rn(pcx, hw); y2 = __fadd_rn(pcy, hh);
    }
    // clip_boxes_to_image (TV ops/boxes.py:149-182)
    x1 = fminf(fmaxf(x1, 0.f), img_w); x2 = fminf(fmaxf(x2, 0.f), img_w);
    y1 = fminf(fmaxf(y1, 0.f), img_h); y2 = fminf(fmaxf(y2, 0.f), img_h);
    const float score = sigmoid_rn(logit);
    // remove_small_boxes (TV ops/boxes.py:123-146) and the score filter (TV rpn.py:285)
    const bool ok = (__fsub_rn(x2, x1) >= g.min_size) && (__fsub_rn(y2, y1) >= g.min_size) &&
                    (score >= g.score_thresh);
    const size_t pos = (size_t)b * g.k_tot + g.c_off[l] + j;
    sbox[pos] = make_float4(x1, y1, x2, y2);
    cscore[pos] = score;
    alive[pos] = ok ? 1 : 0;
    runkey[pos] = (uint32_t)(b * DGOD_MAX_LEVELS + l);
  }
  cluster.sync();          // no CTA leaves while its list may still be read
}

// Merge the per-level kept lists of an image by descending score (ties: lower level, then
// earlier candidate — the stable order of the reference's sort) and emit the first post_nms.
__global__ void __launch_bounds__(256)
rpn_merge_kernel(const RpnDev g, int post_nms_top_n, const float4* __restrict__ sbox,
                 const float* __restrict__ cscore, const int32_t* __restrict__ compact_pos,
                 const int32_t* __restrict__ run_count, float* __restrict__ out_boxes,
                 float* __restrict__ out_scores, int32_t* __restrict__ out_count) {
  pdl_wait();                                                 // launched behind the scan kernel
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  {
    int tot = 0;
    for (int l = 0; l < g.n_levels; ++l) tot += run_count[b * DGOD_MAX_LEVELS + l];
    tot = min(tot, post_nms_top_n);
    if (t == 0) out_count[b] = tot;
    // rows behind the image's last proposal are zero (the outputs are fixed-capacity)
    for (int r = tot + t; r < post_nms_top_n; r += gridDim.x * blockDim.x) {
      reinterpret_cast<float4*>(out_boxes)[(size_t)b * post_nms_top_n + r] = make_float4(0.f, 0.f, 0.f, 0.f);
      out_scores[(size_t)b * post_nms_top_n + r] = 0.f;
    }
  }
  if (t >= g.k_tot) return;
  int l = 0;
  while (l + 1 < g.n_levels && t >= g.c_off[l + 1]) ++l;
  const int j = t - g.c_off[l];
  if (j >= run_count[b * DGOD_MAX_LEVELS + l]) return;
  const size_t img0 = (size_t)b * g.k_tot;
  const int pos = compact_pos[img0 + g.c_off[l] + j];
  const float s = cscore[pos];
  // rank = own position + the number of entries of every other level that sort before this one: the binary searches of
  // the other levels run in lockstep (their dependent load pairs are in flight together: one chain of log2(k) round trips
  // instead of one per level)
  int rank = j;
  int lo[DGOD_MAX_LEVELS], len[DGOD_MAX_LEVELS];
  int longest = 0;
#pragma unroll
  for (int o = 0; o < DGOD_MAX_LEVELS; ++o) {
    lo[o] = 0;
    len[o] = (o < g.n_levels && o != l) ? run_count[b * DGOD_MAX_LEVELS + o] : 0;
    longest = max(longest, len[o]);
  }
  for (int it = longest; it > 0; it >>= 1) {             // floor(log2(longest)) + 1 halvings settle every search
#pragma unroll
    for (int o = 0; o < DGOD_MAX_LEVELS; ++o) {
      if (len[o] > 0) {                                   // scores along a level's list are non-increasing
        const int half = len[o] >> 1;
        const float v = cscore[compact_pos[img0 + g.c_off[o] + lo[o] + half]];
        const bool before = o < l ? (v >= s) : (v > s);
        lo[o] = before ? lo[o] + half + 1 : lo[o];
        len[o] = before ? len[o] - half - 1 : half;
      }
    }
  }
#pragma unroll
  for (int o = 0; o < DGOD_MAX_LEVELS; ++o) rank += lo[o];
  if (rank < post_nms_top_n) {
    reinterpret_cast<float4*>(out_boxes)[(size_t)b * post_nms_top_n + rank] = sbox[pos];
    out_scores[(size_t)b * post_nms_top_n + rank] = s;
  }
}

struct RpnBuffers {
  float4* sbox; float* cscore; uint8_t* alive; uint32_t* runkey; int32_t* compact;
  int32_t* run_count; unsigned long long* keepbits; unsigned long long* diag_cols; unsigned long long* mask;
};

static int fill_geometry(const dgod_rpn_config* cfg, RpnDev& g, int flat) {
  DGOD_REQUIRE(cfg, "rpn: cfg is null");
  DGOD_REQUIRE(cfg->n_levels >= 1 && cfg->n_levels <= DGOD_MAX_LEVELS, "rpn: n_levels out of range");
  DGOD_REQUIRE(cfg->anchors_per_loc >= 1 && cfg->anchors_per_loc <= DGOD_MAX_CELL_ANCHORS,
               "rpn: anchors_per_loc out of range");
  DGOD_REQUIRE(cfg->n_img >= 0 && cfg->pre_nms_top_n > 0 && cfg->post_nms_top_n > 0, "rpn: bad sizes");
  DGOD_REQUIRE(cfg->pre_nms_top_n <= 8192, "rpn: pre_nms_top_n > 8192 is not supported");
  g.n_levels = cfg->n_levels; g.A = cfg->anchors_per_loc; g.n_img = cfg->n_img; g.flat = flat;
  g.min_size = cfg->min_size; g.score_thresh = cfg->score_thresh; g.xform_clip = cfg->bbox_xform_clip;
  int ao = 0, co = 0;
  for (int l = 0; l < cfg->n_levels; ++l) {
    DGOD_REQUIRE(cfg->height[l] > 0 && cfg->width[l] > 0, "rpn: empty feature level");
    g.H[l] = cfg->height[l]; g.W[l] = cfg->width[l];
    g.sh[l] = cfg->stride_h[l]; g.sw[l] = cfg->stride_w[l];
    g.n_l[l] = cfg->height[l] * cfg->width[l] * cfg->anchors_per_loc;
    g.k_l[l] = g.n_l[l] < cfg->pre_nms_top_n ? g.n_l[l] : cfg->pre_nms_top_n;
    g.a_off[l] = ao; g.c_off[l] = co;
    ao += g.n_l[l]; co += g.k_l[l];
    for (int a = 0; a < cfg->anchors_per_loc; ++a)
      for (int c = 0; c < 4; ++c) g.cell[l][a][c] = cfg->cell_anchors[l][a][c];
  }
  g.a_total = ao; g.k_tot = co;
  return DGOD_OK;
}

static size_t carve_rpn(Workspace& ws, RpnBuffers& b, const RpnDev& g) {
  const size_t n_pos = (size_t)(g.n_img > 0 ? g.n_img : 1) * g.k_tot;
  int max_k = 1;
  for (int l = 0; l < g.n_levels; ++l) max_k = g.k_l[l] > max_k ? g.k_l[l] : max_k;
  b.sbox = ws.take<float4>(n_pos);
  b.cscore = ws.take<float>(n_pos);
  b.alive = ws.take<uint8_t>(n_pos);
  b.runkey = ws.take<uint32_t>(n_pos);
  b.compact = ws.take<int32_t>(n_pos);
  b.run_count = ws.take<int32_t>((size_t)(g.n_img > 0 ? g.n_img : 1) * DGOD_MAX_LEVELS);
  b.keepbits = ws.take<unsigned long long>(n_pos / 64 + 1);
  b.diag_cols = ws.take<unsigned long long>(n_pos);
  b.mask = ws.take<unsigned long long>(nms_mask_rows(n_pos) * nms_mask_row_words(max_k));
  return ws.used;
}

static int rpn_run(const dgod_rpn_config* cfg, RpnDev& g, const float* proposals_flat,
                   const float* objectness_flat, const float* image_sizes, float* out_boxes,
                   float* out_scores, int32_t* out_count, void* workspace, size_t workspace_bytes,
                   cudaStream_t st) {
  if (g.n_img == 0) return DGOD_OK;
  DGOD_REQUIRE(image_sizes && out_boxes && out_scores && out_count, "rpn: null pointer");
  Workspace ws(workspace, workspace_bytes);
  RpnBuffers b;
  carve_rpn(ws, b, g);
  if (!workspace || !ws.ok()) {
    set_error("rpn: workspace too small (%zu < %zu)", workspace_bytes, ws.used);
    return DGOD_ERR_WORKSPACE;
  }
  const int n_pos = g.n_img * g.k_tot;
  int max_k = 1;
  for (int l = 0; l < g.n_levels; ++l) max_k = g.k_l[l] > max_k ? g.k_l[l] : max_k;
  int kp = 2;
  while (kp < max_k) kp <<= 1;
  // shared memory: the CTA's survivor list (all k may come from one slice) + the key cache of its slice
  int max_slice = 1;
  for (int l = 0; l < g.n_levels; ++l) max_slice = max(max_slice, (g.n_l[l] + kTopkCluster - 1) / kTopkCluster);
  const size_t list_bytes = (size_t)kp * sizeof(unsigned long long);
  const size_t cache_cap = (size_t)(200 * 1024) - list_bytes;
  const int cache_keys = (size_t)max_slice * 4 <= cache_cap ? max_slice : 0;     // too large an image: re-read the logits
  const size_t smem = list_bytes + (size_t)cache_keys * sizeof(uint32_t);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    DGOD_CUDA(cudaFuncSetAttribute(rpn_topk_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  // no memset: every (image, level) is a run (k_l >= 1), so the scan writes every run_count the merge reads, and the
  // merge zero-fills the output rows behind each image's count
  rpn_topk_decode_kernel<<<dim3(g.n_img * kTopkCluster, g.n_levels), kTopkThreads, smem, st>>>(
      g, proposals_flat, objectness_flat, image_sizes, kp, cache_keys, b.sbox, b.cscore, b.alive, b.runkey);
  DGOD_LAUNCHED();
  int rc = launch_nms_mask(b.sbox, b.runkey, n_pos, max_k, float_round_down(cfg->nms_thresh), b.mask, b.diag_cols, st);
  if (rc) return rc;
  rc = launch_nms_scan(b.mask, b.diag_cols, b.runkey, b.alive, n_pos, max_k, nullptr, b.compact, b.run_count, st);
  if (rc) return rc;
  DGOD_CUDA(launch_pdl(rpn_merge_kernel, dim3(cdiv(g.k_tot, 256), g.n_img), dim3(256), 0, st, g, cfg->post_nms_top_n, (const float4*)b.sbox,
                       (const float*)b.cscore, (const int32_t*)b.compact, (const int32_t*)b.run_count, out_boxes, out_scores, out_count));
  DGOD_LAUNCHED();
  return DGOD_OK;
}

}  // namespace dgod

using namespace dgod;

extern "C" size_t dgod_rpn_workspace_bytes(const dgod_rpn_config* cfg) {
  RpnDev g;
  if (fill_geometry(cfg, g, 0) != DGOD_OK) return 0;
  Workspace ws(nullptr, 0);
  RpnBuffers b;
  return carve_rpn(ws, b, g);
}

extern "C" int dgod_rpn_proposals(const dgod_rpn_config* cfg, const float* const* objectness,
                                  const float* const* deltas, const float* image_sizes,
                                  float* out_boxes, float* out_scores, int32_t* out_count,
                                  void* workspace, size_t workspace_bytes, dgod_stream_t stream) {
  RpnDev g;
  int rc = fill_geometry(cfg, g, 0);
  if (rc) return rc;
  DGOD_REQUIRE(objectness && deltas, "dgod_rpn_proposals: null pointer array");
  for (int l = 0; l < g.n_levels; ++l) {
    DGOD_REQUIRE(g.n_img == 0 || (objectness[l] && deltas[l]), "dgod_rpn_proposals: null level pointer");
    g.obj[l] = objectness[l];
    g.del[l] = deltas[l];
  }
  return rpn_run(cfg, g, nullptr, nullptr, image_sizes, out_boxes, out_scores, out_count, workspace,
                 workspace_bytes, (cudaStream_t)stream);
}

extern "C" int dgod_rpn_filter(const dgod_rpn_config* cfg, const float* proposals,
                               const float* objectness, const float* image_sizes, float* out_boxes,
                               float* out_scores, int32_t* out_count, void* workspace,
                               size_t workspace_bytes, dgod_stream_t stream) {
  RpnDev g;
  int rc = fill_geometry(cfg, g, 1);
  if (rc) return rc;
  DGOD_REQUIRE(g.n_img == 0 || (proposals && objectness), "dgod_rpn_filter: null pointer");
  for (int l = 0; l < DGOD_MAX_LEVELS; ++l) { g.obj[l] = nullptr; g.del[l] = nullptr; }
  return rpn_run(cfg, g, proposals, objectness, image_sizes, out_boxes, out_scores, out_count,
                 workspace, workspace_bytes, (cudaStream_t)stream);
}
