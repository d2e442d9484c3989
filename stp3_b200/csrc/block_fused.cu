// Tail of a TemporalBlock as ONE back-to-back tcgen05 kernel (stp3/layers/temporal.py:426-489):
//
//   path 0 = causal (2,3,3) conv/BN/ReLU of mid_0      path 1 = (1,3,3) conv/BN/ReLU of mid_1      path 2 = 1x1x1 conv/BN/ReLU of x
//   out    = relu(BN(aggregation 1x1x1 of [path 0 | path 1 | path 2 | pyramid pooling])) + (projection(x) | x)
//
// (mid_0 / mid_1 = the paths' 1x1x1 entry convolutions, produced by a preceding stp3_conv_fwd launch).  The unfused form
// writes the three paths into a 128-channel concat tensor and reads it back (2 x 246 MB per block and 4-sample step) and
// needs one more pass over x for path 2 / the projection.  Here a CTA pair keeps a 16x16-pixel tile on chip:
//
//   (taps of one kernel column share ONE activation load: the box holds 8 + 2 image rows, tap j reads it shifted by j rows)
//   main(u) : up to three MMA chains accumulate the paths side by side into ONE TMEM accumulator (chain c at column
//             offset 48*c / 64: the concat happens in TMEM), each chain with its own input tensor, tap list, MMA width N
//             and K-step range
//   convert : epilogue warps read the accumulator in 8-channel pieces, add the (per-image) bias, ReLU, split to bf16 hi/lo
//             and write the compacted 128-channel operand P into shared memory (128B-swizzled K-major)
//   proj(u) : acc2 = P . W_agg  (N = 64);  acc3 = x_tile . W_projection (N = 64) when the block changes its width
//   final   : out = relu(acc2 + per-image bias) + (acc3 + per-image bias | x), hi/lo planes, per-image column sums
//
// proj(u) is issued after main(u+1), like in aspp_fused.cu.  Same precision scheme (bf16 hi/lo planes, three MMAs per
// product, fp32 accumulation in TMEM).
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace stp3 {

constexpr int kBlkThreads = 320;
constexpr int kBlkMaxChains = 3;
constexpr int kBlkMaxTaps = 32;                 // 18 + 9 + 1 path taps + the projection's
constexpr int kBlkNA = 3, kBlkNB = 4;          // the kernel is TMA-latency bound: as many activation bytes in flight as fit
constexpr int kBlkBoxRows = 10;                   // 8 image rows + 2: one activation load serves the three dy taps of a kernel column
constexpr int kBlkAStage = 2 * kBlkBoxRows * 16 * 128;
constexpr int kBlkMaxGroups = 16;
constexpr int kBlkBRows = 32;                     // weight rows per CTA and plane: chains are at most 64 wide
constexpr int kBlkBStage = 2 * kBlkBRows * 128;
constexpr int kBlkPPlane = 128 * 128;
constexpr int kAcc1Stride = 160;                  // TMEM columns between the two hidden accumulators (144 used)
constexpr int kAcc2Col = 320, kAcc3Col = 384;

struct BlkChain {
  int src;                  // 0 = mid tensor, 1 = x tensor
  int cin_off;              // first channel of the 64-channel K block read from the source
  int tap0, ntaps;
  int grp0, ngrp;           // tap groups: runs of <= 3 taps with the same (dt, dx) and consecutive dy share one activation load
  int n_mma;                // MMA width (multiple of 16)
  int tmem_col;             // column offset inside the hidden accumulator
  int ks_first, ks_end;     // UMMA_K = 16 steps that carry data
  int wblk0;                // first weight block (one per tap)
};

struct BlkParams {
  int early_trigger;       // debug switch: griddepcontrol.launch_dependents at the top of the kernel
  int n_img, T, H, W;
  int tiles_x, tiles_y, n_tiles;
  int n_chain;
  BlkChain chain[kBlkMaxChains];
  int has_res_proj;         // acc3 = x . W_projection
  BlkChain res;
  signed char tap[kBlkMaxTaps][4];        // (dt, dy, dx)
  unsigned char gstart[kBlkMaxGroups], gsize[kBlkMaxGroups];   // first tap / number of taps of every group
  int piece_col[16];        // TMEM column (inside the hidden accumulator) of the 8-channel piece pp of P, -1 = zeros
  const float* hid_bias;    // [n_img][128] bias of the hidden channels in P order
  const float* img_bias;    // [n_img][64]  aggregation bias (+ pyramid-pooling branch)
  const float* res_bias;    // [n_img][64]  projection bias (has_res_proj) or null
  const __nv_bfloat16* res_hi;            // identity residual (x planes) when !has_res_proj
  const __nv_bfloat16* res_lo;
  int res_cstride;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  int out_cstride;
  float* sum_part;          // [gridDim.x * 4][n_img][64] or null
  int wagg_blk0;            // two weight blocks (K blocks of P) of the aggregation conv
};

// true if every input element the tap group touches for this (whole, 16x16) tile is zero padding
__device__ __forceinline__ bool blk_group_is_padding(const BlkParams& p, int g, int tidx, int oy_tile, int ox0) {
  const int t0 = p.gstart[g];
  const int tt = tidx + p.tap[t0][0];
  const int ylo = oy_tile + p.tap[t0][1], yhi = oy_tile + 15 + p.tap[t0][1] + p.gsize[g] - 1;
  const int xlo = ox0 + p.tap[t0][2], xhi = ox0 + 15 + p.tap[t0][2];
  return tt < 0 || tt >= p.T || yhi < 0 || ylo >= p.H || xhi < 0 || xlo >= p.W;
}

__global__ void __launch_bounds__(kBlkThreads, 1)
block_fused_kernel(const __grid_constant__ CUtensorMap tm_m_hi, const __grid_constant__ CUtensorMap tm_m_lo,
                   const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo,
                   const __grid_constant__ CUtensorMap tm_w, const BlkParams p) {
  const uint32_t rank = ptx::cluster_ctarank();
  const int cta = (int)(blockIdx.x >> 1), n_cta = (int)(gridDim.x >> 1);
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* a_ring = smem;
  unsigned char* b_ring = a_ring + kBlkNA * kBlkAStage;
  unsigned char* p_buf = b_ring + kBlkNB * kBlkBStage;          // [kb2][hi | lo][128 rows x 128 B]
  float* s_hb = reinterpret_cast<float*>(p_buf + 4 * kBlkPPlane);   // [4 converting warps][128] hidden bias of the image
  float* s_wb = s_hb + 8 * 64;                                      // [4 finishing warps][128] img_bias (64) | res_bias (64)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_wb + 8 * 64);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kBlkNA;
  uint64_t* b_full = a_empty + kBlkNA;
  uint64_t* b_empty = b_full + kBlkNB;
  uint64_t* acc1_full = b_empty + kBlkNB;
  uint64_t* acc1_empty = acc1_full + 2;
  uint64_t* p_full = acc1_empty + 2;
  uint64_t* p_empty = p_full + 2;
  uint64_t* acc2_full = p_empty + 2;
  uint64_t* acc2_empty = acc2_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.early_trigger) ptx::griddep_launch_dependents();
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_m_hi); ptx::prefetch_tmap(&tm_m_lo); ptx::prefetch_tmap(&tm_x_hi); ptx::prefetch_tmap(&tm_x_lo);
    ptx::prefetch_tmap(&tm_w);
    for (int i = 0; i < kBlkNA; ++i) { ptx::mbar_init(&a_full[i], 2); ptx::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kBlkNB; ++i) { ptx::mbar_init(&b_full[i], 2); ptx::mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc1_full[i], 1);
      ptx::mbar_init(&acc1_empty[i], 8);          // the 4 converting warps of both CTAs
      ptx::mbar_init(&p_full[i], 8);
      ptx::mbar_init(&p_empty[i], 1);
    }
    ptx::mbar_init(acc2_full, 1);
    ptx::mbar_init(acc2_empty, 8);                // the 4 finishing warps of both CTAs
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc_pair<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ===================== TMA producer =====================
    ptx::griddep_wait();
    int as = 0, bs = 0; uint32_t aph = 0, bph = 0;
    auto load_a = [&](int src, int c, int x, int y, int t, int b) {
      ptx::mbar_wait(&a_empty[as], aph ^ 1);
      if (ptx::elect_one_sync()) {
        unsigned char* sa = a_ring + (size_t)as * kBlkAStage;
        const uint32_t bar = ptx::mapa(ptx::smem_u32(&a_full[as]), 0);
        ptx::mbar_arrive_expect_tx_cluster(bar, (uint32_t)kBlkAStage);
        ptx::tma_load_5d_pair(sa, src ? &tm_x_hi : &tm_m_hi, bar, c, x, y, t, b);
        ptx::tma_load_5d_pair(sa + kBlkAStage / 2, src ? &tm_x_lo : &tm_m_lo, bar, c, x, y, t, b);
      }
      __syncwarp();
      if (++as == kBlkNA) { as = 0; aph ^= 1; }
    };
    // weight block `blk` = [hi 128 rows][lo 128 rows]; this CTA multiplies n_mma / 2 of its 64 rows: only those are fetched
    // (boxes of 8 rows) -- with 28 taps per tile the weights, not the activations, were the larger L2 -> SM stream
    auto load_b = [&](int blk, int n_mma) {
      ptx::mbar_wait(&b_empty[bs], bph ^ 1);
      if (ptx::elect_one_sync()) {
        const uint32_t bar = ptx::mapa(ptx::smem_u32(&b_full[bs]), 0);
        unsigned char* dst = b_ring + (size_t)bs * kBlkBStage;
        const int rows = n_mma / 2;
        ptx::mbar_arrive_expect_tx_cluster(bar, (uint32_t)(2 * rows * 128));
        for (int r8 = 0; r8 < rows; r8 += 8) {
          ptx::tma_load_2d_pair(dst + r8 * 128, &tm_w, bar, 0, blk * 256 + (int)rank * 64 + r8);
          ptx::tma_load_2d_pair(dst + kBlkBRows * 128 + r8 * 128, &tm_w, bar, 0, blk * 256 + 128 + (int)rank * 64 + r8);
        }
      }
      __syncwarp();
      if (++bs == kBlkNB) { bs = 0; bph ^= 1; }
    };
    int pend_x = 0, pend_y = 0, pend_t = 0, pend_b = 0;
    bool pend = false;
    auto proj_loads = [&]() {
      load_b(p.wagg_blk0, 64); load_b(p.wagg_blk0 + 1, 64);
      if (p.has_res_proj) { load_a(1, p.res.cin_off, pend_x, pend_y, pend_t, pend_b); load_b(p.res.wblk0, 64); }
    };
    for (int tile = cta; tile < p.n_tiles; tile += n_cta) {
      const int img = tile / tiles_per_img, rem = tile % tiles_per_img;
      const int oy_tile = (rem / p.tiles_x) * 16, ox0 = (rem % p.tiles_x) * 16;
      const int oy0 = oy_tile + (int)rank * 8;
      const int bidx = img / p.T, tidx = img % p.T;
      for (int c = 0; c < p.n_chain; ++c) {
        const BlkChain& ch = p.chain[c];
        for (int g = ch.grp0; g < ch.grp0 + ch.ngrp; ++g) {
          if (blk_group_is_padding(p, g, tidx, oy_tile, ox0)) continue;
          const int t0 = p.gstart[g];
          load_a(ch.src, ch.cin_off, ox0 + p.tap[t0][2], oy0 + p.tap[t0][1], tidx + p.tap[t0][0], bidx);
          for (int j = 0; j < p.gsize[g]; ++j) load_b(ch.wblk0 + (t0 + j - ch.tap0), ch.n_mma);
        }
      }
      if (pend) proj_loads();
      pend = true; pend_x = ox0; pend_y = oy0; pend_t = tidx; pend_b = bidx;
    }
    if (pend) proj_loads();
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (leader) =====================
    int as = 0, bs = 0; uint32_t aph = 0, bph = 0;
    int buf1 = 0; uint32_t acc1_ph = 0, pph = 0, t2ph = 0;
    auto issue = [&](uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t idesc, int k0, int k1,
                     uint32_t accumulate) {
      const uint64_t da_hi = ptx::umma_desc_k_sw128(a_hi), da_lo = ptx::umma_desc_k_sw128(a_lo);
      const uint64_t db_hi = ptx::umma_desc_k_sw128(b_hi), db_lo = ptx::umma_desc_k_sw128(b_lo);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < k0 || k >= k1) continue;
        const uint64_t koff = (uint64_t)((k * 32) >> 4);
        ptx::umma_bf16_pair(tmem_d, da_hi + koff, db_hi + koff, idesc, accumulate | (uint32_t)(k > k0));
        ptx::umma_bf16_pair(tmem_d, da_hi + koff, db_lo + koff, idesc, 1);
        ptx::umma_bf16_pair(tmem_d, da_lo + koff, db_hi + koff, idesc, 1);
      }
    };
    const uint32_t idesc64 = ptx::umma_idesc_bf16(256, 64);
    auto proj = [&]() {
      ptx::mbar_wait(acc2_empty, t2ph ^ 1);        // the previous unit's output has left acc2 / acc3
      ptx::tc_fence_after();
      for (int kb2 = 0; kb2 < 2; ++kb2) {
        ptx::mbar_wait(&p_full[kb2], pph);
        ptx::mbar_wait(&b_full[bs], bph);
        ptx::tc_fence_after();
        if (ptx::elect_one_sync()) {
          const uint32_t a_hi = ptx::smem_u32(p_buf + (size_t)kb2 * 2 * kBlkPPlane);
          const uint32_t b_hi = ptx::smem_u32(b_ring + (size_t)bs * kBlkBStage);
          issue(tmem_base + kAcc2Col, a_hi, a_hi + kBlkPPlane, b_hi, b_hi + kBlkBRows * 128, idesc64, 0, 4, kb2 > 0 ? 1u : 0u);
          ptx::umma_commit_pair(&b_empty[bs]);
          ptx::umma_commit_pair(&p_empty[kb2]);
        }
        __syncwarp();
        if (++bs == kBlkNB) { bs = 0; bph ^= 1; }
      }
      pph ^= 1;
      if (p.has_res_proj) {
        ptx::mbar_wait(&a_full[as], aph);
        ptx::mbar_wait(&b_full[bs], bph);
        ptx::tc_fence_after();
        if (ptx::elect_one_sync()) {
          const uint32_t a_hi = ptx::smem_u32(a_ring + (size_t)as * kBlkAStage);
          const uint32_t b_hi = ptx::smem_u32(b_ring + (size_t)bs * kBlkBStage);
          issue(tmem_base + kAcc3Col, a_hi, a_hi + kBlkAStage / 2, b_hi, b_hi + kBlkBRows * 128, idesc64, p.res.ks_first, p.res.ks_end, 0u);
          ptx::umma_commit_pair(&b_empty[bs]);
          ptx::umma_commit_pair(&a_empty[as]);
        }
        __syncwarp();
        if (++as == kBlkNA) { as = 0; aph ^= 1; }
        if (++bs == kBlkNB) { bs = 0; bph ^= 1; }
      }
      if (ptx::elect_one_sync()) ptx::umma_commit_pair(acc2_full);
      __syncwarp();
      t2ph ^= 1;
    };
    bool pend = false;
    for (int tile = cta; tile < p.n_tiles; tile += n_cta) {
      const int img = tile / tiles_per_img, rem = tile % tiles_per_img;
      const int oy_tile = (rem / p.tiles_x) * 16, ox0 = (rem % p.tiles_x) * 16;
      const int tidx = img % p.T;
      ptx::mbar_wait(&acc1_empty[buf1], acc1_ph ^ 1);
      ptx::tc_fence_after();
      for (int c = 0; c < p.n_chain; ++c) {
        const BlkChain& ch = p.chain[c];
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf1 * kAcc1Stride + ch.tmem_col);
        const uint32_t idesc = ptx::umma_idesc_bf16(256, ch.n_mma);
        uint32_t accumulate = 0;
        for (int g = ch.grp0; g < ch.grp0 + ch.ngrp; ++g) {
          if (blk_group_is_padding(p, g, tidx, oy_tile, ox0)) continue;
          ptx::mbar_wait(&a_full[as], aph);
          ptx::tc_fence_after();
          const uint32_t a_hi0 = ptx::smem_u32(a_ring + (size_t)as * kBlkAStage);
          for (int j = 0; j < p.gsize[g]; ++j) {
            ptx::mbar_wait(&b_full[bs], bph);
            ptx::tc_fence_after();
            if (ptx::elect_one_sync()) {
              const uint32_t a_hi = a_hi0 + (uint32_t)(j * 16 * 128);          // tap j of the group: shifted by j image rows
              const uint32_t b_hi = ptx::smem_u32(b_ring + (size_t)bs * kBlkBStage);
              issue(tmem_d, a_hi, a_hi + kBlkAStage / 2, b_hi, b_hi + kBlkBRows * 128, idesc, ch.ks_first, ch.ks_end, accumulate);
              ptx::umma_commit_pair(&b_empty[bs]);
            }
            __syncwarp();
            accumulate = 1;
            if (++bs == kBlkNB) { bs = 0; bph ^= 1; }
          }
          if (ptx::elect_one_sync()) ptx::umma_commit_pair(&a_empty[as]);
          __syncwarp();
          if (++as == kBlkNA) { as = 0; aph ^= 1; }
        }
      }
      if (ptx::elect_one_sync()) ptx::umma_commit_pair(&acc1_full[buf1]);
      __syncwarp();
      if (++buf1 == 2) { buf1 = 0; acc1_ph ^= 1; }
      if (pend) proj();
      pend = true;
    }
    if (pend) proj();
  } else if (warp >= 2) {
    // ===================== epilogue: two independent warp groups =====================
    // warps 2..5 CONVERT (hidden accumulator -> operand P), warps 6..9 FINISH (acc2 / acc3 -> output): the conversion of
    // tile u+1 overlaps the output of tile u instead of queueing behind it in the same warps.  Each group has one warp per
    // TMEM lane quarter (warp id % 4).
    const int e = warp - 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    ptx::griddep_wait();
    if (e < 4) {
      // ---------- convert: 16 pieces of 8 channels per thread (row r of both K blocks of P)
      int buf1 = 0; uint32_t acc1_ph = 0, pph = 0;
      const uint32_t sw = (uint32_t)(r & 7);
      float* hb = s_hb + e * 128;
      for (int tile = cta; tile < p.n_tiles; tile += n_cta) {
        const int img = tile / tiles_per_img;
        {                                          // hidden-bias row of this image (written by a preceding small kernel)
          const volatile float* hsrc = p.hid_bias + (size_t)img * 128;
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) hb[lane + 32 * i] = hsrc[lane + 32 * i];
          __syncwarp();
        }
        ptx::mbar_wait(&acc1_full[buf1], acc1_ph);
        ptx::mbar_wait(&p_empty[0], pph ^ 1);      // the projection of the previous tile has read P
        ptx::mbar_wait(&p_empty[1], pph ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf1 * kAcc1Stride);
#pragma unroll 1
        for (int pp = 0; pp < 16; ++pp) {
          const int col = p.piece_col[pp];
          uint32_t hw[4] = {0u, 0u, 0u, 0u}, lw[4] = {0u, 0u, 0u, 0u};
          if (col >= 0) {                          // warp-uniform
            uint32_t acc[8];
            ptx::tmem_ld_32x32b_x8(tmem_acc + col, acc);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float x0 = fmaxf(__uint_as_float(acc[2 * i]) + hb[pp * 8 + 2 * i], 0.f);
              const float x1 = fmaxf(__uint_as_float(acc[2 * i + 1]) + hb[pp * 8 + 2 * i + 1], 0.f);
              const uint32_t h = ptx::pack_bf16x2(x0, x1);
              hw[i] = h;
              lw[i] = ptx::pack_bf16x2(x0 - __uint_as_float(h << 16), x1 - __uint_as_float(h & 0xFFFF0000u));
            }
          }
          unsigned char* dst = p_buf + (size_t)(pp >> 3) * 2 * kBlkPPlane + (size_t)r * 128 + ((((uint32_t)pp & 7u) ^ sw) << 4);
          *reinterpret_cast<uint4*>(dst) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(dst + kBlkPPlane) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        ptx::tc_fence_before();
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&acc1_empty[buf1]), 0));
          ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&p_full[0]), 0));
          ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&p_full[1]), 0));
        }
        pph ^= 1;
        if (++buf1 == 2) { buf1 = 0; acc1_ph ^= 1; }
      }
    } else {
      // ---------- finish: out = relu(acc2 + bias) + residual, all 64 channels of row r, per-image column sums
      uint32_t t2ph = 0;
      float* wb = s_wb + (e - 4) * 128;            // img_bias (64) | res_bias (64)
      float sacc[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) sacc[i] = 0.f;
      int sum_img = -1;
      auto flush_sums = [&](int img_) {            // per column: sum over the 32 lanes (once per image and warp)
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          float v = sacc[i];
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
          if (lane == (i & 31)) p.sum_part[(((size_t)blockIdx.x * 4 + q) * p.n_img + img_) * 64 + i] = v;
          sacc[i] = 0.f;
        }
      };
      for (int tile = cta; tile < p.n_tiles; tile += n_cta) {
        const int img = tile / tiles_per_img, rem = tile % tiles_per_img;
        const int oy = (rem / p.tiles_x) * 16 + (int)rank * 8 + (r >> 4), ox = (rem % p.tiles_x) * 16 + (r & 15);
        if (p.sum_part && img != sum_img) {        // tiles come in image order
          if (sum_img >= 0) flush_sums(sum_img);
          sum_img = img;
        }
        {
          const volatile float* isrc = p.img_bias + (size_t)img * 64;
          __syncwarp();
          wb[lane] = isrc[lane]; wb[lane + 32] = isrc[lane + 32];
          if (p.res_bias) {
            const volatile float* rsrc = p.res_bias + (size_t)img * 64;
            wb[64 + lane] = rsrc[lane]; wb[96 + lane] = rsrc[lane + 32];
          }
          __syncwarp();
        }
        const bool valid = oy < p.H && ox < p.W;
        const size_t pix = ((size_t)img * p.H + oy) * p.W + ox;
        ptx::mbar_wait(acc2_full, t2ph);
        ptx::tc_fence_after();
        t2ph ^= 1;
        const uint32_t tmem_2 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)kAcc2Col;
        const uint32_t tmem_3 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)kAcc3Col;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t acc[16], acc3[16], rh[8], rl[8];
          if (!p.has_res_proj && valid) {          // identity residual: the block's input planes
            const size_t off = pix * p.res_cstride + j * 16;
            ptx::ld_global_v8(p.res_hi + off, rh);
            ptx::ld_global_v8(p.res_lo + off, rl);
          }
          ptx::tmem_ld_32x32b_x16(tmem_2 + j * 16, acc);
          if (p.has_res_proj) ptx::tmem_ld_32x32b_x16(tmem_3 + j * 16, acc3);
          ptx::tmem_ld_wait();
          if (valid) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(__uint_as_float(acc[i]) + wb[j * 16 + i], 0.f);
            if (p.has_res_proj) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += __uint_as_float(acc3[i]) + wb[64 + j * 16 + i];
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                v[2 * i] += __uint_as_float(rh[i] << 16) + __uint_as_float(rl[i] << 16);
                v[2 * i + 1] += __uint_as_float(rh[i] & 0xFFFF0000u) + __uint_as_float(rl[i] & 0xFFFF0000u);
              }
            }
            uint32_t hw[8], lw[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t h = ptx::pack_bf16x2(v[2 * i], v[2 * i + 1]);
              hw[i] = h;
              lw[i] = ptx::pack_bf16x2(v[2 * i] - __uint_as_float(h << 16), v[2 * i + 1] - __uint_as_float(h & 0xFFFF0000u));
            }
            const size_t off = pix * p.out_cstride + j * 16;
            ptx::st_global_v8(p.out_hi + off, hw);
            ptx::st_global_v8(p.out_lo + off, lw);
            if (p.sum_part) {
#pragma unroll
              for (int i = 0; i < 16; ++i) sacc[j * 16 + i] += v[i];
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(acc2_empty), 0));
      }
      if (p.sum_part && sum_img >= 0) flush_sums(sum_img);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 1) ptx::tmem_dealloc_pair<512>(tmem_base);
}

typedef CUresult (*PFN_tmapEncodeTiledB)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                         const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiledB blk_encode_fn() {
  static PFN_tmapEncodeTiledB fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiledB>(ptr);
  }
  return fn;
}

static int blk_num_sms() {
  static const int n = [] {
    int dev = 0, v = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
  }();
  return n;
}

}  // namespace stp3

using namespace stp3;

extern "C" size_t stp3_block_fused_scratch_bytes(int n_img) {
  return (size_t)blk_num_sms() * 4 * (size_t)(n_img > 0 ? n_img : 0) * 64 * sizeof(float);
}

extern "C" int stp3_block_fused_fwd(const stp3_block_desc* d, const void* mid_hi, const void* mid_lo, const void* x_hi,
                                    const void* x_lo, const void* w, const float* hid_bias, const float* img_bias,
                                    const float* res_bias, void* y_hi, void* y_lo, float* col_sums, void* scratch,
                                    size_t scratch_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  STP3_CHECK_ARG(d && mid_hi && mid_lo && x_hi && x_lo && w && hid_bias && img_bias && y_hi && y_lo,
                 "stp3_block_fused_fwd: null pointer argument");
  STP3_CHECK_ARG(d->B > 0 && d->T > 0 && d->H > 0 && d->W > 0, "non-positive dimension");
  STP3_CHECK_ARG(d->mid_cstride % 64 == 0 && d->x_cstride % 64 == 0 && d->out_cstride % 16 == 0 && d->out_cstride >= 64,
                 "channel strides: mid / x multiples of 64, out a multiple of 16 and >= 64");
  STP3_CHECK_ARG(d->n_chain >= 1 && d->n_chain <= kBlkMaxChains, "1 .. 3 path chains");
  STP3_CHECK_ARG((d->has_res_proj != 0) == (res_bias != nullptr), "res_bias is given exactly when the block has a projection");
  auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; };
  STP3_CHECK_ARG(al32(y_hi) && al32(y_lo) && al32(x_hi) && al32(x_lo), "planes must be 32-byte aligned");
  PFN_tmapEncodeTiledB enc = blk_encode_fn();
  if (!enc) return set_error(STP3_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");

  BlkParams p;
  p.n_img = d->B * d->T; p.T = d->T; p.H = d->H; p.W = d->W;
  p.tiles_x = ceil_div(d->W, 16); p.tiles_y = ceil_div(d->H, 16);
  const long long nt = (long long)p.n_img * p.tiles_x * p.tiles_y;
  STP3_CHECK_ARG(nt < (1ll << 31), "grid too large");
  p.n_tiles = (int)nt;
  p.n_chain = d->n_chain;
  int t = 0, blk = 0, n_grp = 0;
  auto fill = [&](BlkChain& c, const stp3_block_chain& s) -> int {
    c.src = s.src; c.cin_off = s.cin_off; c.tap0 = t; c.ntaps = s.n_taps; c.n_mma = s.n_mma; c.tmem_col = s.tmem_col;
    c.ks_first = s.k_lo / 16; c.ks_end = s.k_hi / 16; c.wblk0 = blk;
    if (!(s.n_taps >= 1 && s.n_taps <= 18 && s.n_mma % 16 == 0 && s.n_mma >= 16 && s.n_mma <= 64 && s.tmem_col % 16 == 0 &&
          s.tmem_col + s.n_mma <= 144 && s.cin_off % 64 == 0 && s.k_lo % 16 == 0 && s.k_hi % 16 == 0 && s.k_lo < s.k_hi &&
          s.k_hi <= 64 && (s.src == 0 || s.src == 1)))
      return set_error(STP3_EINVAL, "stp3_block_fused_fwd: bad chain descriptor");
    bool centre = false;
    for (int i = 0; i < s.n_taps; ++i, ++t) {
      if (t >= kBlkMaxTaps) return set_error(STP3_EINVAL, "too many taps");
      p.tap[t][0] = s.taps[i][0]; p.tap[t][1] = s.taps[i][1]; p.tap[t][2] = s.taps[i][2]; p.tap[t][3] = 0;
      centre |= s.taps[i][0] == 0 && s.taps[i][1] == 0 && s.taps[i][2] == 0;
    }
    if (!centre) return set_error(STP3_EINVAL, "every chain needs its centre tap");
    // tap groups: consecutive taps with the same (dt, dx) and dy advancing by one
    c.grp0 = n_grp;
    for (int i = c.tap0; i < c.tap0 + c.ntaps;) {
      int g = 1;
      while (g < 3 && i + g < c.tap0 + c.ntaps && p.tap[i + g][0] == p.tap[i][0] && p.tap[i + g][2] == p.tap[i][2] &&
             p.tap[i + g][1] == p.tap[i][1] + g)
        ++g;
      if (n_grp >= kBlkMaxGroups) return set_error(STP3_EINVAL, "too many tap groups");
      p.gstart[n_grp] = (unsigned char)i; p.gsize[n_grp] = (unsigned char)g; ++n_grp;
      i += g;
    }
    c.ngrp = n_grp - c.grp0;
    blk += s.n_taps;
    return STP3_OK;
  };
  for (int c = 0; c < d->n_chain; ++c) { const int rc = fill(p.chain[c], d->chain[c]); if (rc) return rc; }
  p.wagg_blk0 = blk; blk += 2;
  p.has_res_proj = d->has_res_proj ? 1 : 0;
  if (p.has_res_proj) {
    STP3_CHECK_ARG(t < kBlkMaxTaps, "too many taps");
    stp3_block_chain rs = d->res;
    STP3_CHECK_ARG(rs.n_taps == 1 && rs.src == 1 && rs.n_mma == 64, "the projection chain is a 1x1 on x with 64 outputs");
    const int rc = fill(p.res, rs); if (rc) return rc;
  } else {
    p.res = p.chain[0];
  }
  for (int i = 0; i < 16; ++i) {
    STP3_CHECK_ARG(d->piece_col[i] < 0 || (d->piece_col[i] % 8 == 0 && d->piece_col[i] + 8 <= 144), "bad piece column");
    p.piece_col[i] = d->piece_col[i];
  }
  p.hid_bias = hid_bias; p.img_bias = img_bias; p.res_bias = res_bias;
  p.res_hi = static_cast<const __nv_bfloat16*>(x_hi); p.res_lo = static_cast<const __nv_bfloat16*>(x_lo);
  p.res_cstride = d->x_cstride;
  p.out_hi = static_cast<__nv_bfloat16*>(y_hi); p.out_lo = static_cast<__nv_bfloat16*>(y_lo); p.out_cstride = d->out_cstride;

  CUtensorMap tm[4], tm_w;
  const void* planes[4] = {mid_hi, mid_lo, x_hi, x_lo};
  for (int i = 0; i < 4; ++i) {
    const int cs = i < 2 ? d->mid_cstride : d->x_cstride;
    const cuuint64_t dims[5] = {(cuuint64_t)cs, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->T, (cuuint64_t)d->B};
    const cuuint64_t strides[4] = {(cuuint64_t)cs * 2, (cuuint64_t)d->W * cs * 2, (cuuint64_t)d->H * d->W * cs * 2,
                                   (cuuint64_t)d->T * d->H * d->W * cs * 2};
    const cuuint32_t box[5] = {64, 16, (cuuint32_t)kBlkBoxRows, 1, 1};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tm[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(planes[i]), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(STP3_ECUDA, "cuTensorMapEncodeTiled(activation %d) failed: %d", i, (int)r);
  }
  {
    const cuuint64_t wd[2] = {64, (cuuint64_t)blk * 256};
    const cuuint64_t ws[1] = {128};
    const cuuint32_t wb[2] = {64, 8};
    const cuuint32_t we[2] = {1, 1};
    CUresult r = enc(&tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), wd, ws, wb, we,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(STP3_ECUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  }
  const size_t smem_bytes = 1024 + (size_t)kBlkNA * kBlkAStage + (size_t)kBlkNB * kBlkBStage + 4 * (size_t)kBlkPPlane +
                            2 * 8 * 64 * sizeof(float) + 32 * 8 + 16;
  static thread_local int attr_dev = -1, occ_val = 0;           // once per device: these calls cost microseconds per eager launch
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  const bool first_call = attr_dev != cur_dev;
  if (first_call) STP3_CUDA_OK(cudaFuncSetAttribute(block_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  const int num_sms = blk_num_sms();
  cudaLaunchConfig_t cfg = {};
  unsigned pairs = (unsigned)(nt < num_sms / 2 ? nt : num_sms / 2);
  cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(kBlkThreads);
  cfg.dynamicSmemBytes = smem_bytes; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int max_clusters = occ_val;
  if (first_call) {
    STP3_CUDA_OK(cudaOccupancyMaxActiveClusters(&max_clusters, block_fused_kernel, &cfg));
    occ_val = max_clusters; attr_dev = cur_dev;
  }
  if (max_clusters < 1) return set_error(STP3_EUNSUPPORTED, "no CTA pair fits on this device");
  if (cfg.gridDim.x > 2u * (unsigned)max_clusters) cfg.gridDim.x = 2u * (unsigned)max_clusters;
  p.sum_part = nullptr;
  if (col_sums) {
    const size_t need = (size_t)cfg.gridDim.x * 4 * p.n_img * 64 * sizeof(float);
    STP3_CHECK_ARG(scratch && scratch_bytes >= need, "col_sums: scratch missing or smaller than stp3_block_fused_scratch_bytes()");
    p.sum_part = static_cast<float*>(scratch);
    STP3_CUDA_OK(cudaMemsetAsync(p.sum_part, 0, need, stream));
  }
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.numAttrs = stp3_pdl_enabled("STP3_FUSED_PDL") ? 2 : 1;
  {
    // No early griddepcontrol.launch_dependents from the back-to-back kernels: with it, a chain block_fused -> col_sum_reduce
    // -> pool_bias -> aspp_fused in flight at once stopped making progress about once in 400 .. 2000 replayed steps
    // (tools/hang_probe.py; 20000 replays are clean without it and with programmatic launch off altogether).  The
    // dependents are released when the grid completes; this kernel itself still starts early behind its predecessor.
    static const bool early = [] { const char* e = getenv("STP3_FUSED_EARLY_TRIGGER"); return e && atoi(e) != 0; }();
    p.early_trigger = early ? 1 : 0;
  }
  STP3_CUDA_OK(cudaLaunchKernelEx(&cfg, block_fused_kernel, tm[0], tm[1], tm[2], tm[3], tm_w, p));
  STP3_CUDA_OK(cudaGetLastError());
  if (col_sums) return launch_col_sum_reduce(p.sum_part, (int)cfg.gridDim.x * 4, p.n_img, col_sums, stream);
  return STP3_OK;
}
