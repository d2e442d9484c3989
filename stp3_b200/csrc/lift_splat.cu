// Fused lift-splat for sm_100a.
//
//   K1  lift_splat_scatter_kernel : one CTA per (b, t, camera, tile of TW image columns)
//         phase A  depth logits and context features of the tile -> shared memory (coalesced tile loads),
//                  softmax over D in place (probabilities never leave the SM; the 371 MB/sample outer product of
//                  the reference, stp3.py:216, is never materialised)
//         phase B  frustum point -> ego frame -> sequential ego-motion warp -> voxel rank, with the reference's
//                  exact fp32 operation order (no FMA, IEEE division, truncation); rank staged beside the prob
//         phase C  outer product + pooling: a warp owns one image column; lanes own channel pairs and keep the
//                  column's context features in registers; walking the column for every depth bin is a segmented
//                  reduction over equal pillar ranks (all rows of a level camera fall into the same pillar),
//                  flushed with one coalesced red.global.add.v2.f32 per lane per segment into a channels-last
//                  (b,t,pillar,C) fp32 grid, plus one byte in a per-pillar occupancy map
//   K2  bev_finalize_kernel : temporal discount recurrence out[t] = out[t-1]*discount + grid[t] and the
//         channels-last -> (C,X,Y) transpose through shared memory.  Reads (and re-zeroes) only the occupied
//         pillars, so the scatter grid is left clean for the next call and never needs a memset.
//
// Reference semantics: /root/reference/stp3/models/stp3.py:186-301, stp3/utils/geometry.py:299-318.
#include <cstdint>

#include "common.cuh"

namespace stp3 {

constexpr int kMaxFrames = 8;      // receptive field S supported by the in-kernel pose chain
constexpr int kScatterThreads = 256;
constexpr int kHChunk = 32;        // image rows whose features a lane keeps in registers at once
constexpr int kCChunk = 64;        // channels staged in shared memory at once (2 per lane)

struct LiftSplatParams {
  const float* feat;
  const float* depth;
  const float* cam_M;
  const float* cam_t;
  const float* ego_R;
  const float* ego_t;
  const float* xs;
  const float* ys;
  const float* ds;
  float off[3];
  float res[3];
  int nx, ny, nz;
  int B, S, N, D, Hf, Wf, C;
  int feat_nhwc;
  int use_depth;
  int TW;        // image columns per CTA
  int tiles_w;   // ceil(Wf / TW)
  int32_t* ranks_out;
  float* grid;          // (B,S,nx*ny*nz,C) fp32, all-zero on entry
  unsigned char* occ;   // (B,S,nx*ny*nz) bytes, all-zero on entry; set to 1 where the grid was written
};

// ((m0*x + m1*y) + m2*z) + t, every product and sum rounded to fp32 separately: bit-identical to the reference's
// CPU batched 3x3 matmul followed by `+= translation` (stp3.py:197-198, 273-277).
__device__ __forceinline__ void affine_exact(const float* __restrict__ m, const float* __restrict__ t,
                                             float& x, float& y, float& z) {
  const float ox = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], x), __fmul_rn(m[1], y)), __fmul_rn(m[2], z)), t[0]);
  const float oy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[3], x), __fmul_rn(m[4], y)), __fmul_rn(m[5], z)), t[1]);
  const float oz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[6], x), __fmul_rn(m[7], y)), __fmul_rn(m[8], z)), t[2]);
  x = ox; y = oy; z = oz;
}

__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// shared-memory row of channel (c0 + cl) in the feature tile: the two channels of a lane (2*lane, 2*lane+1) live in
// rows lane and lane+32, so that with an odd row stride the register fill of phase C is bank-conflict free
__device__ __forceinline__ int feat_row(int cl) { return ((cl & 1) << 5) | (cl >> 1); }

__global__ void __launch_bounds__(kScatterThreads, 2)
lift_splat_scatter_kernel(const LiftSplatParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int D = p.D, Hf = p.Hf, Wf = p.Wf, C = p.C, TW = p.TW;
  const int npix = Hf * TW;
  const int npts = D * npix;
  const int fstride = npix | 1;                                    // odd row stride of the feature tile
  int2* s_pt = reinterpret_cast<int2*>(smem_raw);                  // [D][Hf][TW] {rank, prob bits}
  float* s_feat = reinterpret_cast<float*>(s_pt + npts);           // [kCChunk rows][fstride]
  float* s_mat = s_feat + kCChunk * fstride;                       // camera 12 + (kMaxFrames-1) * 12 pose floats
  float* s_ys = s_mat + 12 * kMaxFrames;                           // [Hf]
  float* s_ds = s_ys + Hf;                                         // [D]
  float* s_red = s_ds + D;                                         // [blockDim] softmax partials

  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  int blk = blockIdx.x;
  const int tile = blk % p.tiles_w; blk /= p.tiles_w;
  const int n = blk % p.N; blk /= p.N;
  const int t = blk % p.S;
  const int b = blk / p.S;
  const int w0 = tile * TW;
  const int img = (b * p.S + t) * p.N + n;       // camera-image index
  const int bt = b * p.S + t;
  const int n_chain = p.S - 1 - t;               // poses t .. S-2 applied in order (stp3.py:270-277)

  // ---- small per-CTA constants
  if (tid < 9) s_mat[tid] = p.cam_M[img * 9 + tid];
  if (tid >= 9 && tid < 12) s_mat[tid] = p.cam_t[img * 3 + tid - 9];
  for (int i = tid; i < n_chain * 12; i += nthr) {
    const int k = i / 12, e = i % 12;
    const int src = b * p.S + t + k;
    s_mat[12 + i] = e < 9 ? p.ego_R[src * 9 + e] : p.ego_t[src * 3 + e - 9];
  }
  for (int i = tid; i < Hf; i += nthr) s_ys[i] = p.ys[i];
  for (int i = tid; i < D; i += nthr) s_ds[i] = p.ds[i];

  // ---- phase A: logits tile -> smem
  if (p.use_depth) {
    const float* dsrc = p.depth + (size_t)img * D * Hf * Wf;
    for (int i = tid; i < npts; i += nthr) {
      const int wl = i % TW;
      const int dh = i / TW;                     // d*Hf + h
      const int w = w0 + wl;
      const float v = w < Wf ? __ldg(dsrc + (size_t)dh * Wf + w) : 0.f;
      s_pt[i].y = __float_as_int(v);
    }
  }
  __syncthreads();
  if (p.use_depth) {
    // softmax over D (stp3.py:215): `parts` threads share one pixel, each owning a slice of the depth axis
    const int npix_r = (npix + 31) & ~31;
    const int parts = nthr / npix_r;
    if (parts >= 1) {
      const int pix = tid % npix_r, part = tid / npix_r;
      const bool act = pix < npix && part < parts;
      const int dchunk = (D + parts - 1) / parts;
      const int da = part * dchunk, db = min(D, da + dchunk);
      float mx = -INFINITY;
      if (act) for (int d = da; d < db; ++d) mx = fmaxf(mx, __int_as_float(s_pt[d * npix + pix].y));
      s_red[tid] = mx;
      __syncthreads();
      if (act) for (int q = 0; q < parts; ++q) mx = fmaxf(mx, s_red[q * npix_r + pix]);
      __syncthreads();
      float sum = 0.f;
      if (act) for (int d = da; d < db; ++d) {
        const float e = __expf(__int_as_float(s_pt[d * npix + pix].y) - mx);
        s_pt[d * npix + pix].y = __float_as_int(e);
        sum += e;
      }
      s_red[tid] = sum;
      __syncthreads();
      if (act) {
        float tot = 0.f;
        for (int q = 0; q < parts; ++q) tot += s_red[q * npix_r + pix];
        const float inv = __frcp_rn(tot);
        for (int d = da; d < db; ++d)
          s_pt[d * npix + pix].y = __float_as_int(__int_as_float(s_pt[d * npix + pix].y) * inv);
      }
    } else {  // very tall tiles: one thread per pixel, several pixels per thread
      for (int pix = tid; pix < npix; pix += nthr) {
        float mx = -INFINITY;
        for (int d = 0; d < D; ++d) mx = fmaxf(mx, __int_as_float(s_pt[d * npix + pix].y));
        float sum = 0.f;
        for (int d = 0; d < D; ++d) {
          const float e = __expf(__int_as_float(s_pt[d * npix + pix].y) - mx);
          s_pt[d * npix + pix].y = __float_as_int(e);
          sum += e;
        }
        const float inv = __frcp_rn(sum);
        for (int d = 0; d < D; ++d)
          s_pt[d * npix + pix].y = __float_as_int(__int_as_float(s_pt[d * npix + pix].y) * inv);
      }
    }
  } else {
    for (int i = tid; i < npts; i += nthr) s_pt[i].y = __float_as_int(1.0f);  // stp3.py:218
  }

  // ---- phase B: voxel rank of every point of the tile (bit-exact with the reference CPU path)
  {
    const float offx = p.off[0], offy = p.off[1], offz = p.off[2];
    const float resx = p.res[0], resy = p.res[1], resz = p.res[2];
    const float fnx = (float)p.nx, fny = (float)p.ny, fnz = (float)p.nz;
    for (int i = tid; i < npts; i += nthr) {
      const int wl = i % TW;
      const int dh = i / TW;
      const int h = dh % Hf;
      const int d = dh / Hf;
      const int w = w0 + wl;
      int rank = -1;
      if (w < Wf) {
        const float dep = s_ds[d];
        float x = __fmul_rn(__ldg(p.xs + w), dep);     // stp3.py:195: (u*d, v*d, d)
        float y = __fmul_rn(s_ys[h], dep);
        float z = dep;
        affine_exact(s_mat, s_mat + 9, x, y, z);       // stp3.py:196-198
        for (int k = 0; k < n_chain; ++k)              // stp3.py:270-277, sequential, rounded every step
          affine_exact(s_mat + 12 + 12 * k, s_mat + 12 + 12 * k + 9, x, y, z);
        // stp3.py:287-289: ((p - (start - res/2)) / res).long()  -- true division, truncation toward zero;
        // trunc(q) in [0, n)  <=>  -1 < q < n  (this keeps the (-1,0) band in cell 0 exactly like .long()).
        const float qx = __fdiv_rn(__fsub_rn(x, offx), resx);
        const float qy = __fdiv_rn(__fsub_rn(y, offy), resy);
        const float qz = __fdiv_rn(__fsub_rn(z, offz), resz);
        const bool keep = (qx > -1.f) && (qx < fnx) && (qy > -1.f) && (qy < fny) && (qz > -1.f) && (qz < fnz);
        if (keep) {
          const int ix = (int)qx, iy = (int)qy, iz = (int)qz;     // cvt.rzi
          rank = ix * (p.ny * p.nz) + iy * p.nz + iz;              // stp3.py:251-255
        }
        if (p.ranks_out) p.ranks_out[((size_t)img * D * Hf + dh) * Wf + w] = rank;
      }
      s_pt[i].x = rank;
    }
  }

  // ---- phase C: outer product + segmented pooling.  work item = (column, depth slice); lanes = channel pairs
  const int warp = tid >> 5, lane = tid & 31;
  const int nwarps = nthr >> 5;
  const int dsplit = max(1, nwarps / TW);
  const int dper = (D + dsplit - 1) / dsplit;
  const size_t nvox = (size_t)p.nx * p.ny * p.nz;
  float* gbase = p.grid + (size_t)bt * nvox * C;
  unsigned char* obase = p.occ + (size_t)bt * nvox;
  const bool vec_ok = (C % 2) == 0;
  for (int c0 = 0; c0 < C; c0 += kCChunk) {
    // stage the context features of channels [c0, c0+64) of this tile (stp3.py:216 reads them D times)
    __syncthreads();
    if (p.feat_nhwc) {
      for (int i = tid; i < kCChunk * npix; i += nthr) {
        const int cl = i % kCChunk;
        const int px = i / kCChunk;              // h*TW + wl
        const int wl = px % TW, h = px / TW;
        const int w = w0 + wl, c = c0 + cl;
        float v = 0.f;
        if (w < Wf && c < C) v = __ldg(p.feat + (((size_t)img * Hf + h) * Wf + w) * C + c);
        s_feat[feat_row(cl) * fstride + px] = v;
      }
    } else {
      for (int i = tid; i < kCChunk * npix; i += nthr) {
        const int px = i % npix;
        const int cl = i / npix;
        const int wl = px % TW, h = px / TW;
        const int w = w0 + wl, c = c0 + cl;
        float v = 0.f;
        if (w < Wf && c < C) v = __ldg(p.feat + (((size_t)img * C + c) * Hf + h) * Wf + w);
        s_feat[feat_row(cl) * fstride + px] = v;
      }
    }
    __syncthreads();
    const int c = c0 + 2 * lane;
    for (int item = warp; item < TW * dsplit; item += nwarps) {
      const int wl = item % TW;
      if (w0 + wl >= Wf) continue;
      const int d0 = (item / TW) * dper;
      const int d1 = min(D, d0 + dper);
      for (int h0 = 0; h0 < Hf; h0 += kHChunk) {
        float f0[kHChunk], f1[kHChunk];
#pragma unroll
        for (int j = 0; j < kHChunk; ++j) {
          const int h = h0 + j;
          f0[j] = 0.f; f1[j] = 0.f;
          if (h < Hf) {
            f0[j] = s_feat[lane * fstride + h * TW + wl];
            f1[j] = s_feat[(lane + 32) * fstride + h * TW + wl];
          }
        }
        int cur = -1;
        float a0 = 0.f, a1 = 0.f;
        for (int d = d0; d < d1; ++d) {
          const int2* row = s_pt + (size_t)(d * Hf + h0) * TW + wl;
#pragma unroll
          for (int j = 0; j < kHChunk; ++j) {
            if (h0 + j < Hf) {
              const int2 v = row[j * TW];
              if (v.x != cur) {                      // warp-uniform: segment boundary
                if (cur >= 0) {
                  if (c < C) {
                    float* dst = gbase + (size_t)cur * C + c;
                    if (vec_ok) red_add_v2(dst, a0, a1);
                    else { atomicAdd(dst, a0); if (c + 1 < C) atomicAdd(dst + 1, a1); }
                  }
                  if (lane == 0 && c0 == 0) obase[cur] = 1;
                }
                cur = v.x; a0 = 0.f; a1 = 0.f;
              }
              const float pr = __int_as_float(v.y);
              a0 = fmaf(pr, f0[j], a0);
              a1 = fmaf(pr, f1[j], a1);
            }
          }
        }
        if (cur >= 0) {
          if (c < C) {
            float* dst = gbase + (size_t)cur * C + c;
            if (vec_ok) red_add_v2(dst, a0, a1);
            else { atomicAdd(dst, a0); if (c + 1 < C) atomicAdd(dst + 1, a1); }
          }
          if (lane == 0 && c0 == 0) obase[cur] = 1;
        }
      }
    }
  }
}

// out[b,t] = out[b,t-1]*discount + grid[b,t]  (stp3.py:296, separate fp32 mul and add like the reference's
// `bev_feature * discount + tmp`), written either channels-last or as (C, X*Y) through a 32x32 shared-memory
// transpose.  One CTA = 32 consecutive pillars of one sample, all channels, all frames.  Only occupied
// (frame, pillar) rows of the grid are read; they are zeroed again and their occupancy byte cleared, which
// leaves the workspace clean for the next call.  pool_sum (B,S,C) += sum over cells (optional).
__global__ void __launch_bounds__(256)
bev_finalize_kernel(float* __restrict__ grid, unsigned char* __restrict__ occ, float* __restrict__ out,
                    float* __restrict__ pool_sum, int S, int C, int nvox, float discount, int out_nhwc) {
  __shared__ float tile[32][33];
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;     // (32, 8); warp == ty
  // occupancy bits of this warp's 4 pillars for every frame: bit (t*4 + k)
  unsigned occ_bits = 0;
  for (int t = 0; t < S; ++t) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int pcell = p0 + ty + 8 * k;
      if (pcell < nvox && occ[((size_t)b * S + t) * nvox + pcell]) occ_bits |= 1u << (t * 4 + k);
    }
  }
  __syncwarp();
  if (tx == 0) {
    for (int t = 0; t < S; ++t)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (occ_bits & (1u << (t * 4 + k))) occ[((size_t)b * S + t) * nvox + p0 + ty + 8 * k] = 0;
  }
  const bool use_tile = !out_nhwc || pool_sum != nullptr;
  for (int c0 = 0; c0 < C; c0 += 32) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int c = c0 + tx;
    for (int t = 0; t < S; ++t) {
      const size_t bt = (size_t)b * S + t;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int pl = ty + 8 * k;                      // pillar within tile
        const int pcell = p0 + pl;
        float v = 0.f;
        if ((occ_bits & (1u << (t * 4 + k))) && c < C) {
          float* src = grid + (bt * nvox + pcell) * C + c;
          v = *src;
          *src = 0.f;
        }
        acc[k] = __fadd_rn(__fmul_rn(acc[k], discount), v);   // out-of-range entries stay exactly zero
        if (out_nhwc && pcell < nvox && c < C) out[(bt * nvox + pcell) * C + c] = acc[k];
        if (use_tile) tile[pl][tx] = acc[k];
      }
      if (use_tile) {
        __syncthreads();
        if (!out_nhwc) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int cl = ty + 8 * k;
            const int cc = c0 + cl, pcell = p0 + tx;
            if (cc < C && pcell < nvox) __stcs(out + (bt * C + cc) * (size_t)nvox + pcell, tile[tx][cl]);
          }
        }
        if (pool_sum && ty == 0 && c < C) {       // per-(b,t,c) spatial sum for the pyramid-pooling branch
          float tot = 0.f;
#pragma unroll
          for (int r = 0; r < 32; ++r) tot += tile[r][tx];
          atomicAdd(pool_sum + bt * C + c, tot);
        }
        __syncthreads();
      }
    }
  }
}

}  // namespace stp3

using namespace stp3;

static size_t grid_bytes(int B, int S, int C, int nx, int ny) {
  const size_t g = (size_t)B * S * nx * ny * C * sizeof(float);
  return (g + 255) & ~(size_t)255;
}

extern "C" size_t stp3_lift_splat_workspace_bytes(int B, int S, int C, int nx, int ny) {
  if (B <= 0 || S <= 0 || C <= 0 || nx <= 0 || ny <= 0) return 0;
  const size_t o = ((size_t)B * S * nx * ny + 255) & ~(size_t)255;
  return grid_bytes(B, S, C, nx, ny) + o;
}

extern "C" int stp3_lift_splat_workspace_init(void* workspace, size_t workspace_bytes, void* stream) {
  STP3_CHECK_ARG(workspace != nullptr, "stp3_lift_splat_workspace_init: null workspace");
  STP3_CUDA_OK(cudaMemsetAsync(workspace, 0, workspace_bytes, reinterpret_cast<cudaStream_t>(stream)));
  return STP3_OK;
}

extern "C" int stp3_lift_splat_fwd(const float* feat, int feat_layout, const float* depth_logits,
                                   const float* cam_M, const float* cam_t, const float* ego_R, const float* ego_t,
                                   const float* xs, const float* ys, const float* ds,
                                   const float* bev_off, const float* bev_res,
                                   int nx, int ny, int nz, float discount,
                                   int B, int S, int N, int D, int Hf, int Wf, int C,
                                   int use_depth_distribution,
                                   int32_t* ranks_out, float* pool_sum,
                                   void* workspace, size_t workspace_bytes,
                                   float* out, int out_layout, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  STP3_CHECK_ARG(feat && cam_M && cam_t && ego_R && ego_t && xs && ys && ds && bev_off && bev_res && out && workspace,
                 "stp3_lift_splat_fwd: null pointer argument");
  STP3_CHECK_ARG(use_depth_distribution == 0 || depth_logits, "depth_logits is NULL but use_depth_distribution=1");
  STP3_CHECK_ARG(B > 0 && S > 0 && N > 0 && D > 0 && Hf > 0 && Wf > 0 && C > 0, "non-positive dimension");
  STP3_CHECK_ARG(S <= kMaxFrames, "receptive field S=%d exceeds the supported %d", S, kMaxFrames);
  STP3_CHECK_ARG(nx > 0 && ny > 0, "empty BEV grid");
  STP3_CHECK_ARG(nz == 1, "nz=%d: the reference (stp3.py:298) and this kernel support a single height bin", nz);
  STP3_CHECK_ARG((long long)nx * ny * nz < (1ll << 31), "BEV grid too large for int32 ranks");
  STP3_CHECK_ARG(feat_layout == 0 || feat_layout == 1, "feat_layout must be 0 (NCHW) or 1 (NHWC)");
  STP3_CHECK_ARG(out_layout == 0 || out_layout == 1, "out_layout must be 0 (C,X,Y) or 1 (X,Y,C)");
  STP3_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  const size_t need = stp3_lift_splat_workspace_bytes(B, S, C, nx, ny);
  if (workspace_bytes < need) return set_error(STP3_ENOSPC, "workspace too small: %zu < %zu", workspace_bytes, need);

  LiftSplatParams p;
  p.feat = feat; p.depth = depth_logits; p.cam_M = cam_M; p.cam_t = cam_t; p.ego_R = ego_R; p.ego_t = ego_t;
  p.xs = xs; p.ys = ys; p.ds = ds;
  for (int i = 0; i < 3; ++i) { p.off[i] = bev_off[i]; p.res[i] = bev_res[i]; }
  p.nx = nx; p.ny = ny; p.nz = nz;
  p.B = B; p.S = S; p.N = N; p.D = D; p.Hf = Hf; p.Wf = Wf; p.C = C;
  p.feat_nhwc = feat_layout; p.use_depth = use_depth_distribution;
  p.ranks_out = ranks_out;
  p.grid = static_cast<float*>(workspace);
  p.occ = static_cast<unsigned char*>(workspace) + grid_bytes(B, S, C, nx, ny);

  // tile width: as wide as shared memory allows (<= 4 columns), at least 1
  int TW = 4;
  auto smem_for = [&](int tw) {
    const int npix = Hf * tw;
    return (size_t)D * npix * sizeof(int2) +
           (size_t)(kCChunk * (npix | 1) + 12 * kMaxFrames + Hf + D + kScatterThreads) * sizeof(float);
  };
  while (TW > 1 && smem_for(TW) > 110 * 1024) TW >>= 1;
  const size_t smem = smem_for(TW);
  STP3_CHECK_ARG(smem <= 227 * 1024, "D*Hf = %d too large for one image column in shared memory", D * Hf);
  p.TW = TW;
  p.tiles_w = ceil_div(Wf, TW);

  STP3_CUDA_OK(cudaFuncSetAttribute(lift_splat_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long nblk = (long long)B * S * N * p.tiles_w;
  STP3_CHECK_ARG(nblk < (1ll << 31), "grid too large");
  lift_splat_scatter_kernel<<<(unsigned)nblk, kScatterThreads, smem, stream>>>(p);
  STP3_CUDA_OK(cudaGetLastError());

  const int nvox = nx * ny * nz;
  STP3_CHECK_ARG(S * 4 <= 32, "S too large for the finalize occupancy mask");
  dim3 fgrid(ceil_div(nvox, 32), B), fblock(32, 8);
  bev_finalize_kernel<<<fgrid, fblock, 0, stream>>>(p.grid, p.occ, out, pool_sum, S, C, nvox, discount, out_layout);
  STP3_CUDA_OK(cudaGetLastError());
  return STP3_OK;
}
