// Frustum point -> ego frame -> ego-motion chain -> BEV pillar rank, in the reference's exact fp32 operation order
// (stp3/models/stp3.py:186-201, 265-289, 239-255).  Shared by the forward and backward lift-splat kernels: every
// kernel that needs the voxel index of a lifted point must derive it through these functions so that the indices stay
// bit-identical with the reference's CPU path.
#pragma once
#include <cuda_runtime.h>

namespace stp3 {

// ((m0*x + m1*y) + m2*z) + t, every product and sum rounded to fp32 separately: bit-identical to the reference's
// batched 3x3 matmul followed by `+= translation` (stp3.py:197-198, 273-277).  No FMA contraction.
__device__ __forceinline__ void affine_exact(const float* __restrict__ m, const float* __restrict__ t,
                                             float& x, float& y, float& z) {
  const float ox = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], x), __fmul_rn(m[1], y)), __fmul_rn(m[2], z)), t[0]);
  const float oy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[3], x), __fmul_rn(m[4], y)), __fmul_rn(m[5], z)), t[1]);
  const float oz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[6], x), __fmul_rn(m[7], y)), __fmul_rn(m[8], z)), t[2]);
  x = ox; y = oy; z = oz;
}

// quantisation constants of the BEV grid (host values of stp3.py:288: offset = start - res/2, true division)
struct BevQuant {
  float off[3];
  float res[3];
  float inv[3];     // exact 1/res when res is a power of two (the division is then an exact scaling)
  int inv_ok[3];
  int nx, ny, nz;
};

// (ego-frame point) -> pillar rank ix*(ny*nz) + iy*nz + iz, or -1 outside the grid.
// ((p - offset) / res).long(): true division, truncation toward zero; trunc(q) in [0, n) <=> -1 < q < n, which keeps
// the (-1, 0) band in cell 0 exactly like .long() followed by the >= 0 mask (stp3.py:239-246).
__device__ __forceinline__ int quantise_rank(const BevQuant& q, float x, float y, float z) {
  const float qx = q.inv_ok[0] ? __fmul_rn(__fsub_rn(x, q.off[0]), q.inv[0]) : __fdiv_rn(__fsub_rn(x, q.off[0]), q.res[0]);
  const float qy = q.inv_ok[1] ? __fmul_rn(__fsub_rn(y, q.off[1]), q.inv[1]) : __fdiv_rn(__fsub_rn(y, q.off[1]), q.res[1]);
  const float qz = q.inv_ok[2] ? __fmul_rn(__fsub_rn(z, q.off[2]), q.inv[2]) : __fdiv_rn(__fsub_rn(z, q.off[2]), q.res[2]);
  const bool keep = (qx > -1.f) && (qx < (float)q.nx) && (qy > -1.f) && (qy < (float)q.ny) && (qz > -1.f) && (qz < (float)q.nz);
  if (!keep) return -1;
  return (int)qx * (q.ny * q.nz) + (int)qy * q.nz + (int)qz;      // cvt.rzi == .long()
}

// The same for the grids the reference configures: x / y resolutions that are powers of two (0.5 m: the divisions are
// exact scalings) and ONE height bin (nz == 1, which the kernels require anyway).  With a single bin the z index is 0 and
// only the mask needs z: trunc(fl(a / r)) == 0  <=>  -1 < fl(a / r) < 1  <=>  -r < a < r  for a = z - off_z, r > 0 --
// a correctly rounded quotient of two floats is below 1 exactly when the numerator is below the denominator (a <= pred(r)
// gives a / r <= 1 - 2^-24, which is representable) -- so the true division (height resolution 20 m is not a power of
// two) is never evaluated.
__device__ __forceinline__ int quantise_rank_pow2xy_nz1(const BevQuant& q, float x, float y, float z) {
  const float qx = __fmul_rn(__fsub_rn(x, q.off[0]), q.inv[0]);
  const float qy = __fmul_rn(__fsub_rn(y, q.off[1]), q.inv[1]);
  const float az = __fsub_rn(z, q.off[2]);
  const bool keep = (qx > -1.f) && (qx < (float)q.nx) && (qy > -1.f) && (qy < (float)q.ny) && (az > -q.res[2]) && (az < q.res[2]);
  if (!keep) return -1;
  return (int)qx * q.ny + (int)qy;
}

// rank of the frustum point (pixel u = xw, v = yh, depth dep): camera transform `cam` (9 + 3 floats), then n_chain
// ego-motion links `chain` (12 floats each), sequential and rounded at every step (stp3.py:270-277)
__device__ __forceinline__ int lifted_point_rank(const BevQuant& q, const float* __restrict__ cam,
                                                 const float* __restrict__ chain, int n_chain, float xw, float yh,
                                                 float dep) {
  float x = __fmul_rn(xw, dep);                  // stp3.py:195: (u*d, v*d, d)
  float y = __fmul_rn(yh, dep);
  float z = dep;
  affine_exact(cam, cam + 9, x, y, z);           // stp3.py:196-198
  for (int k = 0; k < n_chain; ++k) affine_exact(chain + 12 * k, chain + 12 * k + 9, x, y, z);
  return quantise_rank(q, x, y, z);
}

}  // namespace stp3
