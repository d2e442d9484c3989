/* stp3_b200 — C ABI of the B200-native (sm_100a) ST-P3 camera->BEV perception hot path.
 *
 * The reference (OpenDriveLab/ST-P3) is 100 % Python/PyTorch and has no FFI of its own; the drop-in boundary is
 * the nn.Module surface.  This header is the C boundary that sits directly underneath those
 * modules: each entry point names the reference interface it replaces (file:line under /root/reference).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's caching allocator in practice), unless the
 *     parameter is documented as a host value;
 *   - no entry point allocates, frees or synchronises: work is enqueued on `stream` (a cudaStream_t passed as
 *     void*) and the call returns immediately;
 *   - return value: 0 on success, a negative STP3_E* code otherwise; stp3_last_error() returns the text for the
 *     calling thread;
 *   - re-entrant; no global mutable state apart from the thread-local error string.
 */
#ifndef STP3_B200_H_
#define STP3_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STP3_OK 0
#define STP3_EINVAL (-1)   /* bad argument (shape, alignment, null pointer) */
#define STP3_ENOSPC (-2)   /* workspace too small */
#define STP3_ECUDA (-3)    /* CUDA runtime / driver error (launch failure, no device) */
#define STP3_EUNSUPPORTED (-4)

/* Library / build identification ("sm_100a", ABI version). */
int stp3_abi_version(void);
const char* stp3_build_info(void);
const char* stp3_last_error(void);

/* ---------------------------------------------------------------------------------------------------------
 * Lift-splat: softmax(depth) (x) context outer product, ego-motion aligned voxel pooling, temporal discount.
 *
 * Replaces, fused into one scatter kernel + one finalize kernel:
 *   STP3.get_geometry                      stp3/models/stp3.py:186-201   (frustum -> ego points)
 *   STP3.encoder_forward (softmax, outer)  stp3/models/stp3.py:214-221   (never materialised)
 *   STP3.projection_to_birds_eye_view      stp3/models/stp3.py:226-301   (ego warp, voxel index, mask, pooling,
 *                                                                          discount recurrence, (C,X,Y) layout)
 *   VoxelsSumming.forward                  stp3/utils/geometry.py:299-318
 *
 * Layouts (all fp32 unless noted)
 *   feat          feat_layout==0: (B,S,N,C,Hf,Wf)  [the reference Encoder's NCHW output]
 *                 feat_layout==1: (B,S,N,Hf,Wf,C)  [channels-last, produced by this library's conv kernels]
 *   depth_logits  (B,S,N,D,Hf,Wf); ignored (may be NULL) when use_depth_distribution==0 (stp3.py:218)
 *   cam_M         (B,S,N,3,3) = R . K^-1   evaluated on the host with the reference's torch calls
 *   cam_t         (B,S,N,3)
 *   ego_R, ego_t  (B,S,3,3), (B,S,3)  from pose_vec2mat(future_egomotion)  (geometry.py:158-172)
 *   xs, ys, ds    (Wf), (Hf), (D): the three axes of STP3.create_frustum (stp3.py:111-130)
 *   bev_off       HOST float[3] = (bev_start_position - bev_resolution/2) as evaluated in fp32 by torch (stp3.py:288)
 *   bev_res       HOST float[3]
 *   ranks_out     optional (B,S,N,D,Hf,Wf) int32: pillar rank ix*(ny*nz)+iy*nz+iz of every lifted point after the
 *                 ego warp, -1 where the reference's mask (stp3.py:239-246) drops it.  Bit-exact with the CPU
 *                 reference; used by the parity tests.
 *   out           out_layout==0: (B,S,C,nx,ny) fp32  [what projection_to_birds_eye_view returns]
 *                 out_layout==1: (B,S,nx,ny,C) fp32  [channels-last, consumed by the temporal block kernels]
 *   pool_sum      optional (B,S,C) fp32: sum over the nx*ny cells of out[b,t,c] (feeds the pyramid-pooling branch
 *                 of TemporalBlock, temporal.py:408-423); must be zero-initialised by the caller.
 *   workspace     >= stp3_lift_splat_workspace_bytes(...) bytes, 256-byte aligned.  It holds the channels-last
 *                 fp32 scatter grid (B,S,nx*ny,C) and a per-pillar occupancy map.  It must be ALL ZERO on entry:
 *                 clear it once with stp3_lift_splat_workspace_init() after allocating it; every successful
 *                 stp3_lift_splat_fwd leaves it zero again (the finalize kernel re-zeroes exactly the pillars it
 *                 read), so no per-call memset is needed.  After a failed call, re-initialise it.
 * nz must be 1 (the reference's squeeze(0) at stp3.py:298 assumes it).
 */
size_t stp3_lift_splat_workspace_bytes(int B, int S, int C, int nx, int ny);
int stp3_lift_splat_workspace_init(void* workspace, size_t workspace_bytes, void* stream);

int stp3_lift_splat_fwd(const float* feat, int feat_layout, const float* depth_logits,
                        const float* cam_M, const float* cam_t, const float* ego_R, const float* ego_t,
                        const float* xs, const float* ys, const float* ds,
                        const float* bev_off /*host[3]*/, const float* bev_res /*host[3]*/,
                        int nx, int ny, int nz, float discount,
                        int B, int S, int N, int D, int Hf, int Wf, int C,
                        int use_depth_distribution,
                        int32_t* ranks_out, float* pool_sum,
                        void* workspace, size_t workspace_bytes,
                        float* out, int out_layout, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STP3_B200_H_ */
