"""Host-side geometry of the lift-splat: BEV grid parameters, 6-DoF poses, and the small per-camera /
per-frame matrices handed to the CUDA kernel.

Mirrors the reference helpers of the same names (stp3/utils/geometry.py:40-59, 124-172) — same
argument meaning and return values — because they decide voxel indices bit-for-bit: everything here
is evaluated with the same torch CPU/GPU calls the reference makes (torch.inverse, matmul, sin/cos),
so the 18 camera matrices and S poses the kernel receives carry the reference's exact fp32 bits.
The kernel never recomputes them.
"""
import torch


def calculate_birds_eye_view_parameters(x_bounds, y_bounds, z_bounds):
    """(min, max, step) per axis -> (resolution f32[3], first-cell centre f32[3], dimension i64[3]).
    Same contract as stp3/utils/geometry.py:40-59."""
    axes = (x_bounds, y_bounds, z_bounds)
    bev_resolution = torch.tensor([a[2] for a in axes])
    bev_start_position = torch.tensor([a[0] + a[2] / 2.0 for a in axes])
    bev_dimension = torch.tensor([(a[1] - a[0]) / a[2] for a in axes], dtype=torch.long)
    return bev_resolution, bev_start_position, bev_dimension


def euler2mat(angle: torch.Tensor) -> torch.Tensor:
    """(…,3) XYZ Euler angles [rad] -> (…,3,3) rotation R = Rx·Ry·Rz (stp3/utils/geometry.py:124-155)."""
    lead = angle.shape[:-1]
    a = angle.reshape(-1, 3)
    rx, ry, rz = a[:, 0], a[:, 1], a[:, 2]
    one, zero = torch.ones_like(rz), torch.zeros_like(rz)

    def rows(*v):
        return torch.stack(v, dim=1).view(-1, 3, 3)

    cz, sz = torch.cos(rz), torch.sin(rz)
    cy, sy = torch.cos(ry), torch.sin(ry)
    cx, sx = torch.cos(rx), torch.sin(rx)
    mz = rows(cz, -sz, zero, sz, cz, zero, zero, zero, one)
    my = rows(cy, zero, sy, zero, one, zero, -sy, zero, cy)
    mx = rows(one, zero, zero, zero, cx, -sx, zero, sx, cx)
    return mx.bmm(my).bmm(mz).view(*lead, 3, 3)


def pose_vec2mat(vec: torch.Tensor) -> torch.Tensor:
    """(…,6) = (tx,ty,tz,rx,ry,rz) -> (…,4,4) homogeneous transform (stp3/utils/geometry.py:158-172)."""
    out = vec.new_zeros(*vec.shape[:-1], 4, 4)
    out[..., :3, :3] = euler2mat(vec[..., 3:].contiguous())
    out[..., :3, 3] = vec[..., :3]
    out[..., 3, 3] = 1.0
    return out


def frustum_axes(final_dim, downsample, d_bound, device=None):
    """The three 1-D axes of the reference frustum (stp3/models/stp3.py:111-130): pixel x (Wf), pixel y (Hf)
    and metric depth (D).  The (D,Hf,Wf,3) frustum is their outer expansion and is never materialised."""
    h, w = final_dim
    hf, wf = h // downsample, w // downsample
    xs = torch.linspace(0, w - 1, wf, dtype=torch.float)
    ys = torch.linspace(0, h - 1, hf, dtype=torch.float)
    ds = torch.arange(*d_bound, dtype=torch.float)
    if device is not None:
        xs, ys, ds = xs.to(device), ys.to(device), ds.to(device)
    return xs, ys, ds


def lift_matrices(intrinsics, extrinsics, future_egomotion):
    """Per-camera and per-frame matrices for the kernel, computed exactly as the reference does:
      cam_M = R · K^-1  (stp3.py:190,196), cam_t = translation (stp3.py:190,198),
      ego_R, ego_t from pose_vec2mat(future_egomotion) (stp3.py:234-235).
    intrinsics (B,S,N,3,3), extrinsics (B,S,N,4,4), future_egomotion (B,S,6)."""
    rotation, translation = extrinsics[..., :3, :3], extrinsics[..., :3, 3]
    cam_M = rotation.matmul(torch.inverse(intrinsics))
    pose = pose_vec2mat(future_egomotion)
    return (cam_M.contiguous().float(), translation.contiguous().float(),
            pose[..., :3, :3].contiguous().float(), pose[..., :3, 3].contiguous().float())


def bev_offset(bev_start_position, bev_resolution):
    """(start - res/2) as the fp32 tensor expression of stp3.py:288."""
    return bev_start_position - bev_resolution / 2.0
