"""Python entry points over the C ABI: tensors in, tensors out, current CUDA stream, no fallback."""
import ctypes
from typing import Optional

import torch

from . import _lib


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("stp3_b200 ops run on CUDA tensors only (there is no CPU path)")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def _host3(v):
    a = torch.as_tensor(v, dtype=torch.float32, device="cpu").reshape(3)
    return (ctypes.c_float * 3)(*[float(x) for x in a])   # exact: a python float holds every fp32


class Workspace:
    """Grow-only device scratch buffer reused between calls (the lift-splat's L2-resident scatter grid)."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None

    def get(self, nbytes: int, device) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
            self.reset()
        return self.buf

    def reset(self):
        """Zero the buffer (required once after allocation, and after any failed call; see include/stp3_b200.h)."""
        if self.buf is not None:
            with torch.cuda.device(self.buf.device):
                code = _lib.lib().stp3_lift_splat_workspace_init(
                    self.buf.data_ptr(), self.buf.numel(), torch.cuda.current_stream(self.buf.device).cuda_stream)
            _lib.check(code, "stp3_lift_splat_workspace_init")


_default_ws = Workspace()


def lift_splat(feat, depth_logits, cam_M, cam_t, ego_R, ego_t, xs, ys, ds, bev_off, bev_res, bev_dim,
               discount: float, *, feat_channels_last: bool = False, out_channels_last: bool = False,
               use_depth_distribution: bool = True, return_ranks: bool = False, pool_sum: bool = False,
               workspace: Optional[Workspace] = None, out: Optional[torch.Tensor] = None,
               out_hilo: Optional[torch.Tensor] = None):
    """Fused lift (softmax-depth (x) context) + ego-aligned voxel pooling + temporal discount.

      feat          (B,S,N,C,Hf,Wf) or, with feat_channels_last, (B,S,N,Hf,Wf,C)
      depth_logits  (B,S,N,D,Hf,Wf)
      cam_M/cam_t/ego_R/ego_t/xs/ys/ds  see utils.geometry.lift_matrices / frustum_axes
      bev_off, bev_res (3,) float (host values); bev_dim (3,) int
    Returns out (B,S,C,X,Y) [or (B,S,X,Y,C)], and optionally ranks (B,S,N,D,Hf,Wf) int32 and pool sums (B,S,C).
    Semantics: stp3/models/stp3.py:186-301 of the reference."""
    _require_cuda(feat, depth_logits)
    dev = feat.device
    if feat_channels_last:
        B, S, N, Hf, Wf, C = feat.shape
    else:
        B, S, N, C, Hf, Wf = feat.shape
    D = ds.numel()
    nx, ny, nz = (int(v) for v in bev_dim)
    feat = _f32c(feat)
    depth_logits = _f32c(depth_logits) if depth_logits is not None else None
    if use_depth_distribution:
        assert depth_logits is not None and tuple(depth_logits.shape) == (B, S, N, D, Hf, Wf), "depth_logits shape"
    cam_M, cam_t, ego_R, ego_t = (_f32c(t.to(dev)) for t in (cam_M, cam_t, ego_R, ego_t))
    xs, ys, ds = (_f32c(t.to(dev)) for t in (xs, ys, ds))
    assert cam_M.shape == (B, S, N, 3, 3) and cam_t.shape == (B, S, N, 3)
    assert ego_R.shape == (B, S, 3, 3) and ego_t.shape == (B, S, 3)
    assert xs.numel() == Wf and ys.numel() == Hf
    L = _lib.lib()
    need = L.stp3_lift_splat_workspace_bytes(B, S, C, nx, ny)
    ws = (workspace or _default_ws).get(need, dev)
    oshape = (B, S, nx, ny, C) if out_channels_last else (B, S, C, nx, ny)
    layout = int(out_channels_last)
    if out_hilo is not None:      # two bf16 planes (2,B,S,nx,ny,C): the activation format of the tensor-core path
        assert tuple(out_hilo.shape) == (2, B, S, nx, ny, C) and out_hilo.dtype == torch.bfloat16 and out_hilo.is_contiguous()
        out, layout = out_hilo, 2
    elif out is None:
        out = torch.empty(oshape, dtype=torch.float32, device=dev)
    else:
        assert tuple(out.shape) == oshape and out.is_contiguous() and out.dtype == torch.float32
    ranks = torch.empty((B, S, N, D, Hf, Wf), dtype=torch.int32, device=dev) if return_ranks else None
    psum = torch.zeros((B, S, C), dtype=torch.float32, device=dev) if pool_sum else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        code = L.stp3_lift_splat_fwd(
            feat.data_ptr(), int(feat_channels_last), depth_logits.data_ptr() if depth_logits is not None else None,
            cam_M.data_ptr(), cam_t.data_ptr(), ego_R.data_ptr(), ego_t.data_ptr(),
            xs.data_ptr(), ys.data_ptr(), ds.data_ptr(), _host3(bev_off), _host3(bev_res),
            nx, ny, nz, float(discount), B, S, N, D, Hf, Wf, C, int(use_depth_distribution),
            ranks.data_ptr() if ranks is not None else None, psum.data_ptr() if psum is not None else None,
            ws.data_ptr(), ws.numel(), out.data_ptr(), layout, stream)
    if code != 0:
        (workspace or _default_ws).reset()
    _lib.check(code, "stp3_lift_splat_fwd")
    res = (out,)
    if return_ranks:
        res += (ranks,)
    if pool_sum:
        res += (psum,)
    return res if len(res) > 1 else out


def lift_splat_backward(grad_out, feat, depth_logits, cam_M, cam_t, ego_R, ego_t, xs, ys, ds, bev_off, bev_res, bev_dim,
                        discount: float, *, use_depth_distribution: bool = True):
    """Gradients of lift_splat()'s (B,S,C,X,Y) output w.r.t. feat (B,S,N,C,Hf,Wf) and depth_logits (B,S,N,D,Hf,Wf):
    stp3_lift_splat_bwd (include/stp3_b200.h).  Returns (grad_feat, grad_depth_logits or None)."""
    _require_cuda(feat, depth_logits)
    dev = feat.device
    B, S, N, C, Hf, Wf = feat.shape
    D = ds.numel()
    nx, ny, nz = (int(v) for v in bev_dim)
    grad_out = _f32c(grad_out)
    assert tuple(grad_out.shape) == (B, S, C, nx, ny)
    feat = _f32c(feat)
    depth_logits = _f32c(depth_logits) if depth_logits is not None else None
    cam_M, cam_t, ego_R, ego_t = (_f32c(t.to(dev)) for t in (cam_M, cam_t, ego_R, ego_t))
    xs, ys, ds = (_f32c(t.to(dev)) for t in (xs, ys, ds))
    L = _lib.lib()
    nbytes = L.stp3_lift_splat_bwd_scratch_bytes(B, S, C, nx, ny)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    g_feat = torch.empty_like(feat)
    g_depth = torch.empty_like(depth_logits) if (use_depth_distribution and depth_logits is not None) else None
    with torch.cuda.device(dev):
        code = L.stp3_lift_splat_bwd(
            grad_out.data_ptr(), feat.data_ptr(), depth_logits.data_ptr() if depth_logits is not None else None,
            cam_M.data_ptr(), cam_t.data_ptr(), ego_R.data_ptr(), ego_t.data_ptr(), xs.data_ptr(), ys.data_ptr(),
            ds.data_ptr(), _host3(bev_off), _host3(bev_res), nx, ny, nz, float(discount), B, S, N, D, Hf, Wf, C,
            int(use_depth_distribution), scratch.data_ptr(), nbytes, g_feat.data_ptr(),
            g_depth.data_ptr() if g_depth is not None else None, torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(code, "stp3_lift_splat_bwd")
    return g_feat, g_depth


class LiftSplatFunction(torch.autograd.Function):
    """Differentiable lift-splat: forward = stp3_lift_splat_fwd, backward = stp3_lift_splat_bwd.  Gradients flow to the
    context features and the depth logits; calibration, ego-motion and the integer voxel indices carry none (like the
    reference, whose indices pass through .long(), stp3.py:289)."""

    @staticmethod
    def forward(ctx, feat, depth_logits, cam_M, cam_t, ego_R, ego_t, xs, ys, ds, bev_off, bev_res, bev_dim, discount,
                use_depth_distribution, workspace):
        out = lift_splat(feat, depth_logits, cam_M, cam_t, ego_R, ego_t, xs, ys, ds, bev_off, bev_res, bev_dim, discount,
                         use_depth_distribution=use_depth_distribution, workspace=workspace)
        ctx.save_for_backward(feat, depth_logits if depth_logits is not None else feat.new_empty(0),
                              cam_M, cam_t, ego_R, ego_t, xs, ys, ds)
        ctx.host = (bev_off, bev_res, bev_dim, discount, use_depth_distribution, depth_logits is not None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        feat, depth_logits, cam_M, cam_t, ego_R, ego_t, xs, ys, ds = ctx.saved_tensors
        bev_off, bev_res, bev_dim, discount, use_dd, has_depth = ctx.host
        g_feat, g_depth = lift_splat_backward(grad_out, feat, depth_logits if has_depth else None, cam_M, cam_t, ego_R,
                                              ego_t, xs, ys, ds, bev_off, bev_res, bev_dim, discount,
                                              use_depth_distribution=use_dd)
        return (g_feat, g_depth) + (None,) * 13


def lift_splat_autograd(feat, depth_logits, cam_M, cam_t, ego_R, ego_t, xs, ys, ds, bev_off, bev_res, bev_dim,
                        discount: float, *, use_depth_distribution: bool = True, workspace: Optional[Workspace] = None):
    """lift_splat() (reference layout: feat NCHW in, (B,S,C,X,Y) out) with gradients to feat and depth_logits."""
    return LiftSplatFunction.apply(feat, depth_logits, cam_M, cam_t, ego_R, ego_t, xs, ys, ds, bev_off, bev_res, bev_dim,
                                   float(discount), bool(use_depth_distribution), workspace)


def lift_splat_frames(feat, depth_logits, cam_M, cam_t, ego_R, ego_t, xs, ys, ds, bev_off, bev_res, bev_dim,
                      f_begin: int, f_count: int, *, feat_channels_last: bool = False,
                      use_depth_distribution: bool = True, workspace: Optional[Workspace] = None,
                      peer_ptrs=None) -> Optional[torch.Tensor]:
    """Frame-sharded lift-splat: RAW (no discount recurrence) channels-last splats (f_count, X, Y, C) fp32 of the flat
    frames f_begin + [0, f_count), f = b*S + t.  See stp3_lift_splat_frames_fwd in include/stp3_b200.h.
    peer_ptrs (list of device pointers, one (B*S, X, Y, C) fp32 buffer per rank): the finalize epilogue stores the frames
    into slot f of every rank's buffer instead (stp3_lift_splat_frames_allgather_fwd); returns None."""
    _require_cuda(feat, depth_logits)
    dev = feat.device
    if feat_channels_last:
        B, S, N, Hf, Wf, C = feat.shape
    else:
        B, S, N, C, Hf, Wf = feat.shape
    D = ds.numel()
    nx, ny, nz = (int(v) for v in bev_dim)
    feat = _f32c(feat)
    depth_logits = _f32c(depth_logits) if depth_logits is not None else None
    cam_M, cam_t, ego_R, ego_t = (_f32c(t.to(dev)) for t in (cam_M, cam_t, ego_R, ego_t))
    xs, ys, ds = (_f32c(t.to(dev)) for t in (xs, ys, ds))
    L = _lib.lib()
    need = L.stp3_lift_splat_workspace_bytes(f_count, 1, C, nx, ny)
    ws = (workspace or _default_ws).get(need, dev)
    if peer_ptrs is not None:
        arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(x) for x in peer_ptrs])
        with torch.cuda.device(dev):
            code = L.stp3_lift_splat_frames_allgather_fwd(
                feat.data_ptr(), int(feat_channels_last), depth_logits.data_ptr() if depth_logits is not None else None,
                cam_M.data_ptr(), cam_t.data_ptr(), ego_R.data_ptr(), ego_t.data_ptr(), xs.data_ptr(), ys.data_ptr(),
                ds.data_ptr(), _host3(bev_off), _host3(bev_res), nx, ny, nz, B, S, N, D, Hf, Wf, C,
                int(use_depth_distribution), int(f_begin), int(f_count), ws.data_ptr(), ws.numel(), len(peer_ptrs), arr,
                torch.cuda.current_stream(dev).cuda_stream)
        if code != 0:
            (workspace or _default_ws).reset()
        _lib.check(code, "stp3_lift_splat_frames_allgather_fwd")
        return None
    out = torch.empty((f_count, nx, ny, C), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        code = L.stp3_lift_splat_frames_fwd(
            feat.data_ptr(), int(feat_channels_last), depth_logits.data_ptr() if depth_logits is not None else None,
            cam_M.data_ptr(), cam_t.data_ptr(), ego_R.data_ptr(), ego_t.data_ptr(), xs.data_ptr(), ys.data_ptr(),
            ds.data_ptr(), _host3(bev_off), _host3(bev_res), nx, ny, nz, B, S, N, D, Hf, Wf, C,
            int(use_depth_distribution), int(f_begin), int(f_count), ws.data_ptr(), ws.numel(), out.data_ptr(),
            torch.cuda.current_stream(dev).cuda_stream)
    if code != 0:
        (workspace or _default_ws).reset()
    _lib.check(code, "stp3_lift_splat_frames_fwd")
    return out


def bev_discount(raw: torch.Tensor, discount: float) -> torch.Tensor:
    """raw (B,S,X,Y,C) fp32 channels-last per-frame splats -> (2,B,S,X,Y,C) bf16 hi/lo planes of
    out[t] = out[t-1]*discount + raw[t]  (stp3.py:296)."""
    _require_cuda(raw)
    raw = _f32c(raw)
    B, S, X, Y, C = raw.shape
    planes = torch.empty((2, B, S, X, Y, C), dtype=torch.bfloat16, device=raw.device)
    with torch.cuda.device(raw.device):
        code = _lib.lib().stp3_bev_discount(raw.data_ptr(), B, S, X, Y, C, float(discount), planes[0].data_ptr(),
                                            planes[1].data_ptr(), torch.cuda.current_stream(raw.device).cuda_stream)
    _lib.check(code, "stp3_bev_discount")
    return planes
