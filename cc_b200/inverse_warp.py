"""Drop-in for the reference's ``inverse_warp`` module (inverse_warp.py), backed by libccb200.

Same names, argument meaning and AssertionError text as the reference; every function launches
hand-written sm_100a kernels (cc_b200/csrc/warp_ops.cu) and supports autograd through
hand-derived backward kernels.  ``train.py:22`` imports inverse_warp, pose2flow, flow2oob, flow_warp."""
from __future__ import division
import torch
from . import _lib

_ROT = {'euler': _lib.ROT_EULER, 'quat': _lib.ROT_QUAT}
_PAD = {'zeros': _lib.PAD_ZEROS, 'border': _lib.PAD_BORDER, None: _lib.PAD_NONE}


def check_sizes(input, input_name, expected):
    """Reference inverse_warp.py:23-28."""
    condition = [input.ndimension() == len(expected)]
    for i, size in enumerate(expected):
        if size.isdigit():
            condition.append(input.size(i) == int(size))
    assert(all(condition)), "wrong size for {}, expected {}, got  {}".format(input_name, 'x'.join(expected), list(input.size()))


def _f(t):
    return _lib.contig(t.detach().float())


class _InverseWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, depth, pose, K, Kinv, rot, pad):
        img, depth, pose, K, Kinv = _f(img), _f(depth), _f(pose), _f(K), _f(Kinv)
        B, _, h, w = img.shape
        out = torch.empty_like(img)
        L = _lib.lib()
        _lib.check(L.ccb_inverse_warp_fwd(_lib.ptr(img), _lib.ptr(depth), _lib.ptr(pose), 6, _lib.ptr(K), _lib.ptr(Kinv),
                                          B, h, w, rot, pad, _lib.ptr(out), _lib.stream(img)), 'inverse_warp_fwd')
        ctx.save_for_backward(img, depth, pose, K, Kinv)
        ctx.cfg = (rot, pad)
        return out

    @staticmethod
    def backward(ctx, g):
        img, depth, pose, K, Kinv = ctx.saved_tensors
        rot, pad = ctx.cfg
        B, _, h, w = img.shape
        g = _f(g)
        L = _lib.lib()
        d_depth = torch.empty_like(depth)
        d_pose = torch.empty_like(pose)
        part = torch.empty(L.ccb_warp_pose_partials_floats(B, h, w), device=img.device)
        _lib.check(L.ccb_inverse_warp_bwd(_lib.ptr(img), _lib.ptr(depth), _lib.ptr(pose), 6, _lib.ptr(K), _lib.ptr(Kinv),
                                          B, h, w, rot, pad, _lib.ptr(g), _lib.ptr(d_depth), _lib.ptr(d_pose),
                                          _lib.ptr(part), _lib.stream(img)), 'inverse_warp_bwd')
        return None, d_depth, d_pose, None, None, None, None


class _Pose2Flow(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, pose, K, Kinv, rot, pad):
        depth, pose, K, Kinv = _f(depth), _f(pose), _f(K), _f(Kinv)
        B, h, w = depth.shape
        out = torch.empty(B, 2, h, w, device=depth.device)
        L = _lib.lib()
        _lib.check(L.ccb_pose2flow_fwd(_lib.ptr(depth), _lib.ptr(pose), 6, _lib.ptr(K), _lib.ptr(Kinv), B, h, w, rot, pad,
                                       _lib.ptr(out), _lib.stream(depth)), 'pose2flow_fwd')
        ctx.save_for_backward(depth, pose, K, Kinv)
        ctx.cfg = (rot, pad)
        return out

    @staticmethod
    def backward(ctx, g):
        depth, pose, K, Kinv = ctx.saved_tensors
        rot, pad = ctx.cfg
        B, h, w = depth.shape
        g = _f(g)
        L = _lib.lib()
        d_depth = torch.empty_like(depth)
        d_pose = torch.empty_like(pose)
        part = torch.empty(L.ccb_warp_pose_partials_floats(B, h, w), device=depth.device)
        _lib.check(L.ccb_pose2flow_bwd(_lib.ptr(depth), _lib.ptr(pose), 6, _lib.ptr(K), _lib.ptr(Kinv), B, h, w, rot, pad,
                                       _lib.ptr(g), _lib.ptr(d_depth), _lib.ptr(d_pose), _lib.ptr(part),
                                       _lib.stream(depth)), 'pose2flow_bwd')
        return d_depth, d_pose, None, None, None, None


class _FlowWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, flow, pad):
        img, flow = _f(img), _f(flow)
        B, Cc, h, w = img.shape
        out = torch.empty_like(img)
        _lib.check(_lib.lib().ccb_flow_warp_fwd(_lib.ptr(img), _lib.ptr(flow), B, Cc, h, w, pad, _lib.ptr(out),
                                                _lib.stream(img)), 'flow_warp_fwd')
        ctx.save_for_backward(img, flow)
        ctx.pad = pad
        return out

    @staticmethod
    def backward(ctx, g):
        img, flow = ctx.saved_tensors
        B, Cc, h, w = img.shape
        g = _f(g)
        d_flow = torch.empty_like(flow) if ctx.needs_input_grad[1] else None
        d_img = torch.zeros_like(img) if ctx.needs_input_grad[0] else None
        _lib.check(_lib.lib().ccb_flow_warp_bwd(_lib.ptr(img), _lib.ptr(flow), B, Cc, h, w, ctx.pad, _lib.ptr(g),
                                                _lib.ptr(d_flow), _lib.ptr(d_img), _lib.stream(img)), 'flow_warp_bwd')
        return d_img, d_flow, None


def euler2mat(angle):
    """Reference inverse_warp.py:82-119 (host-side torch ops; tiny, not on the per-pixel path)."""
    B = angle.size(0)
    x, y, z = angle[:, 0], angle[:, 1], angle[:, 2]
    zeros = z.detach() * 0
    ones = zeros.detach() + 1
    cosz, sinz = torch.cos(z), torch.sin(z)
    zmat = torch.stack([cosz, -sinz, zeros, sinz, cosz, zeros, zeros, zeros, ones], dim=1).view(B, 3, 3)
    cosy, siny = torch.cos(y), torch.sin(y)
    ymat = torch.stack([cosy, zeros, siny, zeros, ones, zeros, -siny, zeros, cosy], dim=1).view(B, 3, 3)
    cosx, sinx = torch.cos(x), torch.sin(x)
    xmat = torch.stack([ones, zeros, zeros, zeros, cosx, -sinx, zeros, sinx, cosx], dim=1).view(B, 3, 3)
    return xmat.bmm(ymat).bmm(zmat)


def quat2mat(quat):
    """Reference inverse_warp.py:122-143."""
    norm_quat = torch.cat([quat[:, :1].detach() * 0 + 1, quat], dim=1)
    norm_quat = norm_quat / norm_quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = norm_quat[:, 0], norm_quat[:, 1], norm_quat[:, 2], norm_quat[:, 3]
    B = quat.size(0)
    w2, x2, y2, z2 = w.pow(2), x.pow(2), y.pow(2), z.pow(2)
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1).view(B, 3, 3)


def pose_vec2mat(vec, rotation_mode='euler'):
    """[B,6] -> [B,3,4].  Reference inverse_warp.py:146-162 (used by test_pose.py:77)."""
    translation = vec[:, :3].unsqueeze(-1)
    rot = vec[:, 3:]
    if rotation_mode == 'euler':
        rot_mat = euler2mat(rot)
    elif rotation_mode == 'quat':
        rot_mat = quat2mat(rot)
    return torch.cat([rot_mat, translation], dim=2)


def inverse_warp(img, depth, pose, intrinsics, intrinsics_inv, rotation_mode='euler', padding_mode='zeros'):
    """Inverse warp a source image to the target image plane.  Reference inverse_warp.py:250-283."""
    check_sizes(img, 'img', 'B3HW')
    check_sizes(depth, 'depth', 'BHW')
    check_sizes(pose, 'pose', 'B6')
    check_sizes(intrinsics, 'intrinsics', 'B33')
    check_sizes(intrinsics_inv, 'intrinsics', 'B33')
    assert(intrinsics_inv.size() == intrinsics.size())
    return _InverseWarp.apply(img, depth, pose, intrinsics, intrinsics_inv, _ROT[rotation_mode], _PAD[padding_mode])


def flow_warp(img, flow, padding_mode='zeros'):
    """Reference inverse_warp.py:164-192."""
    check_sizes(img, 'img', 'BCHW')
    check_sizes(flow, 'flow', 'B2HW')
    return _FlowWarp.apply(img, flow, _PAD[padding_mode])


def pose2flow(depth, pose, intrinsics, intrinsics_inv, rotation_mode='euler', padding_mode=None):
    """Converts pose parameters to rigid optical flow.  Reference inverse_warp.py:195-220."""
    check_sizes(depth, 'depth', 'BHW')
    check_sizes(pose, 'pose', 'B6')
    check_sizes(intrinsics, 'intrinsics', 'B33')
    check_sizes(intrinsics_inv, 'intrinsics', 'B33')
    assert(intrinsics_inv.size() == intrinsics.size())
    return _Pose2Flow.apply(depth, pose, intrinsics, intrinsics_inv, _ROT[rotation_mode], _PAD[padding_mode])


def flow2oob(flow):
    """Boolean out-of-bounds map (validation only).  Reference inverse_warp.py:222-238."""
    check_sizes(flow, 'flow', 'B2HW')
    bs, _, h, w = flow.size()
    u, v = flow[:, 0], flow[:, 1]
    gx = torch.arange(0, w, device=flow.device, dtype=flow.dtype).view(1, 1, w).expand_as(u)
    gy = torch.arange(0, h, device=flow.device, dtype=flow.dtype).view(1, h, 1).expand_as(v)
    X = 2 * ((gx + u) / (w - 1.0) - 0.5)
    Y = 2 * ((gy + v) / (h - 1.0) - 0.5)
    return (X.abs() > 1).add(Y.abs() > 1) > 0
