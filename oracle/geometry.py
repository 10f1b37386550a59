"""Oracle: geometry / warp layer (reference ``inverse_warp.py``).  TEST INFRASTRUCTURE.

All functions are fp32 torch restatements; `align_corners=False` is what torch>=1.3
executes for the reference's un-annotated ``grid_sample`` calls.
"""
import torch
import torch.nn.functional as F


def check_sizes(t, name, expected):
    """Reference inverse_warp.py:23-28 (same message so error tests read the same)."""
    cond = [t.ndimension() == len(expected)]
    for i, s in enumerate(expected):
        if s.isdigit():
            cond.append(t.size(i) == int(s))
    assert all(cond), "wrong size for {}, expected {}, got  {}".format(
        name, 'x'.join(expected), list(t.size()))


def euler2mat(angle):
    """R = X(rx) @ Y(ry) @ Z(rz).  Reference inverse_warp.py:82-119."""
    B = angle.size(0)
    x, y, z = angle[:, 0], angle[:, 1], angle[:, 2]
    zero = z.detach() * 0
    one = zero + 1
    cz, sz = torch.cos(z), torch.sin(z)
    cy, sy = torch.cos(y), torch.sin(y)
    cx, sx = torch.cos(x), torch.sin(x)
    zmat = torch.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], 1).view(B, 3, 3)
    ymat = torch.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], 1).view(B, 3, 3)
    xmat = torch.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], 1).view(B, 3, 3)
    return xmat.bmm(ymat).bmm(zmat)


def quat2mat(quat):
    """Reference inverse_warp.py:122-143."""
    nq = torch.cat([quat[:, :1].detach() * 0 + 1, quat], dim=1)
    nq = nq / nq.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = nq[:, 0], nq[:, 1], nq[:, 2], nq[:, 3]
    B = quat.size(0)
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], 1).view(B, 3, 3)


def pose_vec2mat(vec, rotation_mode='euler'):
    """[B,6] (tx,ty,tz,rx,ry,rz) -> [B,3,4].  Reference inverse_warp.py:146-162."""
    t = vec[:, :3].unsqueeze(-1)
    rot = vec[:, 3:]
    R = euler2mat(rot) if rotation_mode == 'euler' else quat2mat(rot)
    return torch.cat([R, t], dim=2)


def _id_grid(h, w, like):
    """[1,3,h,w] homogeneous pixel grid (x=col, y=row, 1). Reference inverse_warp.py:13-20."""
    ys = torch.arange(0, h, dtype=like.dtype, device=like.device).view(1, h, 1).expand(1, h, w)
    xs = torch.arange(0, w, dtype=like.dtype, device=like.device).view(1, 1, w).expand(1, h, w)
    return torch.stack((xs, ys, torch.ones_like(xs)), dim=1)


def pixel2cam(depth, intrinsics_inv):
    """cam = depth * (Kinv @ [x,y,1]).  Reference inverse_warp.py:31-45."""
    b, h, w = depth.size()
    grid = _id_grid(h, w, depth).expand(b, 3, h, w).contiguous().view(b, 3, -1)
    cam = intrinsics_inv.bmm(grid).view(b, 3, h, w)
    return cam * depth.unsqueeze(1)


def cam2pixel(cam_coords, proj_rot, proj_tr, padding_mode):
    """Project to normalised [-1,1] coords; 'zeros' rewrites OOB coords to 2 (no grad).
    Reference inverse_warp.py:48-79."""
    b, _, h, w = cam_coords.size()
    flat = cam_coords.view(b, 3, -1)
    p = proj_rot.bmm(flat) if proj_rot is not None else flat
    if proj_tr is not None:
        p = p + proj_tr
    X, Y = p[:, 0], p[:, 1]
    Z = p[:, 2].clamp(min=1e-3)
    Xn = 2 * (X / Z) / (w - 1) - 1
    Yn = 2 * (Y / Z) / (h - 1) - 1
    if padding_mode == 'zeros':
        xm = ((Xn > 1) | (Xn < -1)).detach()
        Xn = torch.where(xm, torch.full_like(Xn, 2.0), Xn)
        ym = ((Yn > 1) | (Yn < -1)).detach()
        Yn = torch.where(ym, torch.full_like(Yn, 2.0), Yn)
    return torch.stack([Xn, Yn], dim=2).view(b, h, w, 2)


def _grid_sample(img, grid, padding_mode):
    return F.grid_sample(img, grid, mode='bilinear', padding_mode=padding_mode or 'zeros',
                         align_corners=False)


def inverse_warp(img, depth, pose, intrinsics, intrinsics_inv, rotation_mode='euler',
                 padding_mode='zeros'):
    """Reference inverse_warp.py:250-283."""
    check_sizes(img, 'img', 'B3HW')
    check_sizes(depth, 'depth', 'BHW')
    check_sizes(pose, 'pose', 'B6')
    check_sizes(intrinsics, 'intrinsics', 'B33')
    check_sizes(intrinsics_inv, 'intrinsics', 'B33')
    assert intrinsics_inv.size() == intrinsics.size()
    cam = pixel2cam(depth, intrinsics_inv)
    P = intrinsics.bmm(pose_vec2mat(pose, rotation_mode))
    grid = cam2pixel(cam, P[:, :, :3], P[:, :, -1:], padding_mode)
    return _grid_sample(img, grid, padding_mode)


def flow_warp(img, flow, padding_mode='zeros'):
    """Reference inverse_warp.py:164-192 (no OOB->2 rewrite: border pixels blend with 0)."""
    check_sizes(img, 'img', 'BCHW')
    check_sizes(flow, 'flow', 'B2HW')
    bs, _, h, w = flow.size()
    u, v = flow[:, 0], flow[:, 1]
    gx = torch.arange(0, w, dtype=u.dtype, device=u.device).view(1, 1, w).expand(1, h, w).expand_as(u)
    gy = torch.arange(0, h, dtype=u.dtype, device=u.device).view(1, h, 1).expand(1, h, w).expand_as(v)
    X = 2 * ((gx + u) / (w - 1.0) - 0.5)
    Y = 2 * ((gy + v) / (h - 1.0) - 0.5)
    return _grid_sample(img, torch.stack((X, Y), dim=3), padding_mode)


def pose2flow(depth, pose, intrinsics, intrinsics_inv, rotation_mode='euler', padding_mode=None):
    """Rigid flow induced by (depth, pose).  Reference inverse_warp.py:195-220."""
    check_sizes(depth, 'depth', 'BHW')
    check_sizes(pose, 'pose', 'B6')
    check_sizes(intrinsics, 'intrinsics', 'B33')
    check_sizes(intrinsics_inv, 'intrinsics', 'B33')
    assert intrinsics_inv.size() == intrinsics.size()
    bs, h, w = depth.size()
    gx = torch.arange(0, w, dtype=depth.dtype, device=depth.device).view(1, 1, w).expand(1, h, w).expand_as(depth)
    gy = torch.arange(0, h, dtype=depth.dtype, device=depth.device).view(1, h, 1).expand(1, h, w).expand_as(depth)
    cam = pixel2cam(depth, intrinsics_inv)
    P = intrinsics.bmm(pose_vec2mat(pose, rotation_mode))
    src = cam2pixel(cam, P[:, :, :3], P[:, :, -1:], padding_mode)
    X = (w - 1) * (src[:, :, :, 0] / 2.0 + 0.5) - gx
    Y = (h - 1) * (src[:, :, :, 1] / 2.0 + 0.5) - gy
    return torch.stack((X, Y), dim=1)


def flow2oob(flow):
    """Reference inverse_warp.py:222-238."""
    check_sizes(flow, 'flow', 'B2HW')
    bs, _, h, w = flow.size()
    u, v = flow[:, 0], flow[:, 1]
    gx = torch.arange(0, w, dtype=u.dtype, device=u.device).view(1, 1, w).expand(1, h, w).expand_as(u)
    gy = torch.arange(0, h, dtype=u.dtype, device=u.device).view(1, h, 1).expand(1, h, w).expand_as(v)
    X = 2 * ((gx + u) / (w - 1.0) - 0.5)
    Y = 2 * ((gy + v) / (h - 1.0) - 0.5)
    return (X.abs() > 1) | (Y.abs() > 1)
