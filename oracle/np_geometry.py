"""Oracle (numpy): the exact fp32 operation order of the projection path, one IEEE op per line.
TEST INFRASTRUCTURE.

This is the arithmetic *recipe* the CUDA kernels follow for the in-bounds / valid masks
(BASELINE.json: bit-exact masks).  It is pinned bit-for-bit against the torch restatement
(oracle.geometry.cam2pixel / pose2flow == reference inverse_warp.py:31-79,195-220 as executed by
the CPU ATen kernels) in tests/test_oracle_golden.py::test_np_recipe_bit_exact.

Findings it encodes (measured in the build container):
  * [B,3,3] x [B,3,N] bmm (N = h*w) rounds like an FMA chain  fma(a2,b2, fma(a1,b1, a0*b0));
  * small bmm ([3,3]x[3,3], [3,3]x[3,4]) rounds WITHOUT fma: (a0*b0 + a1*b1) + a2*b2;
  * tensor / python-scalar is a true IEEE division on CPU.
"""
import numpy as np

f32 = np.float32


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def small_matmul(A, Bm):
    """[...,3,3] @ [...,3,n] with (p0+p1)+p2 rounding, fp32."""
    A = A.astype(f32)
    Bm = Bm.astype(f32)
    p0 = A[..., :, 0:1] * Bm[..., 0:1, :]
    p1 = A[..., :, 1:2] * Bm[..., 1:2, :]
    p2 = A[..., :, 2:3] * Bm[..., 2:3, :]
    return ((p0 + p1).astype(f32) + p2).astype(f32)


def euler_R(rx, ry, rz, sincos=None):
    """R = (X @ Y) @ Z built exactly like reference euler2mat (inverse_warp.py:92-119)."""
    if sincos is None:
        sx, cx, sy, cy, sz, cz = [fn(a.astype(f32)).astype(f32) for a in (rx, ry, rz) for fn in (np.sin, np.cos)]
    else:
        sx, cx, sy, cy, sz, cz = sincos
    B = rx.shape[0]
    z, o = np.zeros(B, f32), np.ones(B, f32)
    Z = np.stack([cz, -sz, z, sz, cz, z, z, z, o], 1).reshape(B, 3, 3)
    Y = np.stack([cy, z, sy, z, o, z, -sy, z, cy], 1).reshape(B, 3, 3)
    X = np.stack([o, z, z, z, cx, -sx, z, sx, cx], 1).reshape(B, 3, 3)
    return small_matmul(small_matmul(X, Y), Z)


def project(depth, K, Kinv, T, rewrite):
    """depth [B,h,w], K/Kinv [B,3,3], T=[R|t] [B,3,4] -> Xn, Yn [B,h,w] (+ rewritten flags)."""
    B, h, w = depth.shape
    P = small_matmul(K, T)                                   # [B,3,4]
    xs = np.arange(w, dtype=f32)[None, None, :].repeat(h, 1)
    ys = np.arange(h, dtype=f32)[None, :, None].repeat(w, 2)
    xs, ys = np.broadcast_to(xs, (B, h, w)), np.broadcast_to(ys, (B, h, w))
    kv = Kinv.astype(f32)
    cam = []
    for r in range(3):
        k0, k1, k2 = [kv[:, r, j][:, None, None] for j in range(3)]
        ray = _fma(np.broadcast_to(k2, xs.shape), np.ones_like(xs),
                   _fma(np.broadcast_to(k1, xs.shape), ys, (k0 * xs).astype(f32)))
        cam.append((ray * depth.astype(f32)).astype(f32))
    p = []
    for r in range(3):
        a0, a1, a2, t = [np.broadcast_to(P[:, r, j][:, None, None], xs.shape) for j in range(4)]
        v = _fma(a2, cam[2], _fma(a1, cam[1], (a0 * cam[0]).astype(f32)))
        p.append((v + t).astype(f32))
    X, Y = p[0], p[1]
    Z = np.maximum(p[2], f32(1e-3))
    Xn = ((f32(2) * (X / Z).astype(f32)).astype(f32) / f32(w - 1)).astype(f32) - f32(1)
    Yn = ((f32(2) * (Y / Z).astype(f32)).astype(f32) / f32(h - 1)).astype(f32) - f32(1)
    Xn, Yn = Xn.astype(f32), Yn.astype(f32)
    xm = (Xn > 1) | (Xn < -1)
    ym = (Yn > 1) | (Yn < -1)
    if rewrite:
        Xn = np.where(xm, f32(2), Xn)
        Yn = np.where(ym, f32(2), Yn)
    return Xn, Yn, xm, ym


def coords_to_flow(Xn, Yn):
    """pose2flow tail (inverse_warp.py:217-218)."""
    B, h, w = Xn.shape
    xs = np.arange(w, dtype=f32)[None, None, :]
    ys = np.arange(h, dtype=f32)[None, :, None]
    u = (f32(w - 1) * ((Xn / f32(2)).astype(f32) + f32(0.5)).astype(f32)).astype(f32) - xs
    v = (f32(h - 1) * ((Yn / f32(2)).astype(f32) + f32(0.5)).astype(f32)).astype(f32) - ys
    return u.astype(f32), v.astype(f32)


def flow_coords(flow):
    """flow_warp grid (inverse_warp.py:181-188)."""
    B, _, h, w = flow.shape
    xs = np.arange(w, dtype=f32)[None, None, :]
    ys = np.arange(h, dtype=f32)[None, :, None]
    X = (xs + flow[:, 0].astype(f32)).astype(f32)
    Y = (ys + flow[:, 1].astype(f32)).astype(f32)
    Xn = (f32(2) * ((X / f32(w - 1.0)).astype(f32) - f32(0.5)).astype(f32)).astype(f32)
    Yn = (f32(2) * ((Y / f32(h - 1.0)).astype(f32) - f32(0.5)).astype(f32)).astype(f32)
    return Xn, Yn
