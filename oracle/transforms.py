"""Oracle: the reference's training transform on uint8 frames (``custom_transforms.py:21-30,47-118`` as composed at
``train.py:165-172``).  TEST INFRASTRUCTURE.

numpy restatement of Compose([RandomHorizontalFlip, RandomScaleCrop, ArrayToTensor, Normalize(.5,.5)]) given the
random decisions.  The resize step restates scipy.misc.imresize's algorithm (PIL BILINEAR: half-pixel-centre triangle
filter, support 1 when up-scaling) in float64 WITHOUT PIL's two uint8 re-quantisations; the fixture frozen from the
reference (tests/golden/transforms_small.npz, made with the real PIL resize) therefore pins it to within one uint8
step per resampling pass, and exactly when no resize happens."""
import numpy as np


def _resize_bilinear(img, sh, sw):
    """img [H,W,3] uint8 -> float64 [sh,sw,3]; sample position (dst + 0.5) * in/out - 0.5, clamped to the frame."""
    H, W, _ = img.shape
    ys = np.clip((np.arange(sh) + 0.5) * (H / sh) - 0.5, 0, H - 1)
    xs = np.clip((np.arange(sw) + 0.5) * (W / sw) - 0.5, 0, W - 1)
    y0, x0 = np.floor(ys).astype(int), np.floor(xs).astype(int)
    y1, x1 = np.minimum(y0 + 1, H - 1), np.minimum(x0 + 1, W - 1)
    wy, wx = (ys - y0)[:, None, None], (xs - x0)[None, :, None]
    f = img.astype(np.float64)
    top = (1 - wx) * f[y0][:, x0] + wx * f[y0][:, x1]
    bot = (1 - wx) * f[y1][:, x0] + wx * f[y1][:, x1]
    return (1 - wy) * top + wy * bot


def apply(frames, K, p, H=None, W=None):
    """frames [B,F,Hs,Ws,3] uint8, K [3,3] or [B,3,3]; p = cc_b200.input_pipeline.draw_params(...) -> (out [B,F,3,H,W] f32, K [B,3,3])."""
    B, F, Hs, Ws, _ = frames.shape
    H, W = H or Hs, W or Ws
    Kb = np.broadcast_to(np.asarray(K, np.float32), (B, 3, 3)).copy()
    out = np.zeros((B, F, 3, H, W), np.float32)
    for b in range(B):
        for f in range(F):
            im = frames[b, f]
            if p['flip'][b]:
                im = np.fliplr(im)
            sh, sw = int(p['scaled_h'][b]), int(p['scaled_w'][b])
            r = _resize_bilinear(im, sh, sw) if (sh, sw) != (Hs, Ws) else im.astype(np.float64)
            oy, ox = int(p['offset_y'][b]), int(p['offset_x'][b])
            c = r[oy:oy + H, ox:ox + W]
            out[b, f] = ((np.transpose(c, (2, 0, 1)).astype(np.float32) / 255) - 0.5) / 0.5
        if p['flip'][b]:
            Kb[b, 0, 2] = Ws - Kb[b, 0, 2]
        Kb[b, 0] *= p['x_scaling'][b]
        Kb[b, 1] *= p['y_scaling'][b]
        Kb[b, 0, 2] -= p['offset_x'][b]
        Kb[b, 1, 2] -= p['offset_y'][b]
    return out, Kb
