"""Input pipeline on the device: the reference's training transform
    Compose([RandomHorizontalFlip(), RandomScaleCrop(), ArrayToTensor(), Normalize(.5, .5)])      (train.py:165-172)
applied by ONE kernel (csrc/io_ops.cu: prep_frames_kernel) to uint8 frames that cross PCIe as uint8 - 4x fewer H2D
bytes than the fp32 tensors the reference's DataLoader ships (train.py:448-451) - plus the matching intrinsics
update (custom_transforms.py:47-58,98-118) and K^-1 (datasets/sequence_folders.py:51-61).

The random parameters are drawn on the host with the reference's own generators and call order
(random.random(); np.random.uniform(1, 1.1, 2); np.random.randint(...) twice), so a seeded run makes the same
augmentation decisions as the reference loader.

Documented deviation: the reference resizes with scipy.misc.imresize (PIL BILINEAR on uint8: fixed-point, the
horizontally and vertically resampled images are each rounded back to uint8); here the same half-pixel-centre
bilinear lookup is evaluated in fp32 without re-quantisation, so a pixel differs from the reference by at most
one uint8 step per pass (<= 2/255 before normalisation); with scale 1 the result is identical."""
import ctypes as C
import random
import numpy as np
import torch
from . import _lib


def draw_params(B, Hs, Ws, H=None, W=None, rng_random=random, rng_np=np.random, flip=True, scale_crop=True):
    """Per-sample augmentation decisions, reference generators and order.  Returns a dict of numpy arrays."""
    H, W = H or Hs, W or Ws
    out = dict(flip=np.zeros(B, np.float32), x_scaling=np.ones(B), y_scaling=np.ones(B), scaled_h=np.full(B, Hs), scaled_w=np.full(B, Ws),
               offset_x=np.zeros(B, np.int32), offset_y=np.zeros(B, np.int32))
    for b in range(B):
        if flip and rng_random.random() < 0.5:                       # custom_transforms.py:52
            out['flip'][b] = 1.0
        if scale_crop:
            xs, ys = rng_np.uniform(1, 1.1, 2)                          # :107
            sh, sw = int(Hs * ys), int(Ws * xs)                          # :108
            out['x_scaling'][b], out['y_scaling'][b], out['scaled_h'][b], out['scaled_w'][b] = xs, ys, sh, sw
            out['offset_y'][b] = rng_np.randint(sh - H + 1)             # :117
            out['offset_x'][b] = rng_np.randint(sw - W + 1)             # :118
    return out


def augment_intrinsics(K, p, Ws):
    """K [B,3,3] float32 numpy -> augmented K (custom_transforms.py:55,110-111,122-123), same fp32 arithmetic."""
    K = np.array(K, dtype=np.float32, copy=True)
    for b in range(K.shape[0]):
        if p['flip'][b]:
            K[b, 0, 2] = Ws - K[b, 0, 2]
        K[b, 0] *= p['x_scaling'][b]
        K[b, 1] *= p['y_scaling'][b]
        K[b, 0, 2] -= p['offset_x'][b]
        K[b, 1, 2] -= p['offset_y'][b]
    return K


class DeviceAugment:
    """frames_u8 [B,F,Hs,Ws,3] uint8 (pinned host or device) + intrinsics [B,3,3] -> (tgt, refs, K, Kinv) on `device`."""

    def __init__(self, device, H=None, W=None, flip=True, scale_crop=True):
        self.device, self.H, self.W, self.flip, self.scale_crop = torch.device(device), H, W, flip, scale_crop

    def __call__(self, frames_u8, intrinsics, params=None, tgt_index=None):
        assert frames_u8.dtype == torch.uint8 and frames_u8.dim() == 5 and frames_u8.size(4) == 3
        B, F, Hs, Ws, _ = frames_u8.shape
        H, W = self.H or Hs, self.W or Ws
        p = params if params is not None else draw_params(B, Hs, Ws, H, W, flip=self.flip, scale_crop=self.scale_crop)
        src = frames_u8.to(self.device, non_blocking=True).contiguous()
        par = torch.from_numpy(np.stack([p['flip'], (p['scaled_w'] / Ws).astype(np.float32), (p['scaled_h'] / Hs).astype(np.float32),
                                         np.zeros(B, np.float32)], 1).astype(np.float32)).to(self.device, non_blocking=True)
        offs = torch.from_numpy(np.stack([p['offset_x'], p['offset_y']], 1).astype(np.int32)).to(self.device, non_blocking=True)
        outs = [torch.empty(B, 3, H, W, device=self.device) for _ in range(F)]
        arr = (C.c_void_p * F)(*[_lib.ptr(o) for o in outs])
        _lib.check(_lib.lib().ccb_prep_frames(_lib.ptr(src, 'frames', torch.uint8), arr, _lib.ptr(par), _lib.ptr(offs, 'offs', torch.int32), B, F, Hs, Ws, H, W,
                                              _lib.stream(src)), 'prep_frames')
        K = augment_intrinsics(intrinsics.cpu().numpy() if torch.is_tensor(intrinsics) else intrinsics, p, Ws)
        Kinv = np.linalg.inv(K).astype(np.float32)                     # sequence_folders.py:61
        t = F // 2 if tgt_index is None else tgt_index                # sequence_folders.py:16-21: the target is the middle frame
        refs = [o for i, o in enumerate(outs) if i != t]
        return outs[t], refs, torch.from_numpy(K).to(self.device), torch.from_numpy(Kinv).to(self.device)
