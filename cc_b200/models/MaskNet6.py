"""MaskNet6 on the libccb200 kernels.  Reference: models/MaskNet6.py:19-123."""
import torch
import torch.nn as nn
from .. import nn as cnn


def conv(in_planes, out_planes, kernel_size=3):
    return nn.Sequential(cnn.Conv2d(in_planes, out_planes, kernel_size, stride=2, padding=(kernel_size - 1) // 2,
                                    act='relu'), cnn.Fused())


def upconv(in_planes, out_planes):
    return nn.Sequential(cnn.ConvTranspose2d(in_planes, out_planes, 4, stride=2, padding=1, act='relu'), cnn.Fused())


class MaskNet6(nn.Module):

    def __init__(self, nb_ref_imgs=4, output_exp=True):
        super().__init__()
        self.nb_ref_imgs = nb_ref_imgs
        self.output_exp = output_exp
        planes = [16, 32, 64, 128, 256, 256]
        ks = [7, 5, 3, 3, 3, 3]
        cin = 3 * (1 + nb_ref_imgs)
        for i in range(6):
            setattr(self, 'conv%d' % (i + 1), conv(cin, planes[i], kernel_size=ks[i]))
            cin = planes[i]
        if self.output_exp:
            up = [256, 256, 128, 64, 32, 16]
            dins = [planes[5], up[0] + planes[4], up[1] + planes[3], up[2] + planes[2], up[3] + planes[1], up[4] + planes[0]]
            for j, n in enumerate(range(6, 0, -1)):
                setattr(self, 'deconv%d' % n, upconv(dins[j], up[j]))
            for j, n in enumerate(range(6, 0, -1)):
                # sigmoid (applied functionally in the reference, MaskNet6.py:104-109) is fused into the head
                setattr(self, 'pred_mask%d' % n, cnn.Conv2d(up[j], nb_ref_imgs, 3, padding=1, act='sigmoid'))

    def init_weights(self):
        cnn.xavier_init_(self)

    def init_mask_weights(self):
        """Reference MaskNet6.py:60-73: re-initialise only the decoder + heads."""
        for m in self.modules():
            if isinstance(m, cnn.ConvTranspose2d):
                nn.init.xavier_uniform_(m.weight.data)
                if m.bias is not None:
                    m.bias.data.zero_()
        for n in range(1, 7):
            m = getattr(self, 'pred_mask%d' % n)
            nn.init.xavier_uniform_(m.weight.data)
            m.bias.data.zero_()

    def forward(self, target_image, ref_imgs):
        assert(len(ref_imgs) == self.nb_ref_imgs)
        x = torch.cat([target_image] + list(ref_imgs), 1)
        enc = []
        for i in range(6):
            x = getattr(self, 'conv%d' % (i + 1))(x)
            enc.append(x)
        if not self.output_exp:
            return (None,) * 6 if self.training else None
        ups = [self.deconv6(enc[5])]
        for n in range(5, 0, -1):
            ups.append(getattr(self, 'deconv%d' % n)(torch.cat((ups[-1], enc[n - 1]), 1)))
        masks = [getattr(self, 'pred_mask%d' % n)(ups[6 - n]) for n in range(1, 7)]
        if self.training:
            return tuple(masks)
        return masks[0]
