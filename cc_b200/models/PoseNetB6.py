"""PoseNetB6 on the libccb200 kernels.  Reference: models/PoseNetB6.py:24-83."""
import torch
import torch.nn as nn
from .. import nn as cnn


def conv(in_planes, out_planes, kernel_size=3):
    return nn.Sequential(cnn.Conv2d(in_planes, out_planes, kernel_size, stride=2, padding=(kernel_size - 1) // 2,
                                    act='relu'), cnn.Fused())


class PoseNetB6(nn.Module):

    def __init__(self, nb_ref_imgs=2):
        super().__init__()
        self.nb_ref_imgs = nb_ref_imgs
        planes = [16, 32, 64, 128, 256, 256, 256, 256]
        ks = [7, 5, 3, 3, 3, 3, 3, 3]
        cin = 3 * (1 + nb_ref_imgs)
        for i in range(8):
            setattr(self, 'conv%d' % (i + 1), conv(cin, planes[i], kernel_size=ks[i]))
            cin = planes[i]
        self.pose_pred = cnn.Conv2d(planes[7], 6 * nb_ref_imgs, 1, padding=0)

    def init_weights(self):
        cnn.xavier_init_(self)

    def forward(self, target_image, ref_imgs):
        assert(len(ref_imgs) == self.nb_ref_imgs)
        x = torch.cat([target_image] + list(ref_imgs), 1)
        for i in range(8):
            x = getattr(self, 'conv%d' % (i + 1))(x)
        pose = self.pose_pred(x)
        pose = pose.mean(3).mean(2)
        return 0.01 * pose.view(pose.size(0), self.nb_ref_imgs, 6)
