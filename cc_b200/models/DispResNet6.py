"""DispResNet6 on the libccb200 kernels.  Reference: models/DispResNet6.py:97-194.

Same constructor, ``init_weights()``, forward arity, train/eval outputs and state_dict keys as the
reference; ReLU / sigmoid / residual-add run in the convolution epilogues."""
import torch
import torch.nn as nn
from .. import nn as cnn


class BasicBlock(nn.Module):
    """Two bias-free 3x3 convs + identity / (1x1 conv + BN) shortcut.  Reference DispResNet6.py:14-43."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = cnn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=False, act='relu')
        self.conv2 = cnn.Conv2d(planes, planes, 3, stride=1, padding=1, bias=False, act='relu')   # relu(conv + residual)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        residual = x if self.downsample is None else self.downsample(x)
        return self.conv2(self.conv1(x), res=residual)


def make_layer(inplanes, block, planes, blocks, stride=1):
    """Reference DispResNet6.py:45-60."""
    downsample = None
    if stride != 1 or inplanes != planes * block.expansion:
        downsample = nn.Sequential(cnn.Conv2d(inplanes, planes * block.expansion, 1, stride=stride, bias=False),
                                   cnn.BatchNorm2d(planes * block.expansion))
    layers = [block(inplanes, planes, stride, downsample)]
    for _ in range(1, blocks):
        layers.append(block(planes * block.expansion, planes))
    return nn.Sequential(*layers)


def downsample_conv(in_planes, out_planes, kernel_size=3):
    p = (kernel_size - 1) // 2
    return nn.Sequential(cnn.Conv2d(in_planes, out_planes, kernel_size, stride=2, padding=p, act='relu'), cnn.Fused(),
                         cnn.Conv2d(out_planes, out_planes, kernel_size, padding=p, act='relu'), cnn.Fused())


def predict_disp(in_planes):
    return nn.Sequential(cnn.Conv2d(in_planes, 1, 3, padding=1, act='sigmoid'), cnn.Fused('sigmoid'))


def upconv(in_planes, out_planes):
    return nn.Sequential(cnn.ConvTranspose2d(in_planes, out_planes, 3, stride=2, padding=1, output_padding=1, act='relu'),
                         cnn.Fused())


def crop_like(input, ref):
    assert(input.size(2) >= ref.size(2) and input.size(3) >= ref.size(3))
    return input[:, :, :ref.size(2), :ref.size(3)]


class DispResNet6(nn.Module):

    def __init__(self, alpha=10, beta=0.01):
        super().__init__()
        self.alpha = alpha
        self.beta = beta
        cp = [32, 64, 128, 256, 512, 512, 512]
        self.conv1 = downsample_conv(3, cp[0], kernel_size=7)
        for n in range(2, 8):
            setattr(self, 'conv%d' % n, make_layer(cp[n - 2], BasicBlock, cp[n - 1], blocks=2, stride=2))
        up = [512, 512, 256, 128, 64, 32, 16]
        ins = [cp[6]] + up[:6]
        for j, n in enumerate(range(7, 0, -1)):
            setattr(self, 'upconv%d' % n, upconv(ins[j], up[j]))
        iconv_in = {7: up[0] + cp[5], 6: up[1] + cp[4], 5: up[2] + cp[3], 4: up[3] + cp[2],
                    3: 1 + up[4] + cp[1], 2: 1 + up[5] + cp[0], 1: 1 + up[6]}
        for j, n in enumerate(range(7, 0, -1)):
            setattr(self, 'iconv%d' % n, make_layer(iconv_in[n], BasicBlock, up[j], blocks=1, stride=1))
        for n, c in zip(range(6, 0, -1), up[1:]):
            setattr(self, 'predict_disp%d' % n, predict_disp(c))

    def init_weights(self):
        cnn.xavier_init_(self)

    def _disp(self, n, feat):
        return self.alpha * getattr(self, 'predict_disp%d' % n)(feat) + self.beta

    def forward(self, x):
        c1 = self.conv1(x)
        c2 = self.conv2(c1)
        c3 = self.conv3(c2)
        c4 = self.conv4(c3)
        c5 = self.conv5(c4)
        c6 = self.conv6(c5)
        c7 = self.conv7(c6)

        i7 = self.iconv7(torch.cat((crop_like(self.upconv7(c7), c6), c6), 1))
        i6 = self.iconv6(torch.cat((crop_like(self.upconv6(i7), c5), c5), 1))
        disp6 = self._disp(6, i6)
        i5 = self.iconv5(torch.cat((crop_like(self.upconv5(i6), c4), c4), 1))
        disp5 = self._disp(5, i5)
        i4 = self.iconv4(torch.cat((crop_like(self.upconv4(i5), c3), c3), 1))
        disp4 = self._disp(4, i4)
        i3 = self.iconv3(torch.cat((crop_like(self.upconv3(i4), c2), c2, crop_like(cnn.upsample2x(disp4), c2)), 1))
        disp3 = self._disp(3, i3)
        i2 = self.iconv2(torch.cat((crop_like(self.upconv2(i3), c1), c1, crop_like(cnn.upsample2x(disp3), c1)), 1))
        disp2 = self._disp(2, i2)
        i1 = self.iconv1(torch.cat((crop_like(self.upconv1(i2), x), crop_like(cnn.upsample2x(disp2), x)), 1))
        disp1 = self._disp(1, i1)

        if self.training:
            return disp1, disp2, disp3, disp4, disp5, disp6
        return disp1
