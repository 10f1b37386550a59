"""Back2Future (multi-frame PWC-style flow net) on the libccb200 kernels.
Reference: models/back2future.py:51-321.  Same constructor (`Model(nlevels)`), `init_weights()`, forward
arity `flow(tgt, [ref-, ref+])`, train-mode `(flow_fwd[6], flow_bwd[6], occ[6])` / eval-mode triple and
state_dict keys (`conv1a.0.weight`, `decoder_fwd6.10.bias`, ...).

`compute_occ=False` skips the five occlusion decoders: in the reference's train step their outputs are
discarded (`flow_fwd, flow_bwd, _ = flow_net(...)`, train.py:463) and they receive no gradient
(SURVEY.md F9), so skipping them changes no result."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from .. import nn as cnn


def conv_feat_block(nIn, nOut):
    return nn.Sequential(cnn.Conv2d(nIn, nOut, 3, stride=2, padding=1, act='leaky', slope=0.2), cnn.Fused('leaky'),
                         cnn.Conv2d(nOut, nOut, 3, stride=1, padding=1, act='leaky', slope=0.2), cnn.Fused('leaky'))


def conv_dec_block(nIn):
    chans = [nIn, 128, 128, 96, 64, 32]
    layers = []
    for i in range(5):
        layers += [cnn.Conv2d(chans[i], chans[i + 1], 3, padding=1, act='leaky', slope=0.2), cnn.Fused('leaky')]
    layers.append(cnn.Conv2d(32, 2, 3, padding=1))
    return nn.Sequential(*layers)


class Model(nn.Module):
    def __init__(self, nlevels, compute_occ=True):
        super().__init__()
        self.nlevels = nlevels
        self.compute_occ = compute_occ
        feat = [(3, 16), (16, 32), (32, 64), (64, 96), (96, 128), (128, 192)]
        for lvl, (ci, co) in enumerate(feat, start=1):
            for tower in 'abc':
                setattr(self, 'conv%d%s' % (lvl, tower), conv_feat_block(ci, co))
        dec_in = {6: 162, 5: 292, 4: 260, 3: 228, 2: 196}
        for lvl in range(6, 1, -1):
            setattr(self, 'decoder_fwd%d' % lvl, conv_dec_block(dec_in[lvl]))
            setattr(self, 'decoder_bwd%d' % lvl, conv_dec_block(dec_in[lvl]))
        for lvl in range(6, 1, -1):
            setattr(self, 'decoder_occ%d' % lvl, conv_dec_block(354 if lvl == 6 else dec_in[lvl]))
        # ImageNet statistics of normalize(): non-persistent buffers (no state_dict key, follow .cuda(); a tensor built
        # from a Python list inside forward would be a pageable H2D copy, which a CUDA-graph capture refuses)
        self.register_buffer('_norm_mean', torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1), persistent=False)
        self.register_buffer('_norm_std', torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1), persistent=False)

    def init_weights(self):
        cnn.xavier_init_(self, bias_uniform=True)

    def normalize(self, ims):
        """Reference back2future.py:118-132 ([-1,1] frames -> ImageNet-normalised)."""
        mean, std = self._norm_mean, self._norm_std
        if mean.device != ims[0].device:                 # module moved by hand / parameters only: follow the frames
            mean, std = mean.to(ims[0].device), std.to(ims[0].device)
        return [(im * 0.5 + 0.5 - mean) / std for im in ims]

    def warp(self, x, flo):
        return cnn.feat_warp(x, flo)

    def forward(self, im_tar, im_refs):
        ims = self.normalize([im_tar] + list(im_refs))
        fa, fb, fc = [ims[0]], [ims[2]], [ims[1]]          # a: I_0, b: I_+, c: I_-
        for lvl in range(1, 7):
            fa.append(getattr(self, 'conv%da' % lvl)(fa[-1]))
            fb.append(getattr(self, 'conv%db' % lvl)(fb[-1]))
            fc.append(getattr(self, 'conv%dc' % lvl)(fc[-1]))
        scale = {5: 0.625, 4: 1.25, 3: 2.5, 2: 5.0}
        flows_f, flows_b, ups_f, ups_b, occs = {}, {}, {}, {}, {}
        up_f = up_b = None
        for lvl in range(6, 1, -1):
            if lvl == 6:
                b_feat, c_feat = fb[6], fc[6]
            else:
                b_feat = self.warp(fb[lvl], scale[lvl] * up_f)
                c_feat = self.warp(fc[lvl], -scale[lvl] * up_f)
            corr = torch.cat((cnn.corr81(fa[lvl], b_feat, False), cnn.corr81(fa[lvl], c_feat, True)), 1)
            if lvl == 6:
                in_f = in_b = corr
                in_o = torch.cat((corr, fa[6]), 1) if self.compute_occ else None
            else:
                in_f = torch.cat((corr, fa[lvl], up_f), 1)
                in_b = torch.cat((corr, fa[lvl], up_b), 1)
                in_o = in_f
            flows_f[lvl] = getattr(self, 'decoder_fwd%d' % lvl)(in_f)
            flows_b[lvl] = getattr(self, 'decoder_bwd%d' % lvl)(in_b)
            if self.compute_occ:
                occs[lvl] = F.softmax(getattr(self, 'decoder_occ%d' % lvl)(in_o), dim=1)
            up_f, up_b = cnn.upsample2x(flows_f[lvl]), cnn.upsample2x(flows_b[lvl])
            ups_f[lvl], ups_b[lvl] = up_f, up_b
        mult = {2: 20.0, 3: 10.0, 4: 5.0, 5: 2.5, 6: 1.25}
        if not self.training:
            occ_full = F.interpolate(occs[2], scale_factor=4, mode='nearest') if self.compute_occ else None
            return mult[2] * cnn.upsample2x(ups_f[2]), -mult[2] * cnn.upsample2x(ups_b[2]), occ_full
        flow_fwd = [mult[l] * cnn.upsample2x(ups_f[l]) for l in range(2, 7)]
        flow_bwd = [-mult[l] * cnn.upsample2x(ups_b[l]) for l in range(2, 7)]
        occ = [F.interpolate(occs[l], scale_factor=4, mode='nearest') for l in range(2, 7)] if self.compute_occ else None
        if self.nlevels == 6:
            flow_fwd.append(0.625 * ups_f[6])
            flow_bwd.append(-0.625 * ups_b[6])
            if self.compute_occ:
                occ.append(F.interpolate(occs[6], scale_factor=2, mode='nearest'))
        return flow_fwd, flow_bwd, occ
