"""Fused multi-tensor Adam over ONE flat fp32 buffer.

The reference builds a single torch.optim.Adam over chain(all nets' parameters)
(train.py:307-310: lr, betas=(momentum, beta), weight_decay 0) and, under nn.DataParallel,
broadcasts 297 MB of parameters and reduces 297 MB of gradients through GPU0 every step.  Here all
trainable parameters are views into one flat buffer, all gradients views into another; a step is one
`ccb_adam_step` launch, and data-parallel training needs exactly one NCCL all-reduce of the flat
gradient buffer (cc_b200/dist.py)."""
import torch
from . import _lib


class FlatAdam:
    def __init__(self, params, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        if weight_decay != 0:
            raise NotImplementedError('cc_b200.FlatAdam: weight_decay is 0 in the reference command line')
        self.params = [p for p in params if p.requires_grad]
        assert self.params, 'no trainable parameters'
        dev = self.params[0].device
        self.lr, self.betas, self.eps = lr, betas, eps
        n = sum(p.numel() for p in self.params)
        self.numel = n
        self.flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32)
        self.state = torch.zeros(4, device=dev, dtype=torch.float32)     # step, 1-b1^t, sqrt(1-b2^t)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.data.reshape(-1))
                p.data = self.flat_p[off:off + k].view_as(p)
                gview = self.flat_g[off:off + k].view_as(p)
                p.grad = gview                 # torch-produced grads accumulate in place into the flat buffer
                p._ccb_grad = gview            # cc_b200.nn backward kernels write here directly
                p._ccb_written = False
                off += k
        self.grad_scale = 1.0

    def zero_grad(self, set_to_none=False):
        self.flat_g.zero_()
        for p in self.params:
            p._ccb_written = False
            if p.grad is None:
                p.grad = p._ccb_grad

    def step(self):
        _lib.check(_lib.lib().ccb_adam_step(_lib.ptr(self.flat_p), _lib.ptr(self.flat_g), _lib.ptr(self.exp_avg),
                                            _lib.ptr(self.exp_avg_sq), self.numel, _lib.ptr(self.state), self.lr,
                                            self.betas[0], self.betas[1], self.eps, self.grad_scale,
                                            _lib.stream(self.flat_p)), 'adam_step')

    # checkpoint contract of the reference: {'epoch', 'state_dict'} of the optimizer (utils.py:55-63)
    def state_dict(self):
        return {'flat': True, 'step': self.state[0].item(), 'exp_avg': self.exp_avg.clone(),
                'exp_avg_sq': self.exp_avg_sq.clone(), 'lr': self.lr, 'betas': self.betas, 'eps': self.eps}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd['exp_avg'])
        self.exp_avg_sq.copy_(sd['exp_avg_sq'])
        self.state.zero_()
        self.state[0] = sd['step']
        self.lr, self.betas, self.eps = sd['lr'], tuple(sd['betas']), sd['eps']
