"""Fused multi-tensor Adam over ONE flat fp32 buffer.

The reference builds a single torch.optim.Adam over chain(all nets' parameters)
(train.py:307-310: lr, betas=(momentum, beta), weight_decay 0) and, under nn.DataParallel,
broadcasts 297 MB of parameters and reduces 297 MB of gradients through GPU0 every step.  Here all
trainable parameters are views into one flat buffer, all gradients views into another; a step is one
`ccb_adam_step` launch, and data-parallel training all-reduces contiguous slices of the flat gradient
buffer (cc_b200/dist.py).

The order of the parameters inside the flat buffers is an internal detail (`relayout()` re-packs them in
gradient-completion order so that the data-parallel buckets are contiguous); checkpoints therefore use
torch.optim.Adam's own per-parameter `state_dict()` format, indexed by the order of the `params` argument
(= the reference's chain(disp, pose, mask, flow) order, train.py:307-310), and load either way."""
import torch
from . import _lib


class FlatAdam:
    def __init__(self, params, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        if weight_decay != 0:
            raise NotImplementedError('cc_b200.FlatAdam: weight_decay is 0 in the reference command line')
        self.params = [p for p in params if p.requires_grad]     # constructor order: the checkpoint index
        assert self.params, 'no trainable parameters'
        self.lr, self.betas, self.eps = lr, betas, eps
        self.numel = sum(p.numel() for p in self.params)
        self.grad_scale = 1.0
        self.flat_p = self.flat_g = self.exp_avg = self.exp_avg_sq = None
        self.state = torch.zeros(4, device=self.params[0].device, dtype=torch.float32)   # step, 1-b1^t, sqrt(1-b2^t)
        self._pack(list(self.params), None)

    # ---- flat layout -------------------------------------------------------------------------------
    def _pack(self, order, old):
        """(Re)build the flat buffers with the parameters in `order`; `old` = {param: (m, v)} state to carry over."""
        dev = order[0].device
        n = self.numel
        flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        exp_avg = torch.zeros(n, device=dev, dtype=torch.float32)
        exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        self.offset = {}
        with torch.no_grad():
            for p in order:
                k = p.numel()
                flat_p[off:off + k].copy_(p.data.reshape(-1))
                if old is not None:
                    exp_avg[off:off + k].copy_(old[p][0].reshape(-1))
                    exp_avg_sq[off:off + k].copy_(old[p][1].reshape(-1))
                    flat_g[off:off + k].copy_(old[p][2].reshape(-1))
                p.data = flat_p[off:off + k].view_as(p)
                gview = flat_g[off:off + k].view_as(p)
                p.grad = gview                 # torch-produced grads accumulate in place into the flat buffer
                p._ccb_grad = gview            # cc_b200.nn backward kernels write here directly
                p._ccb_written = False
                self.offset[p] = (off, k)
                off += k
        self.order = list(order)
        self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq = flat_p, flat_g, exp_avg, exp_avg_sq

    def _views(self, buf, p):
        off, k = self.offset[p]
        return buf[off:off + k].view_as(p)

    def relayout(self, order):
        """Re-pack the flat buffers with the parameters in `order` (a permutation of self.params), keeping
        values, gradients and Adam moments.  Used by the data-parallel bucket scheduler; must precede any CUDA-graph capture."""
        assert len(order) == len(self.params) and set(map(id, order)) == set(map(id, self.params))
        old = {p: (self._views(self.exp_avg, p).clone(), self._views(self.exp_avg_sq, p).clone(),
                   self._views(self.flat_g, p).clone()) for p in self.params}
        self._pack(list(order), old)

    # ---- step ----------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none=False):
        """`set_to_none` is accepted for torch compatibility and ignored: the gradients ARE the flat buffer."""
        self.flat_g.zero_()
        for p in self.params:
            p._ccb_written = False
            if p.grad is None:
                p.grad = p._ccb_grad

    def step(self):
        # Gradients are read from the flat buffer only.  If someone replaced p.grad (net.zero_grad(set_to_none=True)
        # followed by a torch-produced gradient installs a fresh tensor), fold that stray gradient in instead of
        # silently dropping it.
        for p in self.params:
            g = p.grad
            if g is not None and g.data_ptr() != p._ccb_grad.data_ptr():
                p._ccb_grad.add_(g)
                p.grad = p._ccb_grad
        _lib.check(_lib.lib().ccb_adam_step(_lib.ptr(self.flat_p), _lib.ptr(self.flat_g), _lib.ptr(self.exp_avg),
                                            _lib.ptr(self.exp_avg_sq), self.numel, _lib.ptr(self.state), self.lr,
                                            self.betas[0], self.betas[1], self.eps, self.grad_scale,
                                            _lib.stream(self.flat_p)), 'adam_step')

    # ---- checkpoint: torch.optim.Adam's state_dict format (reference utils.py:55-63 saves optimizer.state_dict()) ----
    def state_dict(self):
        step = torch.tensor(float(self.state[0].item()))
        state = {}
        if float(step) > 0:
            for i, p in enumerate(self.params):
                state[i] = {'step': step.clone(), 'exp_avg': self._views(self.exp_avg, p).detach().clone(),
                            'exp_avg_sq': self._views(self.exp_avg_sq, p).detach().clone()}
        group = {'lr': self.lr, 'betas': tuple(self.betas), 'eps': self.eps, 'weight_decay': 0, 'amsgrad': False,
                 'maximize': False, 'foreach': None, 'capturable': False, 'differentiable': False, 'fused': None,
                 'params': list(range(len(self.params)))}
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, sd):
        if 'flat' in sd:                       # round-1 format of this class
            assert self.order == self.params, 'flat optimizer checkpoints predate relayout()'
            self.exp_avg.copy_(sd['exp_avg'])
            self.exp_avg_sq.copy_(sd['exp_avg_sq'])
            step, g = sd['step'], sd
        else:
            g = sd['param_groups'][0]
            assert len(g['params']) == len(self.params), 'optimizer checkpoint has %d parameters, this model %d' % (
                len(g['params']), len(self.params))
            step = 0.0
            with torch.no_grad():
                for i, p in enumerate(self.params):
                    st = sd['state'].get(g['params'][i])
                    if st is None:
                        self._views(self.exp_avg, p).zero_()
                        self._views(self.exp_avg_sq, p).zero_()
                    else:
                        self._views(self.exp_avg, p).copy_(st['exp_avg'])
                        self._views(self.exp_avg_sq, p).copy_(st['exp_avg_sq'])
                        step = max(step, float(st['step']))
        self.state.zero_()
        self.state[0] = float(step)
        self.lr, self.betas, self.eps = g['lr'], tuple(g['betas']), g['eps']

    # ---- snapshot / restore (Trainer.capture warms up on real steps and must not train) --------------------------
    def snapshot(self):
        return (self.flat_p.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(), self.state.clone())

    def restore(self, snap):
        self.flat_p.copy_(snap[0]); self.exp_avg.copy_(snap[1]); self.exp_avg_sq.copy_(snap[2]); self.state.copy_(snap[3])
