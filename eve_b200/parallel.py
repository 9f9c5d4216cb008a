"""Data-parallel training step of the EVE hot path: backward -> ONE gradient allreduce ->
fused clip + Adam on flat buffers.

Reference being replaced: the single-GPU step of src/core/training.py:485-502
(``loss.backward()``, ``clip_grad_norm_(model.parameters(), 5.0)``, ``optimizer.step()``) with
the optimizer of src/train.py:49-55 (``Adam(lr=config.learning_rate, weight_decay=...)``).
The reference has no multi-GPU code at all; clips are independent (per-sample InstanceNorm,
per-clip losses), so each rank runs the same model on its own clips and the only exchange
is one sum-allreduce of the flat fp32 gradient buffer between backward and clipping -- after
it every rank clips against the *global* gradient norm and applies the identical update,
exactly as one GPU with the concatenated batch would.
"""
import torch
import torch.distributed as dist

from . import ops
from .config import get_config


class FlatAdamTrainer(object):
    """Owns flat parameter / gradient / Adam-moment buffers; the model's parameters become
    views into the flat parameter buffer (names, shapes and state_dict keys unchanged)."""

    ALIGN = 64   # floats

    def __init__(self, model, lr=None, weight_decay=None, betas=(0.9, 0.999), eps=1e-8,
                 max_norm=None, process_group=None):
        cfg = get_config()
        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError('FlatAdamTrainer: the model has no trainable parameter')
        dev = self.params[0].device
        self.sizes = [p.numel() for p in self.params]
        # every parameter starts on a 256-byte boundary (the kernels read weights / norm gains
        # with 16-byte vector loads); the padding stays zero in all four buffers
        self.offsets, off = [], 0
        for n in self.sizes:
            self.offsets.append(off)
            off += (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        for p, o, n in zip(self.params, self.offsets, self.sizes):
            self.flat[o:o + n].copy_(p.detach().reshape(-1))
            p.data = self.flat[o:o + n].view(p.shape)
        self.grad = torch.zeros_like(self.flat)
        self.grad_views = [self.grad[o:o + n].view(p.shape)
                           for p, o, n in zip(self.params, self.offsets, self.sizes)]
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self._lr = float(cfg.learning_rate if lr is None else lr)
        # the kernel reads the rate from this device scalar: a scheduler (training.py:382-440)
        # may change `trainer.lr` between replays of a captured step
        self.lr_dev = torch.full((1,), self._lr, dtype=torch.float32, device=dev) \
            if dev.type == 'cuda' else None
        self.weight_decay = float(cfg.weight_decay if weight_decay is None else weight_decay)
        self.betas, self.eps = betas, eps
        if max_norm is None:
            if cfg.do_gradient_clipping and cfg.gradient_clip_by != 'norm':
                raise ValueError("gradient_clip_by='%s' is not supported by the fused step "
                                 "(only 'norm', training.py:492-498)" % cfg.gradient_clip_by)
            max_norm = cfg.gradient_clip_amount if cfg.do_gradient_clipping else 0.0
        self.max_norm = float(max_norm)
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.steps = 0
        # device-side copy of the step counter (what the fused kernel uses: graph-replayable)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev) if dev.type == 'cuda' else None
        self.device = dev
        self.last_grad_norm = None
        # every rank must start from the same parameters (what DDP's constructor enforces)
        if self.world > 1:
            dist.broadcast(self.flat, src=0, group=self.group)

    @property
    def lr(self):
        return self._lr

    @lr.setter
    def lr(self, value):
        self._lr = float(value)
        if self.lr_dev is not None:
            self.lr_dev.fill_(self._lr)      # stream-ordered: takes effect for the next step

    @property
    def param_groups(self):
        """Minimal torch.optim view for LR schedulers that write ``group['lr']``."""
        trainer = self

        class _Group(dict):
            def __setitem__(self, key, value):
                dict.__setitem__(self, key, value)
                if key == 'lr':
                    trainer.lr = value

        return [_Group(lr=self._lr, weight_decay=self.weight_decay, betas=self.betas, eps=self.eps,
                       params=self.params)]

    def state_dict(self):
        """Adam moments and step count (what CheckpointManager stores as optimizer_N.pt,
        checkpoint_manager.py:70-72), keyed by position like torch.optim."""
        return {'flat_exp_avg': self.exp_avg.detach().clone(),
                'flat_exp_avg_sq': self.exp_avg_sq.detach().clone(),
                'steps': int(self.steps), 'lr': self._lr, 'sizes': list(self.sizes)}

    def load_state_dict(self, state):
        if list(state['sizes']) != list(self.sizes):
            raise ValueError('FlatAdamTrainer.load_state_dict: parameter layout mismatch')
        self.exp_avg.copy_(state['flat_exp_avg'])
        self.exp_avg_sq.copy_(state['flat_exp_avg_sq'])
        self.steps = int(state['steps'])
        if self.step_dev is not None:
            self.step_dev.fill_(self.steps)
        self.lr = state.get('lr', self._lr)

    def gather_grads(self):
        """Pack p.grad of every trainable parameter into the flat gradient buffer (one fused
        multi-tensor copy); parameters that did not take part in the graph count as zero."""
        have = [(v, p.grad) for v, p in zip(self.grad_views, self.params) if p.grad is not None]
        missing = [v for v, p in zip(self.grad_views, self.params) if p.grad is None]
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        for v in missing:
            v.zero_()
        for p in self.params:
            p.grad = None

    def apply(self, update=None):
        """allreduce (sum) + clip by the global norm + Adam, on the flat buffers.

        ``update`` replaces the fused CUDA kernel (tests of the communication logic on a
        GPU-less box inject a torch restatement); the product path leaves it None and fails
        loudly on CPU tensors."""
        if self.world > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=self.group)
        self.steps += 1
        if update is not None:
            self.last_grad_norm = update(self)
            return
        self.last_grad_norm = ops.adam_clip_step(
            self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.steps, self._lr,
            self.betas, self.eps, self.weight_decay, self.max_norm, 1.0 / self.world,
            step_dev=self.step_dev, lr_dev=self.lr_dev)

    def step(self, loss, update=None):
        loss.backward()
        self.gather_grads()
        self.apply(update)
