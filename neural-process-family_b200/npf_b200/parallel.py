"""Data parallelism over tasks: every rank owns a contiguous shard of the meta-batch and a replica of the parameters;
the only exchange is ONE all-reduce of a flat fp32 gradient buffer per step.

``FlatGradients`` allocates one contiguous buffer and makes every ``p.grad`` a view into it, so the backward kernels'
results land in the bucket directly and the collective is a single NCCL call (0.5-2 MB: latency-bound on NVLink 5 /
NVSwitch).  Works with any ``torch.distributed`` backend (``gloo`` on CPU for the host-logic tests)."""
import torch
import torch.distributed as dist

__all__ = ["FlatGradients", "FlatAdam", "shard_tasks", "sync_batchnorm_", "sync_moments"]


class FlatGradients:
    def __init__(self, module, process_group=None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = process_group
        # every parameter's slice starts on a 128-byte boundary (vectorised / float4-atomic gradient kernels need 16 B)
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 31) // 32 * 32
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self._attach()

    def _attach(self):
        base = self.flat.data_ptr()
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off: off + p.numel()].view_as(p)
            # the weight-gradient kernels add straight into the bucket for THESE parameters only (ops._gbuf)
            p._npf_direct_grad = p.is_cuda
        self._grad_ptrs = [base + 4 * off for off in self.offsets]

    def detach(self):
        """Give the parameters back to plain autograd accumulation (drops the views into the bucket)."""
        for p in self.params:
            p._npf_direct_grad = False
            p.grad = None

    def zero_(self):
        """Zero the bucket (one memset) and re-attach the views if anything dropped or replaced one of them
        (``optimizer.zero_grad(set_to_none=True)``, ``model.zero_grad()``, a ``.to()`` of the module)."""
        self.flat.zero_()
        for p, ptr in zip(self.params, self._grad_ptrs):
            g = p.grad
            if g is None or g.data_ptr() != ptr:
                self._attach()
                break

    @property
    def world_size(self):
        if self.group is None and not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(self.group)

    def all_reduce_mean(self):
        """grad <- mean over ranks of the per-rank (local-batch-mean) gradients: equals the gradient of the global
        batch mean for equal shards.  No-op on a single rank."""
        w = self.world_size
        if w == 1:
            return self.flat
        import os
        if self.flat.is_cuda and os.environ.get("NPF_ALLREDUCE_AVG", "1") == "1":      # NCCL averages inside the collective: no separate 1/G kernel
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        else:                      # gloo (CPU host-logic tests) has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / w)
        return self.flat


class FlatAdam:
    """torch.optim.Adam (the optimizer of upstream utils/train.py:50) over the flat gradient bucket: the parameters are
    re-pointed into ONE contiguous fp32 buffer laid out like ``flat`` (same 128-byte aligned slots), so a step is a single
    elementwise kernel (``npf_adam_step``) over (param, grad, exp_avg, exp_avg_sq) instead of ~40 small foreach launches.
    ``lr`` may be changed between steps (``opt.lr = ...``: ExponentialLR is ``opt.lr *= gamma`` per epoch);
    ``grad_scale`` pre-multiplies the gradient (1 / world size); ``max_grad_norm`` clips by the global norm like
    ``torch.nn.utils.clip_grad_norm_`` (skorch ``GradientNormClipping`` of the latent notebooks) with the norm reduced and
    consumed on the device (``npf_sqnorm`` + ``npf_adam_step_clipped``): no host sync, capturable in a CUDA graph."""

    def __init__(self, flat, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.flat, self.lr, self.betas, self.eps, self.weight_decay = flat, lr, betas, eps, weight_decay
        self.step_count = 0
        g = flat.flat
        self.param = torch.zeros_like(g)
        with torch.no_grad():
            for p, off in zip(flat.params, flat.offsets):
                view = self.param[off: off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view                     # the module now reads / the kernels now update the flat storage
        self.exp_avg = torch.zeros_like(g)
        self.exp_avg_sq = torch.zeros_like(g)
        self._sqnorm = torch.zeros(1, dtype=torch.float32, device=g.device)

    def last_grad_norm(self):
        """Global gradient norm seen by the last clipped step (device scalar; reading it synchronises)."""
        return self._sqnorm.sqrt()

    def global_grad_norm(self):
        return self.flat.flat.norm()              # padding between slots is zero

    def step(self, grad_scale=1.0, max_grad_norm=None):
        from . import _cabi
        if not self.param.is_cuda:
            raise RuntimeError("FlatAdam: CUDA only (there is no CPU fallback)")
        self.step_count += 1
        n = self.param.numel()
        if max_grad_norm is not None:
            st = torch.cuda.current_stream().cuda_stream
            _cabi.call("npf_sqnorm", self.flat.flat.data_ptr(), n, self._sqnorm.data_ptr(), st)
            _cabi.call("npf_adam_step_clipped", self.param.data_ptr(), self.flat.flat.data_ptr(), self.exp_avg.data_ptr(),
                       self.exp_avg_sq.data_ptr(), n, self.step_count, float(self.lr), float(self.betas[0]), float(self.betas[1]),
                       float(self.eps), float(self.weight_decay), float(grad_scale), self._sqnorm.data_ptr(), float(max_grad_norm), st)
            return
        if self.param.is_cuda:
            _cabi.call("npf_adam_step", self.param.data_ptr(), self.flat.flat.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), n,
                       self.step_count, float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay),
                       float(grad_scale), torch.cuda.current_stream().cuda_stream)
            return
        raise RuntimeError("FlatAdam: CUDA only (there is no CPU fallback)")

    def state_dict(self):
        return dict(step=self.step_count, lr=self.lr, exp_avg=self.exp_avg.clone(), exp_avg_sq=self.exp_avg_sq.clone())

    # ---- interchange with torch.optim.Adam (the ``optimizer.pt`` upstream's skorch Checkpoint writes next to ``params.pt``)
    def torch_state_dict(self):
        """The state in ``torch.optim.Adam.state_dict()`` layout (per-parameter moments, index keys, one param group)."""
        state = {}
        for i, (p, off) in enumerate(zip(self.flat.params, self.flat.offsets)):
            sl = slice(off, off + p.numel())
            state[i] = dict(step=torch.tensor(float(self.step_count)), exp_avg=self.exp_avg[sl].view_as(p).detach().cpu().clone(),
                            exp_avg_sq=self.exp_avg_sq[sl].view_as(p).detach().cpu().clone())
        group = dict(lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=self.weight_decay, amsgrad=False,
                     params=list(range(len(self.flat.params))))
        return dict(state=state, param_groups=[group])

    def load_torch_state_dict(self, sd):
        """Resume from a ``torch.optim.Adam`` state dict -- current index-keyed layout or the legacy id-keyed one of the
        upstream checkpoints: the i-th entry of ``param_groups[0]['params']`` belongs to the i-th parameter."""
        groups = sd["param_groups"]
        if len(groups) != 1:
            raise ValueError("FlatAdam.load_torch_state_dict: exactly one param group expected")
        g, keys = groups[0], groups[0]["params"]
        if len(keys) != len(self.flat.params):
            raise ValueError(f"optimizer state has {len(keys)} parameters, the model {len(self.flat.params)}")
        if g.get("amsgrad", False):
            raise NotImplementedError("FlatAdam: amsgrad state is not supported")
        steps = set()
        for key, p, off in zip(keys, self.flat.params, self.flat.offsets):
            st = sd["state"].get(key)
            if st is None:        # parameter never stepped
                continue
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError(f"optimizer state shape {tuple(st['exp_avg'].shape)} does not match parameter {tuple(p.shape)}")
            sl = slice(off, off + p.numel())
            self.exp_avg[sl].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[sl].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(st["step"]))
        if len(steps) > 1:
            raise ValueError(f"FlatAdam keeps one step count for all parameters, the state has {sorted(steps)}")
        self.step_count = steps.pop() if steps else 0
        self.lr, self.betas = float(g["lr"]), tuple(g["betas"])
        self.eps, self.weight_decay = float(g["eps"]), float(g["weight_decay"])

    def load_state_dict(self, sd):
        self.step_count, self.lr = int(sd["step"]), float(sd["lr"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])


class _AllReduceSum(torch.autograd.Function):
    """Differentiable sum-all-reduce: the gradient of a sum over ranks is the sum over ranks of the gradients."""

    @staticmethod
    def forward(ctx, t, group):
        ctx.group = group
        out = t.detach().clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None


def sync_moments(mean, var, n_local, group=None):
    """Per-rank batch statistics (mean, biased var over ``n_local`` positions) -> statistics of the GLOBAL batch, by ONE
    all-reduce of the 2C + 1 numbers (sum x, sum x^2, count) per BatchNorm layer in forward and one of their gradients in
    backward (SURVEY.md section 8e caveat).  Returns (mean, var, n_total); differentiable."""
    C = mean.numel()
    packed = torch.cat([mean * n_local, (var + mean * mean) * n_local, mean.new_full((1,), float(n_local))])
    tot = _AllReduceSum.apply(packed, group)
    n = tot[2 * C]
    g_mean = tot[:C] / n
    g_var = tot[C:2 * C] / n - g_mean * g_mean
    return g_mean, g_var, n


def sync_batchnorm_(module, group=None):
    """Mark every BatchNorm of ``module`` (the ``Normalization=nn.BatchNorm*`` CNNs of the notebooks) as synchronised over
    ``group``: in train mode the folded pre-activation affine then uses the statistics of the global meta-batch, which
    makes an N-rank run equal to the single-process reference on the same global batch (without it every rank normalises
    with its own shard: "replicas with local BN").  In place; returns the module.  Eval mode is unaffected."""
    for m in module.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m._npf_sync_group = (group,)
    return module


def shard_tasks(batch, rank, world_size):
    """Contiguous split of the task axis: rank g gets tasks [g*B/G, (g+1)*B/G).  ``batch`` is a dict of tensors whose
    first axis is the task axis; B must be divisible by the world size."""
    out = {}
    for k, v in batch.items():
        B = v.shape[0]
        if B % world_size:
            raise ValueError(f"meta-batch {B} is not divisible by world size {world_size}")
        s = B // world_size
        out[k] = v[rank * s: (rank + 1) * s]
    return out
