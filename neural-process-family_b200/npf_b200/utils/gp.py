"""Synthetic Gaussian-process tasks generated ON THE DEVICE.

Role and sampling scheme of upstream ``GPDataset`` (utils/data/gaussian_process.py:25-231): positions
``X ~ U(min, max)`` sorted per task (189-194), targets ``y ~ N(0, k(X, X))`` for a scikit-learn kernel object
(201-231, ``n_same_samples`` draws share one set of positions), positions rescaled to [-1, 1] (196-199), tasks shuffled.
The covariance is factorised per task by ``npf_gp_sample`` (pivoted Cholesky in shared memory); nothing touches the host.
The kernel argument is a scikit-learn kernel object exactly as upstream passes it (``RBF``, ``Matern(nu=1.5)``,
``ExpSineSquared``, optionally ``WhiteKernel + .``) -- read by attribute, scikit-learn itself is not imported -- or the
equivalent ``dict(kind=..., length_scale=..., periodicity=..., noise_level=...)``.

``tol`` (default 1e-5) is the residual variance at which the factorisation stops: an order of magnitude above the fp32
rounding noise of the residual diagonal (~ sqrt(rank) * 6e-8), two below the smallest predictive variance the models can
express (min sigma 0.01).

CUDA only; randomness from ``torch``'s CUDA generator (seed with ``torch.manual_seed``)."""
import torch

from .. import _cabi

__all__ = ["GPSampler", "kernel_hyperparameters", "kernel_hyperparameter_bounds"]

KINDS = {"rbf": 0, "matern15": 1, "periodic": 2}


def kernel_hyperparameters(kernel):
    """scikit-learn kernel object (or dict) -> dict(kind, length_scale, periodicity, noise_level)."""
    if isinstance(kernel, dict):
        out = dict(kind=kernel["kind"], length_scale=float(kernel["length_scale"]), periodicity=float(kernel.get("periodicity", 1.0)),
                   noise_level=float(kernel.get("noise_level", 0.0)))
        if out["kind"] not in KINDS:
            raise ValueError(f"unknown kernel kind {out['kind']}")
        return out
    name = type(kernel).__name__
    if name == "Sum":  # WhiteKernel(noise) + stationary kernel, either order
        parts = [kernel.k1, kernel.k2]
        white = [k for k in parts if type(k).__name__ == "WhiteKernel"]
        rest = [k for k in parts if type(k).__name__ != "WhiteKernel"]
        if len(white) != 1 or len(rest) != 1:
            raise NotImplementedError("npf_b200.GPSampler: only `WhiteKernel + <stationary kernel>` sums are implemented")
        out = kernel_hyperparameters(rest[0])
        out["noise_level"] = float(white[0].noise_level)
        return out
    if name == "RBF":
        return dict(kind="rbf", length_scale=float(kernel.length_scale), periodicity=1.0, noise_level=0.0)
    if name == "Matern":
        if float(kernel.nu) != 1.5:
            raise NotImplementedError("npf_b200.GPSampler: Matern kernels with nu != 1.5 are not implemented (no upstream dataset uses them)")
        return dict(kind="matern15", length_scale=float(kernel.length_scale), periodicity=1.0, noise_level=0.0)
    if name == "ExpSineSquared":
        return dict(kind="periodic", length_scale=float(kernel.length_scale), periodicity=float(kernel.periodicity), noise_level=0.0)
    raise NotImplementedError(f"npf_b200.GPSampler: kernel {name} is not implemented")


def kernel_hyperparameter_bounds(kernel):
    """(lo, hi) per hyper-parameter name for the ones a `is_vary_kernel_hyp` dataset resamples: the kernel's own
    ``<name>_bounds`` (scikit-learn attribute or dict key); "fixed" / absent bounds leave that hyper-parameter at the kernel's value
    (upstream draws ``uniform(*hyperparam.bounds)`` for every entry of ``kernel.hyperparameters``, gaussian_process.py:239-242)."""
    out = {}

    def take(obj, name):
        b = obj.get(name + "_bounds") if isinstance(obj, dict) else getattr(obj, name + "_bounds", None)
        if b is None or (isinstance(b, str) and b == "fixed"):
            return
        lo, hi = (float(v) for v in b)
        if not (0 < lo <= hi):
            raise ValueError(f"{name}_bounds must satisfy 0 < lo <= hi, got {b}")
        out[name] = (lo, hi)

    if isinstance(kernel, dict):
        for name in ("length_scale", "periodicity", "noise_level"):
            take(kernel, name)
        return out
    name = type(kernel).__name__
    if name == "Sum":
        for part in (kernel.k1, kernel.k2):
            out.update(kernel_hyperparameter_bounds(part))
        return out
    if name == "WhiteKernel":
        take(kernel, "noise_level")
    else:
        take(kernel, "length_scale")
        if name == "ExpSineSquared":
            take(kernel, "periodicity")
    return out


class GPSampler:
    """``is_vary_kernel_hyp=True`` (upstream gaussian_process.py:39-41, 206-207, 233-242): every group of ``n_same_samples`` functions
    comes from a kernel whose hyper-parameters were drawn uniformly in their bounds -- drawn here on the device, one row of
    (length_scale, periodicity, noise_level) per group, and consumed by ``npf_gp_sample_hyp``."""

    def __init__(self, kernel, min_max=(-2, 2), n_points=128, n_same_samples=20, tol=1e-5, device="cuda", is_vary_kernel_hyp=False):
        self.hyp = kernel_hyperparameters(kernel)
        self.is_vary_kernel_hyp = bool(is_vary_kernel_hyp)
        self.bounds = kernel_hyperparameter_bounds(kernel) if self.is_vary_kernel_hyp else {}
        self.min_max, self.n_points, self.n_same_samples, self.tol = min_max, n_points, n_same_samples, tol
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("npf_b200.GPSampler runs on CUDA devices only (there is no CPU fallback)")

    def sample_hyperparameters(self, T):
        """[T, 3] device rows (length_scale, periodicity, noise_level): U(lo, hi) where bounds exist, the kernel's value otherwise."""
        h = self.hyp
        cols = []
        for name in ("length_scale", "periodicity", "noise_level"):
            if name in self.bounds and not (name == "periodicity" and h["kind"] != "periodic"):
                lo, hi = self.bounds[name]
                cols.append(torch.rand(T, device=self.device) * (hi - lo) + lo)
            else:
                cols.append(torch.full((T,), h[name], device=self.device))
        return torch.stack(cols, dim=1).contiguous()

    def sample_targets(self, X, n_same_samples=1, eps=None, return_factor=False, hyp=None):
        """X [T, N] raw positions -> Y [T, n_same_samples, N] (and the factor L [T, N, N], rank [T] when asked).  ``hyp`` [T, 3]:
        per-task (length_scale, periodicity, noise_level); default: drawn when ``is_vary_kernel_hyp``, else the kernel's own."""
        X = X.to(self.device, torch.float32).contiguous()
        T, N = X.shape
        if hyp is None and self.is_vary_kernel_hyp:
            hyp = self.sample_hyperparameters(T)
        if hyp is not None:
            hyp = hyp.to(self.device, torch.float32).contiguous()
            assert tuple(hyp.shape) == (T, 3), f"hyp must be [T, 3], got {tuple(hyp.shape)}"
        S = n_same_samples
        if eps is None:
            eps = torch.randn(T, S, N, device=self.device, dtype=torch.float32)
        eps = eps.to(self.device, torch.float32).contiguous()
        assert tuple(eps.shape) == (T, S, N)
        Y = torch.empty(T, S, N, device=self.device, dtype=torch.float32)
        L = torch.empty(T, N, N, device=self.device, dtype=torch.float32) if return_factor else None
        rank = torch.empty(T, device=self.device, dtype=torch.int32) if return_factor else None
        h = self.hyp
        with torch.cuda.device(self.device):
            if hyp is not None:
                _cabi.call("npf_gp_sample_hyp", X.data_ptr(), eps.data_ptr(), Y.data_ptr(), None if L is None else L.data_ptr(),
                           None if rank is None else rank.data_ptr(), hyp.data_ptr(), T, N, S, KINDS[h["kind"]], float(self.tol),
                           torch.cuda.current_stream().cuda_stream)
                return (Y, L, rank) if return_factor else Y
            _cabi.call("npf_gp_sample", X.data_ptr(), eps.data_ptr(), Y.data_ptr(), None if L is None else L.data_ptr(),
                       None if rank is None else rank.data_ptr(), T, N, S, KINDS[h["kind"]], h["length_scale"], h["periodicity"],
                       h["noise_level"], float(self.tol), torch.cuda.current_stream().cuda_stream)
        return (Y, L, rank) if return_factor else Y

    def get_samples(self, n_samples, n_points=None, test_min_max=None):
        """(data [n_samples, n_points, 1] in [-1, 1], targets [n_samples, n_points, 1]) like upstream ``get_samples``."""
        n_points = self.n_points if n_points is None else n_points
        lo, hi = self.min_max if test_min_max is None else test_min_max
        S = max(1, min(self.n_same_samples, n_samples))
        T = -(-n_samples // S)
        X = torch.rand(T, n_points, device=self.device) * (hi - lo) + lo
        X = X.sort(dim=-1).values
        Y = self.sample_targets(X, S)                                         # [T, S, N]
        Xr = X.unsqueeze(1).expand(T, S, n_points).reshape(T * S, n_points)[:n_samples]
        Y = Y.reshape(T * S, n_points)[:n_samples]
        perm = torch.randperm(n_samples, device=self.device)                  # not n_same_samples consecutive look-alikes
        a, b = self.min_max                                                   # rescale_range(X, min_max, (-1, 1)), 196-199
        Xr = (Xr[perm] - a) / (b - a) * 2 - 1
        return Xr.unsqueeze(-1).contiguous(), Y[perm].unsqueeze(-1).contiguous()
