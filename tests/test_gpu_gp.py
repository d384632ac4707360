"""-m gpu: the device Gaussian-process sampler (npf_gp_sample behind npf_b200.utils.gp.GPSampler).  A sampler is pinned
by its LAW, not by a stream: the factor it samples with must reproduce the covariance scikit-learn builds for the same
kernel and positions (tests/golden/gp/kernels.npz), y must be exactly L eps, and the empirical covariance of many draws
must converge to it.  Tolerances: |L L^T - K| <= 5e-5 (fp32 factorisation of matrices with entries <= 1.1, stopping
tolerance 1e-5); the fp64 oracle's rank +- 3 (the stopping test sits on a steep eigenvalue decay)."""
import os

import numpy as np
import pytest
import torch

from _util import ROOT
from oracle import gp_oracle as G

pytestmark = pytest.mark.gpu
FIX = np.load(os.path.join(ROOT, "tests", "golden", "gp", "kernels.npz"))
NAMES = sorted(k[:-2] for k in FIX.files if k.endswith("_x"))
KIND = {0: "rbf", 1: "matern15", 2: "periodic"}


def _sampler(name):
    from npf_b200.utils.gp import GPSampler
    kind, ls, per, noise = FIX[name + "_hyp"]
    return GPSampler(dict(kind=KIND[int(kind)], length_scale=ls, periodicity=per, noise_level=noise))


@pytest.mark.parametrize("name", NAMES)
def test_factor_reproduces_sklearn_covariance(name):
    x, K = FIX[name + "_x"], FIX[name + "_K"]
    N = len(x)
    s = _sampler(name)
    X = torch.from_numpy(np.stack([x, x[::-1].copy()])).float()          # second task: same points, reversed order
    eps = torch.randn(2, 5, N, generator=torch.Generator().manual_seed(0))
    Y, L, rank = s.sample_targets(X, 5, eps=eps, return_factor=True)
    L64, Y64 = L.double().cpu().numpy(), Y.double().cpu().numpy()
    assert np.abs(L64[0] @ L64[0].T - K).max() < 5e-5
    assert np.abs(L64[1] @ L64[1].T - K[::-1, ::-1]).max() < 5e-5
    _, r_ref = G.pivoted_cholesky(K, s.tol)
    assert abs(int(rank[0]) - r_ref) <= 3 and (L64[0][:, int(rank[0]):] == 0).all()
    want = np.einsum("tik,tsk->tsi", L64, eps.double().numpy())
    assert np.abs(Y64 - want).max() < 1e-5 * max(1.0, np.abs(want).max())


def test_empirical_covariance_converges():
    name = "noisy_matern"
    x, K = FIX[name + "_x"][:64], FIX[name + "_K"][:64, :64]
    s = _sampler(name)
    torch.manual_seed(0)
    S = 20000
    Y = s.sample_targets(torch.from_numpy(x).float()[None], S)[0].double().cpu().numpy()   # [S, 64]
    C = Y.T @ Y / S
    # entries of a Wishart mean: std <= sqrt((K_ii K_jj + K_ij^2) / S) <= 1.1 * sqrt(2 / S) ~ 0.011; 6 sigma
    assert np.abs(C - K).max() < 0.07 and abs(Y.mean()) < 0.02


def test_get_samples_contract_and_limits():
    from npf_b200.utils.gp import GPSampler
    s = GPSampler(dict(kind="rbf", length_scale=0.2), n_points=128, n_same_samples=20)
    torch.manual_seed(1)
    X, Y = s.get_samples(50)
    assert X.shape == (50, 128, 1) and Y.shape == (50, 128, 1) and X.is_cuda and Y.is_cuda
    assert float(X.min()) >= -1 and float(X.max()) <= 1 and bool((X[:, 1:] >= X[:, :-1]).all())
    assert torch.isfinite(Y).all() and 0.3 < float(Y.std()) < 2.0
    with pytest.raises(NotImplementedError):
        s.sample_targets(torch.zeros(1, 300), 1)                              # beyond one CTA's shared memory
    # a model trains on the generated tasks through the device split
    import npf_b200
    from npf_b200.utils import datasplit as ds
    Xc, Yc, Xt, Yt = ds.CntxtTrgtGetter(contexts_getter=ds.GetRandomIndcs(a=3, b=30))(X, Y)
    m = npf_b200.CNP(1, 1).cuda().train()
    loss = npf_b200.CNPFLoss()(m(Xc, Yc, Xt, Yt), Yt)
    loss.backward()
    assert torch.isfinite(loss)


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_per_task_hyperparameters(kind):
    """npf_gp_sample_hyp (`is_vary_kernel_hyp`, upstream gaussian_process.py:206-207, 233-242): every task factorises the covariance of
    ITS OWN (length_scale, periodicity, noise_level) row -- checked against the scikit-learn-pinned covariance of the oracle."""
    from npf_b200.utils.gp import GPSampler
    x = FIX[NAMES[0] + "_x"]
    N = len(x)
    s = GPSampler(dict(kind=KIND[kind], length_scale=0.5, periodicity=0.7))
    hyp = torch.tensor([[0.15, 0.4, 0.0], [0.4, 0.9, 0.02], [1.1, 1.7, 0.0], [0.25, 0.55, 0.1]])
    X = torch.from_numpy(np.stack([x] * 4)).float()
    eps = torch.randn(4, 3, N, generator=torch.Generator().manual_seed(2))
    Y, L, rank = s.sample_targets(X, 3, eps=eps, return_factor=True, hyp=hyp)
    L64 = L.double().cpu().numpy()
    for t in range(4):
        K = G.cov_matrix(x, kind, float(hyp[t, 0]), float(hyp[t, 1]), float(hyp[t, 2]))
        assert np.abs(L64[t] @ L64[t].T - K).max() < 5e-5, (kind, t)
    want = np.einsum("tik,tsk->tsi", L64, eps.double().numpy())
    assert np.abs(Y.double().cpu().numpy() - want).max() < 1e-5 * max(1.0, np.abs(want).max())
    assert len(set(int(r) for r in rank)) > 1                                # different kernels, different numerical ranks


def test_vary_kernel_hyp_dataset():
    from npf_b200.utils.gp import GPSampler
    s = GPSampler(dict(kind="rbf", length_scale=0.2, length_scale_bounds=(0.1, 1.0), noise_level_bounds=(0.001, 0.05)),
                  n_points=64, n_same_samples=5, is_vary_kernel_hyp=True)
    torch.manual_seed(3)
    h = s.sample_hyperparameters(1000)
    assert h.shape == (1000, 3) and 0.1 <= float(h[:, 0].min()) and float(h[:, 0].max()) <= 1.0
    assert abs(float(h[:, 0].mean()) - 0.55) < 0.03 and bool((h[:, 1] == 1.0).all())          # uniform in the bounds; periodicity untouched
    assert 0.001 <= float(h[:, 2].min()) and float(h[:, 2].max()) <= 0.05
    X, Y = s.get_samples(43)
    assert X.shape == (43, 64, 1) and Y.shape == (43, 64, 1) and torch.isfinite(Y).all()
    fixed = GPSampler(dict(kind="rbf", length_scale=0.2, length_scale_bounds=(0.1, 1.0)), n_points=64)     # bounds ignored unless asked
    assert fixed.bounds == {} and not fixed.is_vary_kernel_hyp
