"""CPU: pins oracle/gp_oracle.py against covariance matrices built by scikit-learn itself (tests/golden/gp/kernels.npz,
oracle/gen_golden_gp.py) and checks the pivoted factorisation it restates: L L^T == k(X, X) up to the stopping tolerance,
with the rank of these matrices far below N (why a plain Cholesky is not an option)."""
import os

import numpy as np
import pytest

from _util import ROOT
from oracle import gp_oracle as G

FIX = np.load(os.path.join(ROOT, "tests", "golden", "gp", "kernels.npz"))
NAMES = sorted(k[:-2] for k in FIX.files if k.endswith("_x"))


@pytest.mark.parametrize("name", NAMES)
def test_kernels_match_sklearn(name):
    kind, ls, per, noise = FIX[name + "_hyp"]
    K = G.cov_matrix(FIX[name + "_x"], int(kind), ls, per, noise)
    assert np.abs(K - FIX[name + "_K"]).max() < 1e-12


@pytest.mark.parametrize("name", NAMES)
def test_pivoted_cholesky_reproduces_the_covariance(name):
    K = FIX[name + "_K"]
    L, rank = G.pivoted_cholesky(K, 1e-6)
    assert np.abs(L @ L.T - K).max() < 5e-6
    assert np.count_nonzero(np.abs(L[:, rank:]).sum(0)) == 0
    if name in ("rbf", "periodic", "rbf_n200"):
        assert rank < K.shape[0] // 2            # numerically rank-deficient: jitter-free plain Cholesky would break down
        with pytest.raises(np.linalg.LinAlgError):
            np.linalg.cholesky(K)
    else:
        assert rank == K.shape[0]
    eps = np.random.RandomState(1).randn(3, K.shape[0])
    kind, ls, per, noise = FIX[name + "_hyp"]
    y, L2, r2 = G.sample(FIX[name + "_x"], eps, int(kind), ls, per, noise)
    assert r2 == rank and np.allclose(y, eps @ L.T)


def test_host_side_kernel_parsing():
    from npf_b200.utils.gp import GPSampler, kernel_hyperparameters

    class RBF:
        length_scale = 0.2

    class Matern:
        length_scale, nu = 0.3, 1.5

    class WhiteKernel:
        noise_level = 0.1

    class Sum:
        k1, k2 = WhiteKernel(), Matern()

    class ExpSineSquared:
        length_scale, periodicity = 0.5, 0.25

    assert kernel_hyperparameters(RBF()) == dict(kind="rbf", length_scale=0.2, periodicity=1.0, noise_level=0.0)
    assert kernel_hyperparameters(Sum()) == dict(kind="matern15", length_scale=0.3, periodicity=1.0, noise_level=0.1)
    assert kernel_hyperparameters(ExpSineSquared())["periodicity"] == 0.25
    Matern.nu = 2.5
    with pytest.raises(NotImplementedError):
        kernel_hyperparameters(Matern())
    with pytest.raises(RuntimeError):
        GPSampler(RBF(), device="cpu")            # no CPU path
    try:
        from sklearn.gaussian_process.kernels import RBF as SkRBF, WhiteKernel as SkWhite
    except ImportError:
        return
    assert kernel_hyperparameters(SkWhite(noise_level=0.1) + SkRBF(length_scale=0.2))["noise_level"] == 0.1


def test_hyperparameter_bounds_parsing():
    """`is_vary_kernel_hyp` (upstream gaussian_process.py:39-41, 233-242): the bounds that are resampled are the kernel's own `*_bounds`;
    fixed ones stay at the kernel's value (upstream would fail on a "fixed" bound: it draws `uniform(*hyperparam.bounds)` for every entry)."""
    from npf_b200.utils.gp import kernel_hyperparameter_bounds
    assert kernel_hyperparameter_bounds(dict(kind="rbf", length_scale=0.2)) == {}
    assert kernel_hyperparameter_bounds(dict(kind="periodic", length_scale=1.0, periodicity=0.5, length_scale_bounds=(0.5, 2),
                                             periodicity_bounds="fixed", noise_level_bounds=(0.01, 0.1))) == \
        {"length_scale": (0.5, 2.0), "noise_level": (0.01, 0.1)}
    with pytest.raises(ValueError):
        kernel_hyperparameter_bounds(dict(kind="rbf", length_scale=0.2, length_scale_bounds=(0.0, 1.0)))
    try:
        from sklearn.gaussian_process.kernels import RBF, ExpSineSquared, WhiteKernel
    except ImportError:
        return
    k = WhiteKernel(0.1, noise_level_bounds=(0.01, 0.5)) + ExpSineSquared(1.0, 0.5, length_scale_bounds=(0.5, 2), periodicity_bounds="fixed")
    got = kernel_hyperparameter_bounds(k)
    assert got == {"noise_level": (0.01, 0.5), "length_scale": (0.5, 2.0)}
    assert sorted(got) == sorted(h.name.split("__")[-1] for h in k.hyperparameters if not h.fixed)   # the free ones, as scikit-learn sees them
    assert kernel_hyperparameter_bounds(RBF(0.2)) == {"length_scale": (1e-5, 1e5)}           # scikit-learn's default bounds
