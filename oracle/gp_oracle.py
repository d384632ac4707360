"""CPU oracle for the device Gaussian-process sampler  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

fp64 numpy restatement of what ``neural-process-family_b200/csrc/gp_sample.cu`` computes in place of the reference's
host-side generator (utils/data/gaussian_process.py:201-231 -> sklearn ``GaussianProcessRegressor.sample_y`` of the
un-fitted regressor: y ~ N(0, k(X, X))): the stationary kernels of utils/ntbks_helpers.py:76-108 as scikit-learn
1.9 defines them (sklearn/gaussian_process/kernels.py: ``RBF.__call__``, ``Matern.__call__`` nu=1.5,
``ExpSineSquared.__call__``, ``WhiteKernel``), and the diagonally pivoted Cholesky with early termination.

Pinned (tests/test_gp_oracle.py) against covariance matrices produced by scikit-learn itself in the build container
(oracle/gen_golden_gp.py -> tests/golden/gp/kernels.npz).  A sampler has no element-wise golden output (upstream draws
through numpy's SVD-based multivariate_normal on the Mersenne Twister); what is pinned is the LAW: L L^T == k(X, X).
"""
import numpy as np

RBF, MATERN15, PERIODIC = 0, 1, 2


def cov_matrix(x, kind, length_scale, periodicity=1.0, noise_level=0.0):
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    d = np.abs(x[:, None] - x[None, :])
    if kind == RBF:
        K = np.exp(-0.5 * (d / length_scale) ** 2)
    elif kind == MATERN15:
        r = np.sqrt(3.0) * d / length_scale
        K = (1.0 + r) * np.exp(-r)
    elif kind == PERIODIC:
        K = np.exp(-2.0 * (np.sin(np.pi * d / periodicity) / length_scale) ** 2)
    else:
        raise ValueError(kind)
    return K + noise_level * np.eye(len(x))


def pivoted_cholesky(K, tol):
    """K [N, N] PSD -> (L [N, N] with zero columns beyond the rank, rank): K = L L^T + E, max diag(E) <= tol.
    Pivot = largest residual variance, smallest index among ties; rows stay in the original order."""
    K = np.asarray(K, dtype=np.float64)
    N = K.shape[0]
    L = np.zeros((N, N))
    d = np.diag(K).copy()
    done = np.zeros(N, dtype=bool)
    rank = 0
    for k in range(N):
        cand = np.where(done, -1.0, d)
        p = int(np.argmax(cand))
        if not cand[p] > tol:
            break
        rank = k + 1
        col = (K[:, p] - L[:, :k] @ L[p, :k]) / np.sqrt(cand[p])
        col[done] = 0.0
        col[p] = np.sqrt(cand[p])
        L[:, k] = col
        d -= col ** 2
        done[p] = True
    return L, rank


def sample(x, eps, kind, length_scale, periodicity=1.0, noise_level=0.0, tol=1e-6):
    """eps [S, N] -> y [S, N] = (L eps_s)_s."""
    L, rank = pivoted_cholesky(cov_matrix(x, kind, length_scale, periodicity, noise_level), tol)
    return np.asarray(eps, dtype=np.float64) @ L.T, L, rank
