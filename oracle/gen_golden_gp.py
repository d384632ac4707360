"""Generate tests/golden/gp/kernels.npz with scikit-learn (build container only):  python oracle/gen_golden_gp.py
Covariance matrices k(X, X) of the reference's GP datasets (utils/ntbks_helpers.py:76-108) on sorted uniform positions in
[-2, 2] (GPDataset._sample_features, utils/data/gaussian_process.py:189-194), exactly as sample_y would build them."""
import os

import numpy as np
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ExpSineSquared, Matern, WhiteKernel

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "gp", "kernels.npz")


def main():
    rng = np.random.RandomState(0)
    cases = {
        "rbf": (RBF(length_scale=0.2), 128, dict(kind=0, length_scale=0.2, periodicity=1.0, noise_level=0.0)),
        "noisy_matern": (WhiteKernel(noise_level=0.1) + Matern(length_scale=0.2, nu=1.5), 128,
                         dict(kind=1, length_scale=0.2, periodicity=1.0, noise_level=0.1)),
        "periodic": (ExpSineSquared(length_scale=0.5, periodicity=0.5), 128, dict(kind=2, length_scale=0.5, periodicity=0.5, noise_level=0.0)),
        "matern_short": (Matern(length_scale=0.05, nu=1.5), 50, dict(kind=1, length_scale=0.05, periodicity=1.0, noise_level=0.0)),
        "rbf_n200": (RBF(length_scale=0.3), 200, dict(kind=0, length_scale=0.3, periodicity=1.0, noise_level=0.0)),
    }
    out = {}
    for name, (kernel, n, hyp) in cases.items():
        x = np.sort(rng.uniform(-2, 2, size=n))
        gpr = GaussianProcessRegressor(kernel=kernel, alpha=0.005)        # as GPDataset builds it
        _, K = gpr.predict(x[:, None], return_cov=True)                  # un-fitted: the prior covariance sample_y draws from
        assert np.allclose(K, kernel(x[:, None]))
        out[name + "_x"] = x
        out[name + "_K"] = K
        out[name + "_hyp"] = np.array([hyp["kind"], hyp["length_scale"], hyp["periodicity"], hyp["noise_level"]])
        w = np.linalg.eigvalsh(K)
        print(name, n, "eig min/max", w.min(), w.max(), "rank@1e-6", int((w > 1e-6).sum()))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
