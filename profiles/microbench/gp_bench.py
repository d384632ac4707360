"""One epoch of synthetic GP tasks (50 000 tasks x 128 points, n_same_samples = 20: utils/ntbks_helpers.py:90-98) generated on
the device (npf_b200.utils.gp.GPSampler) against scikit-learn's sample_y on the host cores for a bounded sample of the same
workload (the reference's generator).  One JSON line per kernel.
    python profiles/microbench/gp_bench.py"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "neural-process-family_b200"))
from npf_b200.utils.gp import GPSampler  # noqa: E402


def main():
    try:
        from sklearn.gaussian_process import GaussianProcessRegressor
        from sklearn.gaussian_process.kernels import RBF, ExpSineSquared, Matern, WhiteKernel
    except ImportError:
        print(json.dumps(dict(error="scikit-learn not importable")))
        return
    n_samples, n_points, same = 50000, 128, 20
    for name, kernel in [("RBF_Kernel", RBF(length_scale=0.2)), ("Periodic_Kernel", ExpSineSquared(length_scale=0.5, periodicity=0.5)),
                         ("Noisy_Matern_Kernel", WhiteKernel(noise_level=0.1) + Matern(length_scale=0.2, nu=1.5))]:
        s = GPSampler(kernel, n_points=n_points, n_same_samples=same)
        s.get_samples(2000)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        X, Y = s.get_samples(n_samples)
        e1.record()
        torch.cuda.synchronize()
        dev_ms = e0.elapsed_time(e1)
        gpr = GaussianProcessRegressor(kernel=kernel, alpha=0.005)
        n_host = 2000                                        # bounded sample: 100 position sets x 20 draws
        t0 = time.perf_counter()
        for _ in range(n_host // same):
            x = np.sort(np.random.uniform(-2, 2, size=(n_points, 1)), axis=0)
            gpr.sample_y(x, n_samples=same, random_state=None)
        host_ms = (time.perf_counter() - t0) * 1e3 * (n_samples / n_host)
        print(json.dumps(dict(workload=f"{name}_50000x128_same20", device_ms_per_epoch=round(dev_ms, 2),
                              host_sklearn_ms_per_epoch_extrapolated=round(host_ms, 1), host_sample=f"{n_host} tasks",
                              ratio=round(host_ms / dev_ms, 1), y_std=round(float(Y.std()), 3))), flush=True)


if __name__ == "__main__":
    main()
