"""CPU-only end-to-end check of the DATA side's law (test infrastructure; run by hand, ~1 min):
tasks drawn with the restated device algorithms (oracle/gp_oracle.py pivoted-Cholesky sampler, oracle/datasplit_oracle.py
Philox subsets, context size U{0..50}) are scored by the reference's published RBF ConvCNP checkpoint (the weights stored
in tests/golden/convcnp_notebook_pretrained.pt) through the CPU oracle of the model.  If the generator and the split
follow upstream's law, the per-task test log-likelihoods must be distributed like upstream's own eval.csv (published mean
175.12, BASELINE.md section 1 -- a mean over 156 batches that share one drawn context size each, s.e. 22.5; the upper
quantiles, dense contexts, are the sharp part of the comparison).
    python oracle/check_data_law.py [n_tasks]"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from oracle import datasplit_oracle as D, gp_oracle as G, npf_oracle as O  # noqa: E402


STRATIFIED = True   # cycle the context size through 0..50 instead of drawing it: same expectation, far smaller variance


def main(n_tasks=1920, batch=32, same=20, seed=123):
    torch.set_num_threads(8)
    random.seed(seed)
    rng = np.random.RandomState(seed)
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "convcnp_notebook_pretrained.pt"), map_location="cpu", weights_only=False)
    sd = {k: v for k, v in fx["state_dict"].items()}
    Xi = O.induced_grid(fx["cfg"]["density_induced"])
    xs, ys = [], []
    while len(xs) < n_tasks:
        x = np.sort(rng.uniform(-2, 2, size=128))
        y, _, _ = G.sample(x, rng.randn(same, 128), G.RBF, 0.2, tol=1e-5)
        xs += [x] * same
        ys += list(y)
    perm = rng.permutation(len(xs))[:n_tasks]
    X = torch.tensor(np.stack(xs)[perm] / 2.0, dtype=torch.float32).unsqueeze(-1)      # rescale [-2, 2] -> [-1, 1]
    Y = torch.tensor(np.stack(ys)[perm], dtype=torch.float32).unsqueeze(-1)
    ll = []
    with torch.no_grad():
        for i in range(0, n_tasks, batch):
            xb, yb = X[i:i + batch], Y[i:i + batch]
            n = (i // batch) % 51 if STRATIFIED else random.randint(0, 50)
            idx = torch.from_numpy(D.random_subset(len(xb), 128, n, seed=rng.randint(2 ** 31 - 1))).long()
            bi = torch.arange(len(xb))[:, None]
            loc, scale = O.convcnp_forward(sd, xb[bi, idx], yb[bi, idx], xb, X_induced=Xi, training=False)
            ll.append(-O.cnpf_loss(loc, scale, yb, reduction=None))
    bm = torch.stack([b.mean() for b in ll])          # the context size is shared by a batch: batches are the independent units
    ll = torch.cat(ll)
    q = np.percentile(ll.numpy(), [5, 25, 50, 75, 90, 99]).round(1)
    print("per-task log-lik quantiles 5/25/50/75/90/99 %:", q.tolist(),
          "| upstream's eval.csv (RBF ConvCNP, 10 000 tasks): [-192.9, 22.6, 271.9, 377.2, 419.7, 449.2]")
    if STRATIFIED:
        cyc = bm[: len(bm) // 51 * 51].view(-1, 51).mean(1)          # one value per full cycle of context sizes
        print(f"tasks {n_tasks}: mean test log-lik per task {cyc.mean().item():.2f} +- {cyc.std().item() / len(cyc) ** 0.5:.2f} "
              f"(s.e. over {len(cyc)} cycles of the 51 context sizes); published 175.12 +- 22.5 (context sizes DRAWN per batch of 64: "
              f"s.e. from the batch means of upstream's eval.csv)")
    else:
        print(f"tasks {n_tasks}: mean test log-lik per task {ll.mean().item():.2f} +- {bm.std().item() / len(bm) ** 0.5:.2f} "
              f"(s.e. over {len(bm)} batches); published 175.12")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1920)
