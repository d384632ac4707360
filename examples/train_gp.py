"""Train a conditional NP on synthetic GP tasks with EVERYTHING on the device -- the end-to-end sanity run BASELINE.md
asks for (upstream: jupyter/reproducibility/{CNP,AttnCNP,ConvCNP}.ipynb through utils/train.py; published test
log-likelihoods in BASELINE.md section 1).

Per epoch: 50 000 fresh tasks from ``GPSampler`` (upstream: GPDataset, is_reuse_across_epochs=False); per batch: context
size ~ U{0..50} on the host RNG, subsets on the device (``CntxtTrgtGetter(GetRandomIndcs(a=0.0, b=50), get_all_indcs)``,
utils/ntbks_helpers.py:272-286), forward + loss + backward as a CUDA-graph replay (``GraphedStep``, one graph per context
size), Adam on the flat bucket (``FlatAdam``), ExponentialLR to lr / decay over the run (utils/train.py:237-254).
Evaluation like upstream: log p(targets | context) summed over the 128 points of a task, mean over the test tasks.

    python examples/train_gp.py --model convcnp --kernel rbf --epochs 30
Prints one JSON line."""
import argparse
import json
import os
import random
import sys
import time
from functools import partial

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neural-process-family_b200"))
import npf_b200  # noqa: E402
from npf_b200 import AttnCNP, CNP, CNPFLoss, ConvCNP, GraphedStep  # noqa: E402
from npf_b200.architectures import MLP, merge_flat_input  # noqa: E402
from npf_b200.parallel import FlatAdam, FlatGradients  # noqa: E402
from npf_b200.utils import datasplit as ds  # noqa: E402
from npf_b200.utils.gp import GPSampler  # noqa: E402

KERNELS = {
    "rbf": dict(kind="rbf", length_scale=0.2),
    "periodic": dict(kind="periodic", length_scale=0.5, periodicity=0.5),
    "noisy_matern": dict(kind="matern15", length_scale=0.2, noise_level=0.1),
}
R = 128


def make_model(name):
    """The notebooks' 1-D configurations (CNP.ipynb / AttnCNP.ipynb cell 7); ConvCNP with its constructor defaults
    (the configuration bench.py measures)."""
    kw = dict(XEncoder=partial(MLP, n_hidden_layers=1, hidden_size=R),
              Decoder=merge_flat_input(partial(MLP, n_hidden_layers=4, hidden_size=R), is_sum_merge=True), r_dim=R)
    if name == "cnp":
        return CNP(1, 1, XYEncoder=merge_flat_input(partial(MLP, n_hidden_layers=2, hidden_size=R * 2), is_sum_merge=True), **kw)
    if name == "attncnp":
        return AttnCNP(1, 1, attention="transformer",
                       XYEncoder=merge_flat_input(partial(MLP, n_hidden_layers=2, hidden_size=R), is_sum_merge=True), **kw)
    if name == "convcnp":
        return ConvCNP(1, 1)
    raise ValueError(name)


def evaluate(model, sampler, getter, n_tasks, batch):
    model.eval()
    crit = CNPFLoss(reduction=None).eval()
    X, Y = sampler.get_samples(n_tasks)
    ll = []
    with torch.no_grad():
        for i in range(0, n_tasks, batch):
            Xc, Yc, Xt, Yt = getter(X[i:i + batch], Y[i:i + batch])
            ll.append(-crit(model(Xc, Yc, Xt, Yt), Yt))
    model.train()
    return torch.cat(ll).mean().item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="convcnp", choices=["cnp", "attncnp", "convcnp"])
    ap.add_argument("--kernel", default="rbf", choices=list(KERNELS))
    ap.add_argument("--epochs", type=int, default=30)
    ap.add_argument("--tasks-per-epoch", type=int, default=50000)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--decay", type=float, default=10.0)
    ap.add_argument("--eval-tasks", type=int, default=10000)
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--seed", type=int, default=123)
    a = ap.parse_args()

    random.seed(a.seed); np.random.seed(a.seed); torch.manual_seed(a.seed)
    npf_b200.set_precision(a.precision)
    model = make_model(a.model).cuda().train()
    crit = CNPFLoss(reduction="mean").train()
    flat = FlatGradients(model)
    opt = FlatAdam(flat, lr=a.lr)
    step = GraphedStep(model, crit, flat=flat, max_graphs=64)
    sampler = GPSampler(KERNELS[a.kernel], min_max=(-2, 2), n_points=128, n_same_samples=20)
    getter = ds.CntxtTrgtGetter(contexts_getter=ds.GetRandomIndcs(a=0.0, b=50), targets_getter=ds.get_all_indcs)
    gamma = (1.0 / a.decay) ** (1.0 / max(a.epochs, 1))
    n_batches = a.tasks_per_epoch // a.batch

    ll0 = evaluate(model, sampler, getter, a.eval_tasks, a.batch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t_gen = t_train = 0.0
    losses = []
    for epoch in range(a.epochs):
        te = time.perf_counter()
        X, Y = sampler.get_samples(n_batches * a.batch)
        torch.cuda.synchronize()
        t_gen += time.perf_counter() - te
        te = time.perf_counter()
        acc = torch.zeros((), device="cuda")
        for i in range(n_batches):
            Xc, Yc, Xt, Yt = getter(X[i * a.batch:(i + 1) * a.batch], Y[i * a.batch:(i + 1) * a.batch])
            acc += step(Xc, Yc, Xt, Yt)
            opt.step()
        opt.lr *= gamma
        torch.cuda.synchronize()
        t_train += time.perf_counter() - te
        losses.append(round(acc.item() / n_batches, 3))
        if hasattr(model, "validate_now"):
            model.validate_now()
    wall = time.perf_counter() - t0
    ll1 = evaluate(model, sampler, getter, a.eval_tasks, a.batch)
    n_tasks = a.epochs * n_batches * a.batch
    print(json.dumps(dict(model=a.model, kernel=a.kernel, precision=a.precision, epochs=a.epochs, tasks_trained=n_tasks, batch=a.batch,
                          lr=a.lr, test_loglik_per_task_before=round(ll0, 2), test_loglik_per_task_after=round(ll1, 2),
                          train_nll_per_epoch=losses, wall_s=round(wall, 2), data_gen_s=round(t_gen, 3), train_s=round(t_train, 2),
                          tasks_per_s_whole_loop=round(n_tasks / wall, 1), n_params=sum(p.numel() for p in model.parameters()),
                          graphs=len(step._graphs))))


if __name__ == "__main__":
    main()
