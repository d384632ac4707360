"""Score upstream's PRETRAINED 1-D checkpoints with the B200 kernels on device-generated GP test tasks and print the
test log-likelihood per task next to the published number.  The weights travel with the repo as
the ``state_dict`` of the golden fixtures (tests/golden/*_pretrained.pt, written from results/pretrained/RBF_Kernel/*).

Run on a B200 in round 2 (profiles/r2/eval_pretrained_r2c1.jsonl).
    python examples/eval_pretrained.py [n_tasks]
The context size is cycled through 0..50 (the expectation of upstream's per-batch draw without its variance: upstream's
own mean is over 156 batches that share one drawn size each, s.e. ~22 for ConvCNP); quantiles are printed as well."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "neural-process-family_b200")]
import npf_b200  # noqa: E402
from npf_b200.utils.configs import build_model, loss_for  # noqa: E402
from npf_b200.utils import datasplit as ds  # noqa: E402
from npf_b200.utils.gp import GPSampler  # noqa: E402

PUBLISHED = {  # fixture -> (loss, published RBF test log-lik per task)
    "cnp_notebook_pretrained": ("cnpf", -16.11),
    "attncnp_transformer_pretrained": ("cnpf", 149.16),
    "convcnp_notebook_pretrained": ("cnpf", 175.12),
    "convlnp_notebook_pretrained": ("nll", 224.63),
}


def main(n_tasks=10200, batch=50):
    torch.manual_seed(123); np.random.seed(123)
    npf_b200.set_precision("fp32")
    sampler = GPSampler(dict(kind="rbf", length_scale=0.2), min_max=(-2, 2), n_points=128, n_same_samples=20)
    X, Y = sampler.get_samples(n_tasks)
    for name, (loss_name, published) in PUBLISHED.items():
        fx = torch.load(os.path.join(ROOT, "tests", "golden", name + ".pt"), map_location="cpu", weights_only=False)   # cfg + upstream weights (data only)
        model = build_model(fx["cfg"])
        model.load_state_dict(fx["state_dict"])
        model.cuda().eval()
        if hasattr(model, "n_z_samples_test"):
            model.n_z_samples_test = 32                      # the notebooks' evaluation setting
        crit = loss_for(loss_name).eval()
        ll = []
        with torch.no_grad():
            for i in range(0, n_tasks, batch):
                n = (i // batch) % 51
                getter = ds.CntxtTrgtGetter(contexts_getter=ds.GetRandomIndcs(a=n, b=n))     # exactly n context points
                Xc, Yc, Xt, Yt = getter(X[i:i + batch], Y[i:i + batch])
                ll.append(-crit(model(Xc, Yc, Xt, Yt), Yt))
        ll = torch.cat(ll).double().cpu().numpy()
        print(json.dumps(dict(checkpoint=name, tasks=int(len(ll)), test_loglik_per_task=round(float(ll.mean()), 2), published=published,
                              quantiles_5_25_50_75_90_99=np.percentile(ll, [5, 25, 50, 75, 90, 99]).round(1).tolist())), flush=True)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 10200)
