"""-m gpu: npf_b200.utils.train.train_models end to end on the device (GraphedStep + FlatAdam under the upstream signature):
a small CNP on synthetic 1-D functions for a few epochs -- the loss must fall, the run directory must hold upstream's files,
and eval.csv must equal a fresh eval_loglike of the reloaded best checkpoint."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_train_models_on_device(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import npf_b200
    from torch.utils.data import TensorDataset
    from npf_b200.utils.train import eval_loglike, train_models
    npf_b200.set_precision("bf16x3")
    g = torch.Generator().manual_seed(0)
    N = 512
    x = torch.rand(N, 48, 1, generator=g) * 2 - 1
    y = torch.sin(3 * x + torch.rand(N, 1, 1, generator=g) * 6.28) + 0.05 * torch.randn(N, 48, 1, generator=g)
    train, test = TensorDataset(x[:448], y[:448]), TensorDataset(x[448:], y[448:])

    def collate(batch):
        X, Y = torch.stack([b[0] for b in batch]), torch.stack([b[1] for b in batch])
        return dict(X_cntxt=X[:, :16].contiguous(), Y_cntxt=Y[:, :16].contiguous(), X_trgt=X.contiguous(), Y_trgt=Y.contiguous()), Y

    kw = dict(criterion=npf_b200.CNPFLoss, chckpnt_dirname=str(tmp_path) + "/", device="cuda", max_epochs=8, batch_size=64, lr=1e-3,
              decay_lr=10, seed=123, test_datasets={"sin": test}, train_split=None, iterator_train__collate_fn=collate,
              iterator_valid__collate_fn=collate)
    try:
        tr = train_models({"sin": train}, {"CNP": npf_b200.CNP(1, 1)}, is_retrain=True, **kw)["sin/CNP/run_0"]
        hist = tr.history
        assert len(hist) == 8 and hist[-1]["train_loss"] < hist[0]["train_loss"] - 1.0, [h["train_loss"] for h in hist]
        run_dir = tmp_path / "sin" / "CNP" / "run_0"
        assert {"params.pt", "optimizer.pt", "history.json", "eval.csv", "model_summary.txt"} <= {p.name for p in run_dir.iterdir()}
        ev = np.loadtxt(run_dir / "eval.csv", delimiter=",")
        assert ev.shape == (64,) and np.isfinite(ev).all()
        tr.device = "cuda"
        ll = eval_loglike(tr, test)
        assert np.allclose(ll, ev, rtol=1e-4, atol=1e-3)
        # upstream's own loader reads the files: a torch.optim.Adam state dict and a plain state_dict
        sd = torch.load(run_dir / "params.pt", map_location="cpu", weights_only=True)
        assert set(sd) == set(npf_b200.CNP(1, 1).state_dict())
        osd = torch.load(run_dir / "optimizer.pt", map_location="cpu", weights_only=True)
        assert osd["param_groups"][0]["params"] == list(range(len(sd))) and int(osd["state"][0]["step"]) == 8 * 7
    finally:
        npf_b200.set_precision("fp32")
