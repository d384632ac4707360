"""skorch-free ``train_models`` / ``eval_loglike`` serving the signatures of upstream utils/train.py:34-305 and
utils/evaluate.py:9-28 (the two functions every notebook drives the models with), on the B200 step API.

    trainers = train_models(datasets, {"ConvCNP": model}, criterion=CNPFLoss, chckpnt_dirname="results/", device=None,
                            max_epochs=100, batch_size=32, lr=1e-3, decay_lr=10, seed=123, is_retrain=True,
                            test_datasets=test_sets, train_split=None,
                            iterator_train__collate_fn=collate, iterator_valid__collate_fn=collate)

What is kept: argument names and meaning, the loop nest datasets x models x runs, the directory layout
``<chckpnt_dirname><data>/<model>/run_<k>/{params.pt, optimizer.pt, history.json, eval.csv, model_summary.txt}`` (best
epoch by validation loss when there is a validation set, last epoch otherwise), exponential learning-rate decay by a
total factor ``decay_lr``, per-run seeding ``seed + run``, early stopping by ``patience``, reloading the checkpoint before
the test evaluation, ``eval.csv`` = per-task test log-likelihood in dataset order (``eval_loglike``: seed 123,
``reduction=None``), the printed summary line, models moved back to the CPU at the end, and the returned
``{suffix: trainer}`` dict whose values expose ``module_``, ``criterion_``, ``history`` (list of per-epoch dicts with
``train_loss``, ``valid_loss``, ``*_loss_best``, ``dur``), ``test_history``, ``device``, ``get_iterator`` and
``validation_step`` like the skorch ``NeuralNet`` the notebooks poke at.

What is different (skorch 0.8 is not a dependency): ``train_split`` is ``None``, a fraction (default 0.1, upstream's
``CVSplit(0.1)``) or a callable ``dataset -> (train, valid)``; ``callbacks`` are plain callables ``cb(trainer, epoch_record)``
run after every epoch; skorch-style ``iterator_{train,valid}__<kw>`` keyword arguments are passed to the DataLoaders; on a
CUDA device with ``optimizer=Adam`` a step is ``GraphedStep`` (forward + loss + backward as one CUDA-graph replay per
input-shape signature) followed by ``FlatAdam`` (one fused kernel); any other optimizer / device runs the plain eager
loop (which is also what the CPU unit test of the host logic exercises with a toy module).
"""
import json
import os
import random
import time
from copy import deepcopy

import numpy as np
import torch
from torch.optim import Adam
from torch.utils.data import DataLoader, Subset

from .checkpoint import load_checkpoint, save_checkpoint

__all__ = ["train_models", "eval_loglike", "Trainer"]

EVAL_FILENAME = "eval.csv"
MOD_SUMM_FILENAME = "model_summary.txt"


def set_seed(seed):
    """upstream utils/helpers.py:49-55"""
    if seed is not None:
        np.random.seed(seed)
        random.seed(seed)
        torch.manual_seed(seed)


def get_exponential_decay_gamma(scheduling_factor, max_epochs):
    """gamma such that lr decays by ``scheduling_factor`` over ``max_epochs`` (upstream utils/helpers.py)"""
    return (1 / scheduling_factor) ** (1 / max_epochs)


def _to_device(obj, device):
    if torch.is_tensor(obj):
        return obj.to(device, non_blocking=True)
    if isinstance(obj, dict):
        return {k: _to_device(v, device) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_device(v, device) for v in obj)
    return obj


class Trainer:
    """The object ``train_models`` returns per (dataset, model, run): the parts of ``skorch.NeuralNet`` the notebooks use."""

    def __init__(self, module, criterion, device, optimizer=Adam, lr=1e-3, batch_size=16, iterator_train=None, iterator_valid=None,
                 max_grad_norm=None):
        self.module_ = module() if isinstance(module, type) or (callable(module) and not isinstance(module, torch.nn.Module)) else module
        self.criterion_ = criterion() if isinstance(criterion, type) or not isinstance(criterion, torch.nn.Module) else criterion
        self.device, self.lr, self.batch_size = device, lr, batch_size
        self.optimizer_cls, self.max_grad_norm = optimizer, max_grad_norm
        self.iterator_train, self.iterator_valid = dict(iterator_train or {}), dict(iterator_valid or {})
        self.history, self.test_history = [], []
        self._opt = self._flat = self._gstep = None

    # ---------------------------------------------------------------- data
    def get_iterator(self, dataset, training=False):
        kw = dict(self.iterator_train if training else self.iterator_valid)
        kw.setdefault("batch_size", self.batch_size if training else 2 * self.batch_size)
        kw.setdefault("shuffle", bool(training))
        return DataLoader(dataset, **kw)

    @staticmethod
    def _unpack(batch):
        """(inputs dict, targets) as upstream's collate returns them; targets default to inputs['Y_trgt']"""
        if isinstance(batch, (list, tuple)) and len(batch) == 2 and isinstance(batch[0], dict):
            X, y = batch
        elif isinstance(batch, dict):
            X, y = batch, batch.get("Y_trgt")
        else:
            raise TypeError("the collate function must return (dict(X_cntxt, Y_cntxt, X_trgt, Y_trgt), Y_trgt)")
        return X, (y if y is not None else X.get("Y_trgt"))

    # ---------------------------------------------------------------- steps
    def initialize(self):
        self.module_.to(self.device)
        self.criterion_.to(self.device) if hasattr(self.criterion_, "to") else None
        if self._opt is not None:
            return self
        on_gpu = torch.device(self.device).type == "cuda"
        if on_gpu and self.optimizer_cls in (Adam, "adam"):
            from ..graph import GraphedStep
            from ..parallel import FlatAdam, FlatGradients
            self._flat = FlatGradients(self.module_)
            self._opt = FlatAdam(self._flat, lr=self.lr)
            self._gstep = GraphedStep(self.module_, self.criterion_, flat=self._flat, max_graphs=64)
        else:
            self._opt = self.optimizer_cls(self.module_.parameters(), lr=self.lr)
        return self

    def train_step(self, Xi, yi):
        self.module_.train()
        self.criterion_.train()
        if self._gstep is not None:
            loss = self._gstep(Xi["X_cntxt"], Xi["Y_cntxt"], Xi["X_trgt"], Xi["Y_trgt"])
            self._opt.lr = self.lr
            self._opt.step(max_grad_norm=self.max_grad_norm)
            return loss
        self._opt.zero_grad()
        for g in self._opt.param_groups:
            g["lr"] = self.lr
        loss = self.criterion_(self.module_(**Xi), yi)
        loss.backward()
        if self.max_grad_norm is not None:
            torch.nn.utils.clip_grad_norm_(self.module_.parameters(), self.max_grad_norm)
        self._opt.step()
        return loss.detach()

    def validation_step(self, Xi, yi, **kwargs):
        self.module_.eval()
        self.criterion_.eval()
        with torch.no_grad():
            Xi, yi = _to_device(Xi, self.device), _to_device(yi, self.device)
            return dict(loss=self.criterion_(self.module_(**Xi), yi))

    def run_epoch(self, dataset, training):
        tot, n = None, 0
        for batch in self.get_iterator(dataset, training=training):
            Xi, yi = self._unpack(batch)
            Xi, yi = _to_device(Xi, self.device), _to_device(yi, self.device)
            bs = int(yi.shape[0])
            loss = self.train_step(Xi, yi) if training else self.validation_step(Xi, yi)["loss"]
            loss = loss.detach().float().mean() * bs           # accumulated on the device: one host sync per epoch
            tot = loss.clone() if tot is None else tot + loss
            n += bs
        return float(tot) / max(n, 1) if tot is not None else float("nan")

    # ---------------------------------------------------------------- checkpoints
    def save_params(self, dirname):
        save_checkpoint(dirname, self.module_, self._opt, history=self.history)

    def load_params(self, checkpoint=None, dirname=None):
        dirname = dirname if dirname is not None else checkpoint
        hist = load_checkpoint(dirname, self.module_, self._opt)
        if hist:
            self.history = hist
        self.module_.to(self.device)
        return self


def eval_loglike(trainer, dataset, seed=123, **kwargs):
    """Log-likelihood of every task of ``dataset`` in order (upstream utils/evaluate.py:9-28): seed fixed so that the
    context / target draws are reproducible, criterion switched to ``reduction=None``, eval mode (always NPML)."""
    set_seed(seed)
    trainer.module_.to(trainer.device)
    old_reduction = trainer.criterion_.reduction
    trainer.criterion_.reduction = None
    all_losses = []
    try:
        for batch in trainer.get_iterator(dataset, training=False):
            Xi, yi = trainer._unpack(batch)
            step = trainer.validation_step(Xi, yi, **kwargs)
            all_losses.append(-step["loss"])             # log likelihood instead of NLL
    finally:
        trainer.criterion_.reduction = old_reduction
    return torch.cat(all_losses, dim=0).detach().cpu().numpy()


def _eval_save_load(trainer, data_test, test_eval_file, is_force_rerun=False):
    """upstream utils/train.py:315-330"""
    test_loglike = None
    if data_test is not None:
        if test_eval_file is not None and os.path.exists(test_eval_file):
            test_loglike = np.loadtxt(test_eval_file, delimiter=",")
        if is_force_rerun or test_loglike is None:
            test_loglike = eval_loglike(trainer, data_test)
        if test_eval_file is not None:
            os.makedirs(os.path.dirname(test_eval_file), exist_ok=True)
            np.savetxt(test_eval_file, test_loglike, delimiter=",")
        return test_loglike.mean(axis=0)


def _best_loss(trainer, mode="valid"):
    for epoch, rec in enumerate(trainer.history[::-1]):
        if rec.get(f"{mode}_loss_best"):
            return rec[f"{mode}_loss"], len(trainer.history) - epoch
    return None, None


def _split(dataset, train_split, seed):
    if train_split is None:
        return dataset, None
    if callable(train_split):
        return train_split(dataset)
    n = len(dataset)
    n_valid = max(1, int(round(float(train_split) * n)))
    perm = np.random.RandomState(0 if seed is None else seed).permutation(n)
    return Subset(dataset, perm[n_valid:].tolist()), Subset(dataset, perm[:n_valid].tolist())


def round_decimals(x, n=4):
    return None if x is None else float(("{:." + str(n) + "f}").format(x))


def train_models(datasets, models, criterion, test_datasets=dict(), valid_datasets=dict(), chckpnt_dirname=None,
                 is_continue_train=False, is_retrain=False, runs=1, starting_run=0, train_split=0.1, device=None, max_epochs=100,
                 batch_size=16, lr=1e-3, optimizer=Adam, callbacks=(), patience=None, decay_lr=None, is_reeval=False, seed=None,
                 datasets_kwargs=dict(), models_kwargs=dict(), **kwargs):
    """Train (``is_retrain=True``) or load every model on every dataset; see the module docstring and upstream
    utils/train.py:34-156 for the arguments."""
    trainers = dict()
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    kwargs = dict(kwargs)
    if "iterator_train__shuffle" not in kwargs:
        kwargs["iterator_train__shuffle"] = True
    if "iterator_valid__batch_size" not in kwargs:
        kwargs["iterator_valid__batch_size"] = batch_size * 2

    for data_name, data_train in datasets.items():
        current_models = models[data_name] if isinstance(list(models.values())[0], dict) else models
        data_test = test_datasets.get(data_name, None)
        data_valid = valid_datasets.get(data_name, None)
        for model_name, model in current_models.items():
            for run in range(starting_run, starting_run + runs):
                suffix = data_name + "/" + model_name + "/run_{}".format(run)
                print("\n--- {} {} ---\n".format("Training" if is_retrain else "Loading", suffix), flush=True)
                run_dir = None if chckpnt_dirname is None else chckpnt_dirname + suffix
                test_eval_file = None if run_dir is None else os.path.join(run_dir, EVAL_FILENAME)
                cur = dict(kwargs)
                cur.update(datasets_kwargs.get(data_name, dict()))
                cur.update(models_kwargs.get(model_name, dict()))
                it_train = {k[len("iterator_train__"):]: v for k, v in cur.items() if k.startswith("iterator_train__")}
                it_valid = {k[len("iterator_valid__"):]: v for k, v in cur.items() if k.startswith("iterator_valid__")}
                if seed is not None:
                    set_seed(seed + run)                                   # FixRandomSeed(seed + run)
                trainer = Trainer(deepcopy(model) if isinstance(model, torch.nn.Module) and runs > 1 else model, criterion, device,
                                  optimizer=optimizer, lr=cur.get("lr", lr), batch_size=cur.get("batch_size", batch_size),
                                  iterator_train=it_train, iterator_valid=it_valid, max_grad_norm=cur.get("max_grad_norm"))
                trainer.initialize()
                if data_valid is not None:
                    d_train, d_valid = data_train, data_valid
                else:
                    d_train, d_valid = _split(data_train, train_split, None if seed is None else seed + run)

                if is_continue_train:
                    assert run_dir is not None, "is_continue_train needs chckpnt_dirname"
                    if os.path.exists(os.path.join(run_dir, "params.pt")):
                        trainer.load_params(dirname=run_dir)

                if is_retrain:
                    gamma = None if decay_lr is None else get_exponential_decay_gamma(decay_lr, max_epochs)
                    best, since_best = {"train": float("inf"), "valid": float("inf")}, 0
                    first_epoch = len(trainer.history)
                    for epoch in range(first_epoch, first_epoch + max_epochs):
                        t0 = time.time()
                        train_loss = trainer.run_epoch(d_train, training=True)
                        valid_loss = trainer.run_epoch(d_valid, training=False) if d_valid is not None else None
                        rec = dict(epoch=epoch + 1, train_loss=train_loss, valid_loss=valid_loss, dur=time.time() - t0, lr=trainer.lr,
                                   train_loss_best=train_loss < best["train"],
                                   valid_loss_best=valid_loss is not None and valid_loss < best["valid"])
                        best["train"] = min(best["train"], train_loss)
                        if valid_loss is not None:
                            best["valid"] = min(best["valid"], valid_loss)
                        trainer.history.append(rec)
                        print("  epoch {:4d}  train_loss {:12.4f}  valid_loss {}  lr {:.2e}  {:.1f}s".format(
                            rec["epoch"], train_loss, "{:12.4f}".format(valid_loss) if valid_loss is not None else "     -      ", trainer.lr,
                            rec["dur"]), flush=True)
                        if run_dir is not None and (d_valid is None or rec["valid_loss_best"]):
                            trainer.save_params(run_dir)                     # Checkpoint(monitor=None | "valid_loss_best")
                        for cb in callbacks:
                            cb(trainer, rec)
                        if gamma is not None:
                            trainer.lr *= gamma                              # LRScheduler(ExponentialLR, gamma)
                        since_best = 0 if (valid_loss is None or rec["valid_loss_best"]) else since_best + 1
                        if patience is not None and since_best >= patience:
                            print("  early stopping: no validation improvement in {} epochs".format(patience), flush=True)
                            break
                    if run_dir is not None:
                        os.makedirs(run_dir, exist_ok=True)
                        with open(os.path.join(run_dir, "history.json"), "w") as f:
                            json.dump(trainer.history, f)
                        with open(os.path.join(run_dir, MOD_SUMM_FILENAME), "w") as f:
                            f.write(str(trainer.module_))

                # load in all cases => even after training the BEST checkpoint is what gets evaluated (upstream 266-268)
                if run_dir is not None and os.path.exists(os.path.join(run_dir, "params.pt")):
                    hist = deepcopy(trainer.history)
                    trainer.load_params(dirname=run_dir)
                    if is_retrain:
                        trainer.history = hist
                elif not is_retrain:
                    raise FileNotFoundError(f"train_models(is_retrain=False): no checkpoint under {run_dir!r}")

                test_loglike = _eval_save_load(trainer, data_test, test_eval_file, is_force_rerun=is_retrain or is_reeval)
                trainer.test_history = deepcopy(trainer.history)
                valid_loss, best_epoch = _best_loss(trainer, mode="valid")
                train_loss, _ = _best_loss(trainer, mode="train")
                print(suffix, "| best epoch:", best_epoch, "| train loss:", round_decimals(train_loss), "| valid loss:",
                      round_decimals(valid_loss), "| test log likelihood:", round_decimals(None if test_loglike is None else float(test_loglike)),
                      flush=True)
                if trainer._flat is not None:
                    trainer._flat.detach()
                    trainer._gstep = None
                trainer.module_.cpu()                                       # upstream 300
                if torch.cuda.is_available():
                    torch.cuda.empty_cache()
                trainers[suffix] = trainer
    return trainers
