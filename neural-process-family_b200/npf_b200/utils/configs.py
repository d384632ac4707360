"""Named model configurations -> npf_b200 models: the portable ``cfg`` dicts stored with the golden fixtures / used by the examples
(``family``, ``notebook``, ``cnn``, ``attention`` ... -- the constructor arguments of the upstream notebooks, ``jupyter/reproducibility/*.ipynb``
cell 7) turned into constructor calls.  (The fixture generator does the same with the reference's own classes.)"""
from functools import partial

import torch.nn as nn

from .. import CNP, LNP, AttnCNP, AttnLNP, ConvCNP, ConvLNP, GridConvCNP, GridConvLNP, CNPFLoss, ELBOLossLNPF, NLLLossLNPF
from ..architectures import CNN, MLP, ResConvBlock, SetConv, discard_ith_arg, merge_flat_input
from .helpers import CircularPad2d, make_abs_conv, make_padded_conv

__all__ = ["build_model", "loss_for"]

R_DIM = 128


def build_model(cfg):
    fam = cfg["family"]
    kw = {}
    if cfg.get("notebook"):
        if fam in ("CNP", "AttnCNP", "LNP"):
            kw["XEncoder"] = partial(MLP, n_hidden_layers=1, hidden_size=R_DIM)
            kw["Decoder"] = merge_flat_input(partial(MLP, n_hidden_layers=4, hidden_size=R_DIM), is_sum_merge=True)
        if fam in ("CNP", "AttnCNP", "LNP", "AttnLNP"):
            kw["r_dim"] = R_DIM
            if not cfg.get("is_self_attn"):
                kw["XYEncoder"] = merge_flat_input(
                    partial(MLP, n_hidden_layers=2, hidden_size=cfg["xy_hidden"]), is_sum_merge=True)
        elif fam in ("ConvCNP", "GridConvCNP"):
            kw["r_dim"] = R_DIM
            kw["Decoder"] = discard_ith_arg(partial(MLP, n_hidden_layers=4, hidden_size=R_DIM), i=0)
        elif fam in ("ConvLNP", "GridConvLNP"):
            kw["r_dim"] = R_DIM
            kw["Decoder"] = discard_ith_arg(nn.Linear, i=0)
            kw["is_q_zCct"] = False
    if "cnn" in cfg:
        c = cfg["cnn"]
        Conv = nn.Conv1d if c["dim"] == 1 else nn.Conv2d
        Norm = {None: nn.Identity, "bn": nn.BatchNorm1d if c["dim"] == 1 else nn.BatchNorm2d}[c.get("norm")]
        if "bn_eps" in c:
            Norm = partial(Norm, eps=c["bn_eps"])
        if cfg.get("circular"):  # `model_2d_extrap` of ConvCNP.ipynb / ConvLNP.ipynb: wrap-around padding everywhere
            Conv = make_padded_conv(Conv, CircularPad2d)
            kw["Conv"] = lambda y_dim: make_padded_conv(make_abs_conv(nn.Conv2d), CircularPad2d)(
                y_dim, y_dim, groups=y_dim, kernel_size=11, padding=11 // 2, bias=False)
        kw["CNN"] = partial(CNN, ConvBlock=ResConvBlock, Conv=Conv, Normalization=Norm, n_blocks=c["n_blocks"],
                            kernel_size=c["kernel_size"], is_chan_last=True, n_conv_layers=c["n_conv_layers"])
    for k in ("density_induced", "attention", "n_z_samples_train", "n_z_samples_test", "is_global", "encoded_path",
              "is_q_zCct", "is_self_attn"):
        if k in cfg:
            kw[k] = cfg[k]
    if fam in ("ConvCNP", "ConvLNP") and cfg.get("notebook"):
        kw["Interpolator"] = SetConv
    cls = dict(CNP=CNP, LNP=LNP, AttnCNP=AttnCNP, AttnLNP=AttnLNP, ConvCNP=ConvCNP, ConvLNP=ConvLNP, GridConvCNP=GridConvCNP,
               GridConvLNP=GridConvLNP)[fam]
    return cls(cfg["x_dim"], cfg["y_dim"], **kw)


def loss_for(name, reduction=None):
    return dict(cnpf=CNPFLoss, nll=NLLLossLNPF, elbo=ELBOLossLNPF)[name](reduction=reduction)
