"""npf_b200 -- Blackwell (B200, sm_100a) forward/backward path for the Neural Process Family.

Drop-in for the model / loss classes of YannDubs/Neural-Process-Family (``from npf import CNP, ...``): same
constructors, same ``forward(X_cntxt, Y_cntxt, X_trgt, Y_trgt=None)``, same ``state_dict`` keys, same returned
``(p_yCc, z_samples, q_zCc, q_zCct)``.  All arithmetic runs in hand-written CUDA kernels behind the C ABI of
``lib/libnpf_b200.so`` (include/npf_b200.h); there is no CPU path.
"""
from . import ops
from .losses import *
from .neuralproc import *
from .graph import GraphedStep, PipelinedStep
from .ops import get_precision, set_precision

__version__ = "0.1.0"
