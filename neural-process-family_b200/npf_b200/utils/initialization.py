"""Parameter initialisation with the same *effective* distributions as upstream npf/utils/initialization.py.

Upstream's ``weights_init`` is a de-facto no-op whenever the module it is called on defines
``reset_parameters`` (its skip test looks at the root, initialization.py:18), which is the case
for every class of the library.  The net effect that matters: Linear layers owned by an ``MLP`` get
``linear_init`` (kaiming-uniform for relu, zero bias); everything else keeps the torch default init.
"""
import torch.nn as nn

__all__ = ["weights_init", "linear_init", "init_param_"]


def weights_init(module, **kwargs):
    """Kept for API compatibility (upstream initialization.py:7-31).  Marks the module as reset; children keep
    the initialisation they were constructed with, exactly as upstream effectively does."""
    module.is_resetted = True


def linear_init(module, activation="relu"):
    """upstream initialization.py:67-94: zero bias; kaiming-uniform (relu / leaky-relu), xavier-uniform otherwise."""
    if module.bias is not None:
        module.bias.data.zero_()
    w = module.weight
    if activation is None:
        return nn.init.xavier_uniform_(w)
    name = activation if isinstance(activation, str) else type(activation).__name__.lower()
    if name in ("relu",):
        return nn.init.kaiming_uniform_(w, nonlinearity="relu")
    if name in ("leaky_relu", "leakyrelu"):
        slope = 0 if isinstance(activation, str) else activation.negative_slope
        return nn.init.kaiming_uniform_(w, a=slope, nonlinearity="leaky_relu")
    if name in ("sigmoid", "tanh", "softmax"):
        gain = nn.init.calculate_gain("tanh" if name == "tanh" else "sigmoid")
        return nn.init.xavier_uniform_(w, gain=gain)
    raise ValueError(f"Unknown activation {activation}")


def init_param_(param, activation=None, is_positive=False, bound=0.05, shift=0):
    """upstream initialization.py:97-124 (uniform in +-bound*gain, optionally positive / shifted)."""
    gain = 1.0 if activation is None else nn.init.calculate_gain(
        activation if isinstance(activation, str) else type(activation).__name__.lower())
    if is_positive:
        nn.init.uniform_(param, 1e-5 + shift, bound * gain + shift)
    else:
        nn.init.uniform_(param, -bound * gain + shift, bound * gain + shift)
