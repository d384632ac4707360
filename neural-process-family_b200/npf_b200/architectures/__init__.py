"""Building blocks obeying upstream's sub-module factory protocol (SURVEY.md section 8b tier 2); each can also be injected
into the unmodified reference classes."""
from .mlp import MLP
from .encoders import DiscardIthArg, MergeFlatInputs, discard_ith_arg, merge_flat_input
from .setcnn import ExpRBF, SetConv
from .cnn import CNN, ResConvBlock
from .attention import DotAttender, MultiheadAttender, TransformerAttender, get_attender
from .selfattn import SelfAttention

__all__ = ["MLP", "merge_flat_input", "discard_ith_arg", "MergeFlatInputs", "DiscardIthArg", "SetConv", "ExpRBF", "ResConvBlock", "CNN",
           "get_attender", "DotAttender", "MultiheadAttender", "TransformerAttender", "SelfAttention"]
