"""Model classes of the family (constructor-compatible with upstream ``npf.neuralproc``): every class below runs its
forward / backward through the CUDA entry points of libnpf_b200.so."""
from .base import BoundedSigmoid, LatentNeuralProcessFamily, MinSoftplus, NeuralProcessFamily
from .np import CNP, LNP
from .attnnp import AttnCNP, AttnLNP
from .convnp import ConvCNP, ConvLNP
from .gridconvnp import GridConvCNP, GridConvLNP

__all__ = ["NeuralProcessFamily", "LatentNeuralProcessFamily", "MinSoftplus", "BoundedSigmoid", "CNP", "LNP", "AttnCNP", "AttnLNP",
           "ConvCNP", "ConvLNP", "GridConvCNP", "GridConvLNP"]
