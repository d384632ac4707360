from .attnnp import *
from .base import *
from .convnp import *
from .gridconvnp import *
from .np import *
