from .attention import *
from .cnn import *
from .encoders import *
from .mlp import *
from .selfattn import *
from .setcnn import *
