from .utils import *
from .cbo_functions import *
from .cost_functions import *
from .graph_functions import *
from .causal_kernels import *
from .causal_acquisition_functions import *
