from .option import Opt
from .stats import Stats
from .solution import aesol, daesol
from .nr import nr_method
from .rodas import Rodas
from .rodas_param import Rodas_param
from .fdae import fdae_solver
from .implicit import backward_euler, implicit_trapezoid
from .rodas_event import ComponentEvents
from .laesolver import BatchedLU, lu_decomposition, solve
from .sicnm import sicnm
