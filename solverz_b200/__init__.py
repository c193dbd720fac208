"""solverz_b200 -- B200-native batched Newton / Rosenbrock backend with the Solverz API surface."""
from .sym import (Var, Param, TimeSeriesParam, AliasVar, Set, Idx, Sum, Mat_Mul, sin, cos, exp, ln, Abs, atan2, tan,
                  sqrt, Sign, heaviside, Saturation, In, GreaterThan, LessThan, And, Or, Not, Min, switch, AntiWindUp)
from .sym import t as t_sym
from .equations import Eqn, Ode, LoopEqn, LoopOde, Model, AE, DAE, FDAE
from .variables import Vars, TimeVars, Address

__version__ = "0.1.0"
from .code_printer import module_printer, made_numerical
from .solvers import sicnm, Opt, Stats, aesol, daesol, nr_method, Rodas, fdae_solver, backward_euler, implicit_trapezoid, ComponentEvents, BatchedLU, lu_decomposition, solve
from .io import save, load, save_result, load_result
