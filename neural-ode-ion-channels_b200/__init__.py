"""B200-native batched ``odeint`` for the hERG/IKr neural-ODE models (NN-f / NN-d).

Public surface (mirrors what the reference scripts use on this path):

* ``odeint(func, y0, t, *, rtol, atol, method, options)`` -- torchdiffeq-compatible entry point
* ``integrate(...)``                                   -- same, plus fused current / loss epilogue
* ``loss_and_grad(...)``                               -- fused loss + gradient through the solver
* ``ODEFunc`` / ``ODEFuncNNf`` / ``ODEFuncNNd``         -- the reference's ODE-func modules
* ``ARCHITECTURES`` / ``build_net``                    -- architectures/s00..s11
* ``protocols``                                        -- voltage-clamp protocol tables
"""
from . import parallel, protocols, reporting  # noqa: F401
from .models import (ARCHITECTURES, PARAMETER_SETS, ODEFunc, ODEFuncNNd, ODEFuncNNf,  # noqa: F401
                     build_net, load_weights)
from .solver import IkrResult, describe, integrate, integrate_many, odeint  # noqa: F401
from .adjoint import loss_and_grad  # noqa: F401
from .hh import HHPopulationModel, integrate_hh  # noqa: F401
from .markov import MARKOV_B06, MarkovGroundTruth, integrate_markov  # noqa: F401
from .regression import fit_regression, mse_loss_and_grad, save_checkpoint  # noqa: F401

__all__ = ['odeint', 'integrate', 'loss_and_grad', 'integrate_hh', 'HHPopulationModel', 'integrate_markov', 'MarkovGroundTruth', 'MARKOV_B06', 'mse_loss_and_grad', 'fit_regression',
           'save_checkpoint', 'integrate_many', 'describe', 'IkrResult', 'ODEFunc', 'ODEFuncNNf', 'ODEFuncNNd',
           'ARCHITECTURES', 'PARAMETER_SETS', 'build_net', 'load_weights', 'protocols', 'parallel', 'reporting']
