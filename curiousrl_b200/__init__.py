"""curiousrl_b200 - B200-native batched iLQR behind the CuriousRL solver/scenario API.

    from curiousrl_b200.algorithm.ilqr_solver import BasiciLQR
    from curiousrl_b200.scenario.dynamic_model import ThreeLinkPlanarManipulator

Importing the package does not need a GPU; constructing a solver session (`algo.init(...)`)
does, and raises otherwise - there is no CPU fallback.
"""
from .utils.Logger import logger
from .algorithm.algo_wrapper import AlgoWrapper
from .scenario.scen_wrapper import ScenarioWrapper

__version__ = "0.1.0"
