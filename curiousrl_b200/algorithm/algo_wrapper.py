"""Base class of every algorithm (reference: CuriousRL/algorithm/algo_wrapper.py:4-28)."""
from abc import ABC, abstractmethod

from ..utils.Logger import logger


class AlgoWrapper(ABC):
    """Stores and echoes the constructor arguments; subclasses implement init(scenario) and solve()."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        self.print_params()

    @abstractmethod
    def init(self, scenario):
        ...

    @abstractmethod
    def solve(self):
        ...

    def print_params(self):
        for key, value in self.kwargs.items():
            logger.info("[+] %s = %s" % (key, value))
