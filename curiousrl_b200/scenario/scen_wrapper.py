"""Scenario capability interface (reference: CuriousRL/scenario/scen_wrapper.py:2-16)."""
from abc import ABC, abstractmethod


class ScenarioWrapper(ABC):
    """What a solver may ask a scenario before accepting it."""

    @abstractmethod
    def with_model(self):
        ...

    @abstractmethod
    def is_action_discrete(self):
        ...

    @abstractmethod
    def is_output_image(self):
        ...

    def play(self):
        raise NotImplementedError
