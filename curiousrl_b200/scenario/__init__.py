from .scen_wrapper import ScenarioWrapper
