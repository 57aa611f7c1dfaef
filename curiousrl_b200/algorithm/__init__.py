from .algo_wrapper import AlgoWrapper
