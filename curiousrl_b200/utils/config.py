"""Global configuration singleton (reference: CuriousRL/utils/config.py:6-32): the device flag and the random seed."""
import numpy as np
import torch


class GlobalConfiguration(object):
    def __new__(cls):
        if not hasattr(cls, '_instance'):
            cls._instance = super(GlobalConfiguration, cls).__new__(cls)
            cls._instance._is_cuda = True
        return cls._instance

    @property
    def is_cuda(self):
        return self._is_cuda

    @property
    def device(self):
        return torch.device("cuda", torch.cuda.current_device()) if self._is_cuda else torch.device("cpu")

    def set_random_seed(self, seed: int):
        torch.manual_seed(seed)
        np.random.seed(seed)
        self.random_seed = seed

    def set_is_cuda(self, is_cuda: bool):
        self._is_cuda = is_cuda


global_config = GlobalConfiguration()
