"""`Data`: a bundle of transitions (reference: CuriousRL/data/data.py:10-162).

Same constructor keys (`state`, `action`, `next_state`, `reward`, `done_flag`), single / multiple data modes, `cat`,
`clone`, `len`.  Tensors live on the device the package computes on (`curiousrl_b200.utils.config.global_config`);
NumPy inputs become float32 tensors there, as in the reference (`torch.from_numpy(...).float().cuda()`).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import Tensor

from ..utils.config import global_config

ACCESSIBLE_KEY = {'state', 'action', 'next_state', 'reward', 'done_flag'}


class Data(object):
    def __init__(self, **kwargs):
        self._data_dict = {}
        for key, value in kwargs.items():
            if key not in ACCESSIBLE_KEY:
                raise Exception("\"" + key + "\" is not an accessible key in data!")
            if isinstance(value, Tensor):
                self._data_dict[key] = value
            else:
                t = torch.from_numpy(np.asarray(value))
                self._data_dict[key] = t.float().to(global_config.device) if global_config.is_cuda else t
        if 'action' in self._data_dict and self._data_dict['action'].dim() == 1:      # single data mode
            for key in self._data_dict:
                self._data_dict[key] = torch.unsqueeze(torch.squeeze(self._data_dict[key]), 0)

    def __len__(self):
        return len(next(iter(self._data_dict.values())))

    def __str__(self):
        return "".join(key + ":\n " + str(value) + "\n" for key, value in self._data_dict.items())

    state = property(lambda self: self._data_dict['state'])
    action = property(lambda self: self._data_dict['action'])
    next_state = property(lambda self: self._data_dict['next_state'])
    reward = property(lambda self: self._data_dict['reward'])
    done_flag = property(lambda self: self._data_dict['done_flag'])

    def cat(self, datas) -> "Data":
        """A new Data made of this one followed by `datas` (one Data or a tuple of them) - same keys required."""
        new_data = self.clone()
        for data in (datas if isinstance(datas, tuple) else (datas,)):
            if data._data_dict.keys() != self._data_dict.keys():
                raise Exception("Cannot perform \"cat\" among Data instances with different keys!")
            for key in data._data_dict:
                new_data._data_dict[key] = torch.cat([new_data._data_dict[key], data._data_dict[key]], dim=0)
        return new_data

    def clone(self) -> "Data":
        out = Data()
        out._data_dict = {key: value.clone() for key, value in self._data_dict.items()}
        return out
