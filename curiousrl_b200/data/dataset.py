"""`Dataset`: fixed-size replay buffer of `Data` (reference: CuriousRL/data/dataset.py:10-137).

A ring buffer per key, allocated on the first update on the package's device; the oldest entries are overwritten.
`fetch_all_data` returns the whole buffer (zeros where nothing has been written yet, as the reference does).
"""
from __future__ import annotations

import numpy as np
import torch

from ..utils.config import global_config
from .data import Data


class Dataset(object):
    def __init__(self, buffer_size, state_dim, action_dim: int):
        if isinstance(state_dim, int):
            state_dim = (state_dim,)
        self._buffer_size = int(buffer_size)
        self._state_dim = state_dim
        self._action_dim = action_dim
        self._update_key = 0
        self._total_update = 0            # number of data obtained so far

    def _allocate(self, data: Data):
        dev = global_config.device if global_config.is_cuda else torch.device("cpu")
        self._dataset_dict = {}
        for key, value in data._data_dict.items():
            if key in ("state", "next_state"):
                self._dataset_dict[key] = torch.zeros((self._buffer_size, *value.shape[1:]), device=dev)
            elif key == "action":
                self._dataset_dict[key] = torch.zeros((self._buffer_size, value.shape[1]), device=dev)
            elif key == "reward":
                self._dataset_dict[key] = torch.zeros((self._buffer_size,), device=dev)
            elif key == "done_flag":
                self._dataset_dict[key] = torch.zeros((self._buffer_size,), dtype=torch.bool, device=dev)

    def update_dataset(self, data: Data):
        """Append `data`; when the buffer is full the oldest entries are replaced (dataset.py:41-82)."""
        if self._total_update == 0:
            self._allocate(data)
        n = len(data)
        if n > self._buffer_size:
            raise Exception("The new data (%d) do not fit the buffer (%d)" % (n, self._buffer_size))
        self._total_update += n
        head = min(n, self._buffer_size - self._update_key)
        for key, buf in self._dataset_dict.items():
            src = data._data_dict[key].to(buf.device)
            buf[self._update_key:self._update_key + head] = src[:head]
            if n > head:
                buf[:n - head] = src[head:]
        self._update_key = (self._update_key + n) % self._buffer_size

    def get_current_buffer_size(self):
        return min(self._total_update, self._buffer_size)

    def fetch_all_data(self) -> Data:
        return self.fetch_data_by_index(list(range(self._buffer_size)))

    def fetch_data_by_index(self, index) -> Data:
        return Data(**{key: buf[index] for key, buf in self._dataset_dict.items()})

    def fetch_data_randomly(self, num_of_data: int) -> Data:
        have = self.get_current_buffer_size()
        if num_of_data > have:
            raise Exception("The current buffer size is %d. Number of random data size is %d." % (have, num_of_data) +
                            "The latter must be less or equal than the former.")
        return self.fetch_data_by_index(np.random.choice(have, size=num_of_data, replace=False))
