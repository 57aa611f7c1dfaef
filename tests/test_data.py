"""`Data` / `Dataset` (reference CuriousRL/data/{data,dataset}.py): CPU semantics, and - where the reference
tree is mounted - the same sequence of updates through the reference's own classes."""
import numpy as np
import pytest
import torch

from curiousrl_b200.data import Data, Dataset
from curiousrl_b200.utils.config import global_config
from oracle import ref_shim


@pytest.fixture(autouse=True)
def _cpu_mode():
    global_config.set_is_cuda(False)
    yield
    global_config.set_is_cuda(True)


def test_data_modes_and_cat():
    single = Data(state=np.arange(4.0), action=np.array([1.0, 2.0]), reward=np.array(3.0))
    assert len(single) == 1 and single.state.shape == (1, 4) and single.action.shape == (1, 2)
    multi = Data(state=np.arange(12.0).reshape(3, 4), action=np.ones((3, 2)))
    assert len(multi) == 3
    both = multi.cat(Data(state=np.zeros((2, 4)), action=np.zeros((2, 2))))
    assert len(both) == 5 and len(multi) == 3                       # cat returns a new object
    many = multi.cat((multi, multi))
    assert len(many) == 9
    with pytest.raises(Exception):
        multi.cat(Data(state=np.zeros((2, 4))))                      # different keys
    with pytest.raises(Exception):
        Data(observation=np.zeros(3))                                # not an accessible key
    clone = multi.clone()
    clone.state[0, 0] = 99.0
    assert multi.state[0, 0] == 0.0


def _fill(ds_cls, data_cls, updates, buffer_size, dim):
    ds = ds_cls(buffer_size, dim, 1)
    for u in updates:
        ds.update_dataset(data_cls(state=u))
    return ds


def test_ring_buffer_wraps_like_the_reference():
    rng = np.random.default_rng(0)
    updates = [rng.normal(size=(k, 3)) for k in (6, 6, 5, 10, 1, 3)]
    ds = _fill(Dataset, Data, updates, 10, 3)
    # hand-rolled ring
    ring, key = np.zeros((10, 3), dtype=np.float32), 0          # the buffers are float32, as in the reference
    for u in updates:
        for row in u:
            ring[key] = row
            key = (key + 1) % 10
    assert np.array_equal(ds.fetch_all_data().state.numpy(), ring)
    assert ds.get_current_buffer_size() == 10 and ds._update_key == key
    assert np.array_equal(ds.fetch_data_by_index([1, 2, 5]).state.numpy(), ring[[1, 2, 5]])
    np.random.seed(1)
    got = ds.fetch_data_randomly(4).state.numpy()
    assert got.shape == (4, 3) and all(any(np.array_equal(g, r) for r in ring) for g in got)
    with pytest.raises(Exception):
        ds.fetch_data_randomly(11)
    part = _fill(Dataset, Data, updates[:1], 10, 3)
    assert part.get_current_buffer_size() == 6
    with pytest.raises(Exception):
        part.fetch_data_randomly(7)


@pytest.mark.skipif(not ref_shim.available(), reason="no reference tree on this machine")
def test_same_contents_as_the_reference_classes():
    ref_shim.install()
    from CuriousRL.utils.config import global_config as ref_config
    ref_config.set_is_cuda(False)
    from CuriousRL.data import Data as RefData, Dataset as RefDataset
    rng = np.random.default_rng(5)
    updates = [rng.normal(size=(k, 4)) for k in (7, 7, 3, 12, 2)]
    ours = _fill(Dataset, Data, updates, 12, 4)
    theirs = _fill(RefDataset, RefData, updates, 12, 4)
    assert torch.equal(ours.fetch_all_data().state.double(), theirs.fetch_all_data().state.double())
    assert ours._update_key == theirs._update_key and ours.get_current_buffer_size() == theirs.get_current_buffer_size()
