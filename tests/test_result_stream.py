"""Batched result / replay channel: utils/result_stream.py + Logger.save_batch /
read_batch - the binary counterpart of Logger.save_to_json / read_from_json
(/root/reference/CuriousRL/utils/Logger.py:66-99).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from curiousrl_b200.utils.Logger import Logger
from curiousrl_b200.utils.result_stream import ResultStream


def _record(rng, B=5, T=7, d=4):
    return dict(trajectory=rng.standard_normal((B, T, d)), obj_fun_value=rng.standard_normal(B),
                iterations=rng.integers(1, 50, B).astype(np.int32), converged=(rng.random(B) > 0.5).astype(np.uint8))


def test_append_read_roundtrip_bit_exact(tmp_path):
    rng = np.random.default_rng(0)
    st = ResultStream("run", root=str(tmp_path))
    recs = [_record(rng, B) for B in (5, 1, 9)]                 # ragged batch sizes between records
    for k, r in enumerate(recs):
        assert st.append(meta={"k": k}, **r) == k
    assert len(st) == 3
    for k, r in enumerate(recs):
        got = st.read(k)
        assert got["meta"] == {"k": k}
        for name, a in r.items():
            assert got[name].dtype == a.dtype and np.array_equal(got[name], a)
            assert isinstance(got[name], np.memmap)             # lazily mapped: only touched rows are read
    assert np.array_equal(st.read(-1)["trajectory"], recs[-1]["trajectory"])
    assert not isinstance(st.read(-1, mmap=False)["trajectory"], np.memmap)


def test_read_semantics_follow_read_from_json(tmp_path):
    st = ResultStream("empty", root=str(tmp_path))
    with pytest.raises(Exception, match="No result stream"):
        st.read()
    st.append(**_record(np.random.default_rng(1)))
    with pytest.raises(Exception, match="exceeds the maximum iteration number"):   # Logger.py:95
        st.read(1)
    with pytest.raises(ValueError):
        st.append(trajectory=np.zeros((3, 2, 2)), iterations=np.zeros(4, dtype=np.int32))   # mismatched batch axes
    with pytest.raises(ValueError):
        st.append()


def test_reference_encoding_of_one_trajectory(tmp_path):
    """`trajectory_as_json` = what the reference's `read_from_json(...)` returns for that trajectory:
    {"trajectory": [T][d][1] lists}, i.e. json.dumps(X.tolist()) of an X [T, d, 1] (basic_ilqr.py:93)."""
    rng = np.random.default_rng(2)
    r = _record(rng)
    st = ResultStream("enc", root=str(tmp_path))
    st.append(**r)
    got = st.trajectory_as_json(-1, index=3)
    X = r["trajectory"][3][:, :, None]
    assert json.loads(json.dumps({"trajectory": X.tolist()})) == got
    assert np.asarray(got["trajectory"]).shape == (7, 4, 1)


def test_torch_tensors_and_partial_records(tmp_path):
    st = ResultStream("t", root=str(tmp_path))
    t = torch.arange(24, dtype=torch.float64).reshape(2, 3, 4)
    st.append(trajectory=t, obj_fun_value=torch.tensor([1.0, 2.0]), skipped=None)
    got = st.read()
    assert set(got) == {"trajectory", "obj_fun_value", "meta"}
    assert np.array_equal(got["trajectory"], t.numpy())
    # a record whose index line is missing (writer died between the arrays and the index) is invisible
    os.makedirs(os.path.join(st.path, "1"))
    np.save(os.path.join(st.path, "1", "trajectory.npy"), np.zeros((2, 3, 4)))
    assert len(st) == 1
    assert np.array_equal(st.read(-1)["trajectory"], t.numpy())


def test_logger_batch_channel(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    lg = Logger()
    saved = (lg.IS_SAVE_JSON, lg.IS_USE_LOGGER, lg.FOLDER_NAME, lg.memory_batch)
    try:
        r = _record(np.random.default_rng(3))
        lg.set_is_save_json(False)
        assert lg.save_batch(**r) is None
        mem = lg.read_batch()
        assert mem["trajectory"] is r["trajectory"]                        # by reference, no copy
        lg.set_folder_name("stream_test").set_is_save_json(True)
        assert lg.save_batch(meta={"solver": "BasiciLQR"}, **r) == 0
        assert lg.save_batch(**_record(np.random.default_rng(4))) == 1
        assert os.path.isfile(os.path.join("logs", "stream_test", "stream", "index.jsonl"))
        assert np.array_equal(lg.read_batch(no_iter=0)["trajectory"], r["trajectory"])
        assert lg.read_batch(no_iter=0)["meta"] == {"solver": "BasiciLQR"}
        assert np.array_equal(lg.read_batch("stream_test", 0)["iterations"], r["iterations"])
        lg.set_is_use_logger(False)
        with pytest.raises(Exception, match="Logger must be used"):
            lg.save_batch(**r)
    finally:
        lg.set_is_use_logger(saved[1])
        lg.IS_SAVE_JSON, lg.FOLDER_NAME, lg.memory_batch = saved[0], saved[2], saved[3]


def test_round_trip_property(tmp_path):
    """Property test: any sequence of records with arbitrary batch sizes, trailing shapes and dtypes reads back bit for bit,
    in order, under both addressing modes (index and -1 = last)."""
    from hypothesis import given, settings, strategies as st, HealthCheck
    import itertools
    counter = itertools.count()
    dtypes = st.sampled_from([np.float64, np.float32, np.int32, np.uint8])
    shape_tail = st.lists(st.integers(1, 4), min_size=0, max_size=2)
    record = st.tuples(st.integers(1, 6), st.lists(st.tuples(shape_tail, dtypes), min_size=1, max_size=3))

    @settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(st.lists(record, min_size=1, max_size=4), st.integers(0, 2 ** 31 - 1))
    def check(records, seed):
        rng = np.random.default_rng(seed)
        stream = ResultStream("prop_%d" % next(counter), root=str(tmp_path))
        written = []
        for B, fields in records:
            arrays = {"f%d" % i: (rng.standard_normal((B, *tail)) * 100).astype(dt) for i, (tail, dt) in enumerate(fields)}
            stream.append(meta={"B": B}, **arrays)
            written.append(arrays)
        assert len(stream) == len(written)
        for k, arrays in enumerate(written):
            got = stream.read(k)
            assert got["meta"] == {"B": next(iter(arrays.values())).shape[0]}
            for name, a in arrays.items():
                assert got[name].dtype == a.dtype and got[name].shape == a.shape and np.array_equal(got[name], a)
        last = stream.read(-1)
        for name, a in written[-1].items():
            assert np.array_equal(last[name], a)

    check()
