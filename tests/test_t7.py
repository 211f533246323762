"""Torch7 container codec (visdial_b200/t7.py) against byte strings assembled by hand from the published format
[upstream torch7 File.lua / generic Tensor.c / Storage.c] and by round trips.  CPU only; parity unpinned (no torch7)."""
import io
import struct

import numpy as np
import pytest

from visdial_b200 import t7

i32 = lambda v: struct.pack("<i", v)
i64 = lambda v: struct.pack("<q", v)
f64 = lambda v: struct.pack("<d", v)
s_ = lambda s: i32(len(s)) + s.encode()


def _float_tensor_bytes(idx_t, idx_s, shape, stride, data):
    b = i32(4) + i32(idx_t) + s_("V 1") + s_("torch.FloatTensor") + i32(len(shape))
    b += b"".join(i64(x) for x in shape) + b"".join(i64(x) for x in stride) + i64(1)
    b += i32(4) + i32(idx_s) + s_("V 1") + s_("torch.FloatStorage") + i64(len(data)) + np.asarray(data, "<f4").tobytes()
    return b


def test_reads_hand_assembled_checkpoint_table():
    # {modelW = FloatTensor{1,2,3,4,5,6} viewed 2x3, optims = {learningRate = 0.001, t = 7}, modelParams = {encoder = 'lf-ques', useIm = false}}
    blob = i32(3) + i32(1) + i32(3)
    blob += i32(2) + s_("modelW") + _float_tensor_bytes(2, 3, (2, 3), (3, 1), [1, 2, 3, 4, 5, 6])
    blob += i32(2) + s_("optims") + i32(3) + i32(4) + i32(2) + i32(2) + s_("learningRate") + i32(1) + f64(0.001) \
        + i32(2) + s_("t") + i32(1) + f64(7.0)
    blob += i32(2) + s_("modelParams") + i32(3) + i32(5) + i32(2) + i32(2) + s_("encoder") + i32(2) + s_("lf-ques") \
        + i32(2) + s_("useIm") + i32(5) + i32(0)
    ck = t7.load(io.BytesIO(blob))
    assert ck["modelW"].dtype == np.float32 and ck["modelW"].tolist() == [[1, 2, 3], [4, 5, 6]]
    assert ck["optims"] == {"learningRate": 0.001, "t": 7}
    assert ck["modelParams"] == {"encoder": "lf-ques", "useIm": False}


def test_strided_view_offset_and_shared_reference():
    # a transposed view with a storage offset, and the same tensor object referenced twice (second time by index only)
    storage = list(range(10))
    t = i32(4) + i32(2) + s_("V 1") + s_("torch.DoubleTensor") + i32(2) + i64(2) + i64(3) + i64(1) + i64(2) + i64(3) \
        + i32(4) + i32(3) + s_("V 1") + s_("torch.DoubleStorage") + i64(10) + np.asarray(storage, "<f8").tobytes()
    blob = i32(3) + i32(1) + i32(2) + i32(1) + f64(1) + t + i32(1) + f64(2) + i32(4) + i32(2)
    out = t7.load(io.BytesIO(blob))
    assert out[1].tolist() == [[2, 4, 6], [3, 5, 7]]          # offset 3 (1-based) -> element 2; strides (1, 2)
    assert out[2] is out[1]
    assert t7.as_list(out)[0] is out[1]


def test_legacy_header_without_version_and_nil():
    blob = i32(3) + i32(1) + i32(2) + i32(2) + s_("a") + i32(0) + i32(2) + s_("w") \
        + i32(4) + i32(2) + s_("torch.LongTensor") + i32(1) + i64(2) + i64(1) + i64(1) \
        + i32(4) + i32(3) + s_("torch.LongStorage") + i64(2) + np.asarray([5, 9], "<i8").tobytes()
    out = t7.load(io.BytesIO(blob))
    assert out["a"] is None and out["w"].tolist() == [5, 9] and out["w"].dtype == np.int64


def test_writer_emits_the_documented_bytes():
    buf = io.BytesIO()
    t7.save(buf, {"w": np.asarray([1.5, 2.5], np.float32)})
    want = i32(3) + i32(1) + i32(1) + i32(2) + s_("w") + _float_tensor_bytes(2, 3, (2,), (1,), [1.5, 2.5])
    assert buf.getvalue() == want


def test_round_trip_of_a_checkpoint_shaped_table():
    rng = np.random.default_rng(0)
    w = rng.standard_normal(1000).astype(np.float32)
    ck = {"modelW": t7.CudaTensor(w), "optims": {"learningRate": 4e-4, "t": 12, "m": t7.CudaTensor(w * 2), "v": t7.CudaTensor(w * w)},
          "modelParams": {"encoder": "mn-att-ques-im-hist", "decoder": "disc", "rnnHiddenSize": 512, "dropout": 0.5, "useIm": True},
          "layout": [{"name": "wordEmbed.weight", "offset": 0, "rows": 10, "cols": 3}], "ids": np.arange(6, dtype=np.int64).reshape(2, 3)}
    buf = io.BytesIO()
    t7.save(buf, ck)
    assert b"torch.CudaTensor" in buf.getvalue() and b"torch.CudaStorage" in buf.getvalue()
    out = t7.load(io.BytesIO(buf.getvalue()))
    assert np.array_equal(out["modelW"], w) and np.array_equal(out["optims"]["v"], w * w)
    assert out["optims"]["t"] == 12 and out["optims"]["learningRate"] == 4e-4
    assert out["modelParams"] == ck["modelParams"]
    assert t7.as_list(out["layout"])[0]["name"] == "wordEmbed.weight"
    assert out["ids"].tolist() == [[0, 1, 2], [3, 4, 5]]


def test_errors():
    with pytest.raises(t7.T7Error):
        t7.load(io.BytesIO(i32(2) + i32(10) + b"abc"))                      # truncated string
    with pytest.raises(t7.T7Error):
        t7.load(io.BytesIO(i32(6)))                                          # a function: not in the schema
    with pytest.raises(t7.T7Error):
        t7.load(io.BytesIO(i32(4) + i32(1) + s_("V 1") + s_("nn.Linear")))  # a module: not in the schema
    with pytest.raises(t7.T7Error):
        t7.save(io.BytesIO(), {"x": object()})


def test_reads_a_torch7_written_checkpoint():
    """Pins t7.py against torch7's own serialiser and records the nngraph getParameters() order WHEN the artefacts exist
    (tests/golden/external/README.md).  No LuaJIT / Torch7 in the build container."""
    import os
    import pytest
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "external")
    path = os.path.join(here, "model_tiny.t7")
    if not os.path.exists(path):
        pytest.skip("external artefact model_tiny.t7 not present (see tests/golden/external/README.md)")
    from visdial_b200 import t7
    ck = t7.load(path)
    assert "modelW" in ck and "modelParams" in ck
    w = np.asarray(ck["modelW"])
    assert w.ndim == 1 and w.dtype == np.float32 and np.isfinite(w).all()
    order = os.path.join(here, "wrapper_params.txt")
    if os.path.exists(order):
        sizes = [int(np.prod([int(x) for x in line.split()[1].split("x")])) for line in open(order) if line.strip()]
        assert sum(sizes) == w.size            # the flattened vector is the concatenation of wrapper:parameters() in that order
