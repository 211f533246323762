"""visdial_b200/h5lite.py: the HDF5 subset the reference's data files use (data/prepro.py:264-277,
dataloader.lua:37-129).  CPU only; parity unpinned (no h5py / libhdf5 in the image): structure checks against the
file-format specification's byte layout plus writer/reader round trips."""
import struct

import numpy as np
import pytest

from visdial_b200 import h5lite
from visdial_b200.synthetic import make_corpus


def test_round_trip_many_datasets_and_dtypes(tmp_path):
    rng = np.random.default_rng(0)
    data = {"ques_train": rng.integers(0, 9000, size=(7, 10, 20)).astype(np.uint32),
            "ques_length_train": rng.integers(0, 21, size=(7, 10)).astype(np.uint32),
            "img_pos_train": np.arange(7, dtype=np.int64),
            "images_train": rng.standard_normal((7, 8, 3, 3)).astype(np.float32),
            "d64": rng.standard_normal((5,)).astype(np.float64), "i32": np.array([[-5, 7]], dtype=np.int32),
            "u8": np.arange(6, dtype=np.uint8).reshape(2, 3), "empty": np.zeros((0, 4), np.uint32)}
    for i in range(14):                                    # > 8 names: several symbol-table nodes under the B-tree
        data["extra_%02d" % i] = np.full((2,), i, np.uint32)
    path = str(tmp_path / "a.h5")
    h5lite.write(path, data)
    out = h5lite.read(path)
    assert sorted(out) == sorted(data)
    for k, v in data.items():
        assert out[k].dtype == v.dtype and out[k].shape == v.shape and np.array_equal(out[k], v), k
    assert list(h5lite.read(path, ["/ques_train"])) == ["ques_train"]
    with pytest.raises(h5lite.H5Error):
        h5lite.read(path, ["nope"])


def test_file_structure_follows_the_specification(tmp_path):
    path = str(tmp_path / "s.h5")
    a = np.array([[1, 2, 3], [4, 5, 0x01020304]], dtype=np.uint32)
    h5lite.write(path, {"opt_list_val": a})
    b = open(path, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0                    # signature, superblock version 0
    assert b[13] == 8 and b[14] == 8                                       # sizes of offsets / lengths
    eof = struct.unpack_from("<Q", b, 24 + 16)[0]
    assert eof == len(b)                                                   # end-of-file address
    root_hdr = struct.unpack_from("<Q", b, 56 + 8)[0]
    assert b[root_hdr] == 1                                                # version-1 object header
    assert struct.unpack_from("<H", b, root_hdr + 16)[0] == 0x0011         # its only message: symbol table
    bt, heap = struct.unpack_from("<QQ", b, root_hdr + 24)
    assert b[bt:bt + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP"
    snod = struct.unpack_from("<Q", b, bt + 24 + 8)[0]
    assert b[snod:snod + 4] == b"SNOD" and struct.unpack_from("<H", b, snod + 6)[0] == 1
    assert a.astype("<u4").tobytes() in b                                  # raw little-endian, contiguous
    assert b"opt_list_val\x00" in b


def test_big_endian_and_signed_types_are_decoded(tmp_path):
    path = str(tmp_path / "be.h5")
    a = np.array([1, -2, 300000], dtype=np.int32)
    h5lite.write(path, {"x": a})
    b = bytearray(open(path, "rb").read())
    i = b.index(a.astype("<i4").tobytes())
    b[i:i + 12] = a.astype(">i4").tobytes()                                # store big-endian ...
    j = b.index(struct.pack("<BBBBI", 0x10, 0x08, 0, 0, 4))                # ... and flip the byte-order bit of the datatype
    b[j + 1] |= 1
    open(path, "wb").write(bytes(b))
    out = h5lite.read(path)["x"]
    assert out.tolist() == a.tolist() and out.dtype == np.int32


@pytest.mark.parametrize("gzip", [False, True])
def test_chunked_layout_with_ragged_edge_chunks(tmp_path, gzip):
    rng = np.random.default_rng(1)
    feats = rng.standard_normal((5, 6, 7)).astype(np.float32)
    toks = rng.integers(0, 100, size=(9, 10)).astype(np.uint32)
    path = str(tmp_path / "c.h5")
    h5lite.write(path, {"images_val": feats, "ans_val": toks}, chunks={"images_val": (2, 4, 7), "ans_val": (4, 4)}, gzip=gzip)
    out = h5lite.read(path)
    assert np.array_equal(out["images_val"], feats) and np.array_equal(out["ans_val"], toks)
    raw = open(path, "rb").read()
    assert raw.count(b"TREE") == 3                                          # the group's B-tree + one chunk index per dataset


def test_prepro_style_files_feed_the_dataloader_dict(tmp_path):
    """visdial_data.h5 / data_img.h5 as prepro.py and prepro_img_*.lua name their datasets -> the per-split dict."""
    params = {"vocabSize": 100, "encoder": "mn-att-ques-im-hist", "imgFeatureSize": 8, "imgSpatialSize": 3,
              "maxQuesCount": 10, "numOptions": 100}
    raw = make_corpus(params, 9, 30, seed=8)
    qa = {k + "_train": v.astype(np.uint32) for k, v in raw.items() if k != "images"}        # prepro.py:267: dtype='uint32'
    h5lite.write(str(tmp_path / "visdial_data.h5"), qa)
    h5lite.write(str(tmp_path / "data_img.h5"), {"images_train": raw["images"]})
    got = h5lite.split(h5lite.read(str(tmp_path / "visdial_data.h5")), "train")
    got.update(h5lite.split(h5lite.read(str(tmp_path / "data_img.h5")), "train"))
    assert sorted(got) == sorted(raw)
    for k in raw:
        assert np.array_equal(got[k], raw[k]), k
    assert got["ques"].dtype == np.uint32 and got["images"].dtype == np.float32


def test_unsupported_files_are_refused(tmp_path):
    p = str(tmp_path / "bad.h5")
    open(p, "wb").write(b"not hdf5 at all" * 10)
    with pytest.raises(h5lite.H5Error):
        h5lite.read(p)
    h5lite.write(p, {"x": np.zeros(3, np.uint32)})
    b = bytearray(open(p, "rb").read())
    b[8] = 2                                                                # superblock version 2 (HDF5 1.8 'latest')
    open(p, "wb").write(bytes(b))
    with pytest.raises(h5lite.H5Error):
        h5lite.read(p)
    with pytest.raises(h5lite.H5Error):
        h5lite.write(p, {"x": np.array(["a"])})


def test_reads_an_h5py_written_file():
    """Pins h5lite against libhdf5-written bytes WHEN the artefact exists (tests/golden/external/README.md says how to make it:
    two create_dataset(dtype='uint32') calls, as data/prepro.py:264-277 does).  No h5py / libhdf5 in the build container."""
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "external")
    path = os.path.join(here, "visdial_data_tiny.h5")
    if not os.path.exists(path):
        pytest.skip("external artefact visdial_data_tiny.h5 not present (see tests/golden/external/README.md)")
    d = h5lite.read(path)
    q = d["ques_train"]
    assert q.dtype == np.uint32 and q.shape == (2, 3, 4) and np.array_equal(q.ravel(), np.arange(24))
    ql = d["ques_length_train"]
    assert ql.shape == (2, 3) and (ql == 1).all()
    img = os.path.join(here, "data_img_tiny.h5")
    if os.path.exists(img):
        a = h5lite.read(img)["images_train"]
        assert a.shape == (2, 3, 4, 4) and a.dtype == np.float32 and np.allclose(a, 0.5)
