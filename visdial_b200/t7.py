"""Torch7 binary serialisation (`torch.save` / `torch.load`, the `.t7` container) — the subset the reference's
checkpoints use: nil, numbers, booleans, strings, tables (with shared references) and torch tensors / storages.

The reference writes its checkpoints with `torch.save(path, {modelW = ..., optims = ..., modelParams = ...})`
(/root/reference/train.lua:99-102,120-121) and reads them back with `torch.load` (train.lua:33-34, evaluate.lua:58,
generate.lua:53).  torch7 itself is NOT in /root/reference (README.md:34-44 installs it from torch/distro); this file
restates its published on-disk format [upstream torch7 File.lua `writeObject/readObject`, generic/Tensor.c
`torch_Tensor_(write)`, generic/Storage.c `torch_Storage_(write)`], binary mode, native little-endian:

    object   := int32 type, payload
    type     := 0 nil | 1 number | 2 string | 3 table | 4 torch object | 5 boolean
    number   := float64                      boolean := int32 (1/0)        string := int32 n, n bytes
    table    := int32 index; if index is new: int32 npairs, npairs x (object key, object value)
    torch    := int32 index; if index is new: string "V <version>", string class name, class payload
    Tensor   := int32 ndim, int64 size[ndim], int64 stride[ndim], int64 storageOffset (1-based), object storage
    Storage  := int64 n, n raw elements

PARITY UNPINNED: no Lua/torch7 runs in this image and the reference holds no `.t7` fixture, so the codec is checked
against byte strings assembled by hand from the format above (tests/test_t7.py) and by round trips, not against files
written by torch7.
"""
from __future__ import annotations

import struct
from typing import Any, BinaryIO, Dict

import numpy as np

TYPE_NIL, TYPE_NUMBER, TYPE_STRING, TYPE_TABLE, TYPE_TORCH, TYPE_BOOLEAN = 0, 1, 2, 3, 4, 5

_STORAGE_DTYPES = {
    "torch.FloatStorage": np.float32, "torch.DoubleStorage": np.float64, "torch.LongStorage": np.int64,
    "torch.IntStorage": np.int32, "torch.ShortStorage": np.int16, "torch.ByteStorage": np.uint8,
    "torch.CharStorage": np.int8, "torch.CudaStorage": np.float32, "torch.CudaDoubleStorage": np.float64,
    "torch.CudaLongStorage": np.int64, "torch.CudaIntStorage": np.int32, "torch.HalfStorage": np.float16,
    "torch.CudaHalfStorage": np.float16,
}
_TENSOR_TO_STORAGE = {k.replace("Storage", "Tensor"): k for k in _STORAGE_DTYPES}
_DTYPE_TO_TENSOR = {np.dtype(np.float32): "torch.FloatTensor", np.dtype(np.float64): "torch.DoubleTensor",
                    np.dtype(np.int64): "torch.LongTensor", np.dtype(np.int32): "torch.IntTensor",
                    np.dtype(np.int16): "torch.ShortTensor", np.dtype(np.uint8): "torch.ByteTensor",
                    np.dtype(np.int8): "torch.CharTensor", np.dtype(np.float16): "torch.HalfTensor"}


class T7Error(ValueError):
    pass


class TorchObject:
    """A torch class this codec does not interpret (kept so that a table still loads)."""

    def __init__(self, class_name: str, version: int):
        self.class_name, self.version = class_name, version

    def __repr__(self):
        return "TorchObject(%s)" % self.class_name


# ---------------------------------------------------------------------------------------------------- reader
class _Reader:
    def __init__(self, f: BinaryIO):
        self.f = f
        self.memo: Dict[int, Any] = {}

    def _read(self, n: int) -> bytes:
        b = self.f.read(n)
        if len(b) != n:
            raise T7Error("unexpected end of file")
        return b

    def int(self) -> int:
        return struct.unpack("<i", self._read(4))[0]

    def long(self) -> int:
        return struct.unpack("<q", self._read(8))[0]

    def string(self) -> str:
        n = self.int()
        if n < 0:
            raise T7Error("negative string length")
        return self._read(n).decode("latin-1")

    def obj(self) -> Any:
        t = self.int()
        if t == TYPE_NIL:
            return None
        if t == TYPE_NUMBER:
            v = struct.unpack("<d", self._read(8))[0]
            return int(v) if v.is_integer() and abs(v) < 2 ** 53 else v
        if t == TYPE_BOOLEAN:
            return self.int() == 1
        if t == TYPE_STRING:
            return self.string()
        if t == TYPE_TABLE:
            idx = self.int()
            if idx in self.memo:
                return self.memo[idx]
            out: Dict[Any, Any] = {}
            self.memo[idx] = out
            for _ in range(self.int()):
                k = self.obj()
                out[k] = self.obj()
            return out
        if t == TYPE_TORCH:
            idx = self.int()
            if idx in self.memo:
                return self.memo[idx]
            s = self.string()
            if s.startswith("V "):
                version, cls = int(s[2:]), self.string()
            else:                                   # files older than the versioned header carry the class name directly
                version, cls = 0, s
            val = self._torch(cls, version, idx)
            self.memo[idx] = val
            return val
        raise T7Error("unsupported object type %d (functions / userdata are not part of the checkpoint schema)" % t)

    def _torch(self, cls: str, version: int, idx: int) -> Any:
        if cls in _STORAGE_DTYPES:
            n = self.long()
            dt = np.dtype(_STORAGE_DTYPES[cls]).newbyteorder("<")
            return np.frombuffer(self._read(n * dt.itemsize), dtype=dt).copy()
        if cls in _TENSOR_TO_STORAGE:
            nd = self.int()
            size = [self.long() for _ in range(nd)]
            stride = [self.long() for _ in range(nd)]
            off = self.long() - 1                   # stored 1-based "to respect Lua convention"
            storage = self.obj()
            if storage is None or nd == 0:
                return np.zeros(size if nd else (0,), dtype=_STORAGE_DTYPES[_TENSOR_TO_STORAGE[cls]])
            it = storage.dtype.itemsize
            view = np.lib.stride_tricks.as_strided(storage[off:], shape=size, strides=[s * it for s in stride])
            return np.array(view)                   # owns its data, C-contiguous
        self.memo[idx] = obj = TorchObject(cls, version)
        raise T7Error("torch class %r is not part of the checkpoint schema (%r)" % (cls, obj))


def load(path_or_file) -> Any:
    """torch.load(path): tensors come back as numpy arrays, Lua tables as dicts (see `as_list` for array-like ones)."""
    if hasattr(path_or_file, "read"):
        return _Reader(path_or_file).obj()
    with open(path_or_file, "rb") as f:
        return _Reader(f).obj()


def as_list(table: Dict[Any, Any]) -> list:
    """A Lua array-like table {1: a, 2: b, ...} -> [a, b, ...]."""
    n = len(table)
    if sorted(table.keys()) != list(range(1, n + 1)):
        raise T7Error("table is not an array")
    return [table[i] for i in range(1, n + 1)]


# ---------------------------------------------------------------------------------------------------- writer
class CudaTensor:
    """Marks an array that must be written as torch.CudaTensor (train.lua:99-102 saves wrapperW as is, i.e. CUDA)."""

    def __init__(self, array: np.ndarray):
        self.array = np.ascontiguousarray(array, dtype=np.float32)


class _Writer:
    def __init__(self, f: BinaryIO):
        self.f = f
        self.seen: Dict[int, int] = {}
        self.keep = []                               # keeps id()s unique for the duration of the write
        self.n = 0

    def int(self, v: int):
        self.f.write(struct.pack("<i", v))

    def long(self, v: int):
        self.f.write(struct.pack("<q", v))

    def string(self, s: str):
        b = s.encode("latin-1")
        self.int(len(b))
        self.f.write(b)

    def _index(self, o: Any) -> bool:
        """writes the object index; True if the object still has to be written"""
        key = id(o)
        if key in self.seen:
            self.int(self.seen[key])
            return False
        self.n += 1
        self.seen[key] = self.n
        self.keep.append(o)
        self.int(self.n)
        return True

    def _storage(self, cls: str, flat: np.ndarray):
        self.int(TYPE_TORCH)
        if self._index(flat):
            self.string("V 1")
            self.string(cls)
            self.long(flat.size)
            self.f.write(np.ascontiguousarray(flat).astype(flat.dtype.newbyteorder("<"), copy=False).tobytes())

    def _tensor(self, o: np.ndarray, cls: str):
        self.int(TYPE_TORCH)
        if not self._index(o):
            return
        a = np.ascontiguousarray(o)
        self.string("V 1")
        self.string(cls)
        self.int(a.ndim)
        for s in a.shape:
            self.long(int(s))
        for s in a.strides:
            self.long(int(s // a.itemsize))
        self.long(1)
        if a.size == 0:
            self.int(TYPE_NIL)
        else:
            self._storage(_TENSOR_TO_STORAGE[cls], a.reshape(-1))

    def obj(self, o: Any):
        if o is None:
            self.int(TYPE_NIL)
        elif isinstance(o, (bool, np.bool_)):
            self.int(TYPE_BOOLEAN)
            self.int(1 if o else 0)
        elif isinstance(o, (int, float, np.integer, np.floating)):
            self.int(TYPE_NUMBER)
            self.f.write(struct.pack("<d", float(o)))
        elif isinstance(o, str):
            self.int(TYPE_STRING)
            self.string(o)
        elif isinstance(o, CudaTensor):
            self._tensor(o.array, "torch.CudaTensor")
        elif isinstance(o, np.ndarray):
            dt = np.dtype(o.dtype.type)
            if dt not in _DTYPE_TO_TENSOR:
                raise T7Error("no torch tensor type for dtype %s" % o.dtype)
            self._tensor(o, _DTYPE_TO_TENSOR[dt])
        elif isinstance(o, (list, tuple)):
            self.obj({i + 1: v for i, v in enumerate(o)})
        elif isinstance(o, dict):
            self.int(TYPE_TABLE)
            if self._index(o):
                self.int(len(o))
                for k, v in o.items():
                    self.obj(k)
                    self.obj(v)
        else:
            raise T7Error("cannot serialise %r" % type(o))


def save(path_or_file, obj: Any):
    """torch.save(path, obj) for dicts / lists / numbers / strings / booleans / numpy arrays."""
    if hasattr(path_or_file, "write"):
        _Writer(path_or_file).obj(obj)
        return
    with open(path_or_file, "wb") as f:
        _Writer(f).obj(obj)
