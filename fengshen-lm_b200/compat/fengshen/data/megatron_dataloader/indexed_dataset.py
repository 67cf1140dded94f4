"""Memory-mapped indexed datasets — the on-disk format of fengshen/data/megatron_dataloader/indexed_dataset.py:344-585
(Megatron-LM / fairseq `mmap` implementation), reader and writer. Files written here are byte-identical to the reference's and
either side reads the other's (tests/test_megatron_dataset_cpu.py).

  <prefix>.bin   the token arrays of all sentences, back to back, in the index's dtype
  <prefix>.idx   b'MMIDIDX\\x00\\x00' | <Q version = 1 | <B dtype code | <Q n_sentences | <Q n_doc_entries
                 | int32 sizes[n_sentences] | int64 byte pointers[n_sentences] | int64 doc_idx[n_doc_entries]
  doc_idx[d] is the first sentence of document d, with a closing entry (= n_sentences once every document was ended).

Only the `mmap` implementation is provided ('infer' recognises it); the legacy 'lazy' / 'cached' TNTIDX format is refused loudly."""
import os
import shutil
import struct
from itertools import accumulate

import numpy as np
import torch

from fengshen.data.megatron_dataloader.utils import print_rank_0

_MMAP_MAGIC = b'MMIDIDX\x00\x00'
_LEGACY_MAGIC = b'TNTIDX\x00\x00'

# dtype codes of the format (indexed_dataset.py:95-104; code 6 is the builtin float == float64 there)
dtypes = {1: np.uint8, 2: np.int8, 3: np.int16, 4: np.int32, 5: np.int64, 6: np.float64, 7: np.double, 8: np.uint16}


def code(dtype):
    for k, v in dtypes.items():
        if v == dtype:
            return k
    raise ValueError(dtype)


def best_fitting_dtype(vocab_size=None):
    """indexed_dataset.py:24-28: uint16 when every token id fits."""
    return np.uint16 if vocab_size is not None and vocab_size < 65500 else np.int32


def index_file_path(prefix_path):
    return prefix_path + '.idx'


def data_file_path(prefix_path):
    return prefix_path + '.bin'


def get_available_dataset_impl():
    return ['mmap']


def _exists(path):
    return os.path.exists(index_file_path(path)) and os.path.exists(data_file_path(path))


def infer_dataset_impl(path):
    if not _exists(path):
        print(f"Dataset does not exist: {path}")
        print("Path should be a basename that both .idx and .bin can be appended to get full filenames.")
        return None
    with open(index_file_path(path), 'rb') as f:
        magic = f.read(8)
    if magic == _MMAP_MAGIC[:8]:
        return 'mmap'
    if magic == _LEGACY_MAGIC:
        return 'cached'
    return None


def dataset_exists(path, impl):
    return _exists(path)


def make_builder(out_file, impl, vocab_size=None):
    if impl != 'mmap':
        raise NotImplementedError(f"fsb200 compat: dataset implementation {impl!r} is not provided (mmap is)")
    return MMapIndexedDatasetBuilder(out_file, dtype=best_fitting_dtype(vocab_size))


def make_dataset(path, impl, skip_warmup=False):
    if not _exists(path):
        print(f"Dataset does not exist: {path}")
        print("Path should be a basename that both .idx and .bin can be appended to get full filenames.")
        return None
    if impl == 'infer':
        impl = infer_dataset_impl(path)
    if impl == 'mmap':
        return MMapIndexedDataset(path, skip_warmup)
    if impl in ('lazy', 'cached'):
        raise NotImplementedError("fsb200 compat: the legacy TNTIDX indexed-dataset format is not provided; re-encode with mmap")
    print(f"Unknown dataset implementation: {impl}")
    return None


def _warmup_mmap_file(path):
    with open(path, 'rb') as stream:
        while stream.read(100 * 1024 * 1024):
            pass


class _Index:
    """The .idx file, mapped. sizes / pointers / doc_idx are zero-copy views."""
    _HDR_MAGIC = _MMAP_MAGIC
    _HEADER = struct.Struct('<9sQBQQ')

    @classmethod
    def write(cls, path, dtype, sizes, doc_idx):
        sizes = np.asarray(sizes, dtype=np.int32)
        pointers = np.zeros(sizes.shape[0], dtype=np.int64)
        if sizes.shape[0] > 1:
            np.cumsum(sizes[:-1].astype(np.int64) * np.dtype(dtype).itemsize, out=pointers[1:])
        doc_idx = np.asarray(doc_idx, dtype=np.int64)
        with open(path, 'wb') as f:
            f.write(cls._HEADER.pack(cls._HDR_MAGIC, 1, code(dtype), sizes.shape[0], doc_idx.shape[0]))
            f.write(sizes.tobytes(order='C'))
            f.write(pointers.tobytes(order='C'))
            f.write(doc_idx.tobytes(order='C'))

    def __init__(self, path, skip_warmup=False):
        with open(path, 'rb') as f:
            raw = f.read(self._HEADER.size)
        if len(raw) < self._HEADER.size or raw[:9] != self._HDR_MAGIC:
            raise ValueError(f"{path}: not an mmap indexed-dataset index (make sure --dataset-impl is configured properly)")
        _, version, dtype_code, self._len, self._doc_count = self._HEADER.unpack(raw)
        if version != 1:
            raise ValueError(f"{path}: unsupported index version {version}")
        self._dtype = dtypes[dtype_code]
        if not skip_warmup:
            print_rank_0("    warming up index mmap file...")
            _warmup_mmap_file(path)
        self._mmap = np.memmap(path, mode='r', order='C')
        buf = memoryview(self._mmap)
        off = self._HEADER.size
        self._sizes = np.frombuffer(buf, dtype=np.int32, count=self._len, offset=off)
        self._pointers = np.frombuffer(buf, dtype=np.int64, count=self._len, offset=off + self._sizes.nbytes)
        self._doc_idx = np.frombuffer(buf, dtype=np.int64, count=self._doc_count,
                                      offset=off + self._sizes.nbytes + self._pointers.nbytes)

    dtype = property(lambda self: self._dtype)
    sizes = property(lambda self: self._sizes)
    doc_idx = property(lambda self: self._doc_idx)

    def __getitem__(self, i):
        return self._pointers[i], self._sizes[i]

    def __len__(self):
        return self._len


class MMapIndexedDataset(torch.utils.data.Dataset):
    Index = _Index

    def __init__(self, path, skip_warmup=False):
        super().__init__()
        self._do_init(path, skip_warmup)

    def __getstate__(self):
        return self._path

    def __setstate__(self, state):
        self._do_init(state, skip_warmup=True)

    def _do_init(self, path, skip_warmup):
        self._path = path
        self._index = _Index(index_file_path(path), skip_warmup)
        if not skip_warmup:
            print_rank_0("    warming up data mmap file...")
            _warmup_mmap_file(data_file_path(path))
        self._bin_mmap = np.memmap(data_file_path(path), mode='r', order='C')
        self._bin = memoryview(self._bin_mmap)

    def __len__(self):
        return len(self._index)

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            ptr, size = self._index[idx]
            return np.frombuffer(self._bin, dtype=self._index.dtype, count=size, offset=ptr)
        if isinstance(idx, slice):
            start, stop, step = idx.indices(len(self))
            if step != 1:
                raise ValueError("Slices into indexed_dataset must be contiguous")
            sizes = self._index._sizes[idx]
            flat = np.frombuffer(self._bin, dtype=self._index.dtype, count=int(sizes.sum()),
                                 offset=self._index._pointers[start])
            return np.split(flat, list(accumulate(sizes))[:-1])
        raise TypeError(f"MMapIndexedDataset indices must be integers or slices, not {type(idx).__name__}")

    def get(self, idx, offset=0, length=None):
        """Part of one sentence: `length` tokens from token `offset` (indexed_dataset.py:514-526)."""
        ptr, size = self._index[idx]
        if length is None:
            length = size - offset
        return np.frombuffer(self._bin, dtype=self._index.dtype, count=length,
                             offset=ptr + offset * np.dtype(self._index.dtype).itemsize)

    sizes = property(lambda self: self._index.sizes)
    doc_idx = property(lambda self: self._index.doc_idx)

    def get_doc_idx(self):
        return self._index._doc_idx

    def set_doc_idx(self, doc_idx_):
        self._index._doc_idx = doc_idx_

    @property
    def supports_prefetch(self):
        return False

    @staticmethod
    def exists(path):
        return _exists(path)


class MMapIndexedDatasetBuilder(object):
    def __init__(self, out_file, dtype=np.int64):
        self._data_file = open(out_file, 'wb', buffering=5000000)
        self._dtype = dtype
        self._sizes = []
        self._doc_idx = [0]

    def add_item(self, tensor):
        arr = np.array(tensor.numpy() if hasattr(tensor, "numpy") else tensor, dtype=self._dtype)
        self._data_file.write(arr.tobytes(order='C'))
        self._sizes.append(arr.size)

    def end_document(self):
        self._doc_idx.append(len(self._sizes))

    def merge_file_(self, another_file, merge_doc_idx=False):
        """Append another dataset's sentences. As in the reference (indexed_dataset.py:569-579) the other file's document
        boundaries are NOT carried over unless merge_doc_idx=True (then they are, shifted — upstream Megatron's behaviour)."""
        index = _Index(index_file_path(another_file), skip_warmup=True)
        if index.dtype != self._dtype:
            raise ValueError("merge_file_: dtype mismatch")
        base = len(self._sizes)
        self._sizes.extend(int(s) for s in index.sizes)
        if merge_doc_idx:
            self._doc_idx.extend(int(d) + base for d in index.doc_idx[1:])
        with open(data_file_path(another_file), 'rb') as f:
            shutil.copyfileobj(f, self._data_file)

    def finalize(self, index_file):
        self._data_file.close()
        _Index.write(index_file, self._dtype, self._sizes, self._doc_idx)
