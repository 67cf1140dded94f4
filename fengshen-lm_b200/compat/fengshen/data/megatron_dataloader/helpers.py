"""`fengshen.data.megatron_dataloader.helpers` — in the reference a pybind11 extension built from helpers.cpp by `make`
(fengshen/data/megatron_dataloader/Makefile:1-9, entry points helpers.cpp:788-793); here the same four functions over
`fsb_index_*` of libfsb200.so (include/fsb200.h). Same argument order, numpy arrays in, numpy arrays out (or filled in place for
build_blending_indices), bit-identical integers including the reference's shuffles. `verbose` only controls the reference's
progress printing and is accepted for signature parity."""
import ctypes

import numpy as np

from fsb200 import lib as _L

_UINT32_MAX = 2 ** 32 - 1


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def _arr(a, dtype, name):
    out = np.ascontiguousarray(a, dtype=dtype)
    if out.ndim != 1:
        raise ValueError(f"helpers: {name} must be one-dimensional")
    return out


def _fail(what):
    raise RuntimeError(f"fsb200: {what} failed: {_L.last_error()}")


def build_sample_idx(sizes, doc_idx, seq_length, num_epochs, tokens_per_epoch):
    """helpers.cpp:101-195 -> int32 [num_samples + 1, 2]: (index into doc_idx, token offset inside that document)."""
    sizes, doc_idx = _arr(sizes, np.int32, "sizes"), _arr(doc_idx, np.int32, "doc_idx")
    if not (seq_length > 1 and num_epochs > 0 and tokens_per_epoch > 1):
        raise ValueError("helpers.build_sample_idx: seq_length > 1, num_epochs > 0, tokens_per_epoch > 1 required")
    num_samples = (int(num_epochs) * int(tokens_per_epoch) - 1) // int(seq_length)
    out = np.empty((num_samples + 1, 2), dtype=np.int32)
    _L.call("fsb_index_build_sample_idx", _ptr(sizes), _ptr(doc_idx), doc_idx.shape[0], int(seq_length), int(num_epochs),
            int(tokens_per_epoch), _ptr(out), out.shape[0])
    return out


def _two_pass(name, cols, n_sentences, args_before_dtype):
    dtype_np, dtype_c = (np.uint64, _L.U64) if n_sentences > _UINT32_MAX else (np.uint32, _L.U32)
    fn = getattr(_L.load(), name)
    rows = fn(*args_before_dtype, dtype_c, None, 0)
    if rows < 0:
        _fail(name)
    out = np.empty((rows, cols), dtype=dtype_np)
    if fn(*args_before_dtype, dtype_c, _ptr(out), rows) != rows:
        _fail(name)
    return out


def build_mapping(docs, sizes, num_epochs, max_num_samples, max_seq_length, short_seq_prob, seed, verbose, min_num_sent):
    """helpers.cpp:461-499 -> uint32 / uint64 [num_samples, 3]: (first sentence, end sentence, target sequence length)."""
    docs, sizes = _arr(docs, np.int64, "docs"), _arr(sizes, np.int32, "sizes")
    return _two_pass("fsb_index_build_mapping", 3, sizes.shape[0],
                     (_ptr(docs), docs.shape[0] - 1, _ptr(sizes), int(num_epochs), int(max_num_samples), int(max_seq_length),
                      float(short_seq_prob), int(seed), int(min_num_sent)))


def build_blocks_mapping(docs, sizes, titles_sizes, num_epochs, max_num_samples, max_seq_length, seed, verbose,
                         use_one_sent_blocks):
    """helpers.cpp:752-786 -> uint32 / uint64 [num_samples, 4]: (first sentence, end sentence, document, block id)."""
    docs, sizes = _arr(docs, np.int64, "docs"), _arr(sizes, np.int32, "sizes")
    titles_sizes = _arr(titles_sizes, np.int32, "titles_sizes")
    if titles_sizes.shape[0] < docs.shape[0] - 1:
        raise ValueError("helpers.build_blocks_mapping: titles_sizes needs one entry per document")
    return _two_pass("fsb_index_build_blocks_mapping", 4, sizes.shape[0],
                     (_ptr(docs), docs.shape[0] - 1, _ptr(sizes), _ptr(titles_sizes), int(num_epochs), int(max_num_samples),
                      int(max_seq_length), int(seed), int(bool(use_one_sent_blocks))))


def build_blending_indices(dataset_index, dataset_sample_index, weights, num_datasets, size, verbose):
    """helpers.cpp:34-99: fills dataset_index (uint8 [size]) and dataset_sample_index (int64 [size]) IN PLACE."""
    for a, dt, n in ((dataset_index, np.uint8, "dataset_index"), (dataset_sample_index, np.int64, "dataset_sample_index")):
        if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags["C_CONTIGUOUS"] and a.ndim == 1 and a.shape[0] >= size):
            raise TypeError(f"helpers.build_blending_indices: {n} must be a contiguous {np.dtype(dt).name} array of length >= size")
    w = _arr(weights, np.float64, "weights")
    if w.shape[0] < num_datasets:
        raise ValueError("helpers.build_blending_indices: fewer weights than datasets")
    _L.call("fsb_index_build_blending_indices", _ptr(dataset_index), _ptr(dataset_sample_index), _ptr(w), int(num_datasets),
            int(size))
