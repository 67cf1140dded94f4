"""Megatron BERT data path in compat (`fengshen.data.megatron_dataloader.{indexed_dataset,dataset_utils,bert_dataset}`): the mmap
indexed-dataset FILE FORMAT (byte-identical to what the unmodified reference writes; both directions read), split / blend
arithmetic, and BertDataset samples — against tests/golden/megatron_dataset.npz, produced by the reference's Python modules running
on its own compiled C++ helpers (oracle/make_golden_megatron_dataset.py). Plus the parts the reference cannot do: building a
missing index map, and the three-way split through build_train_valid_test_datasets (the reference's imports a module that does
not exist in its tree)."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
for p in ("fengshen-lm_b200", os.path.join("fengshen-lm_b200", "compat")):
    sys.path.insert(0, os.path.join(ROOT, p))

import megatron_dataset_cases as C  # noqa: E402
from make_golden_index_builders import quiet_stdout  # noqa: E402

NAMES = ("utils", "indexed_dataset", "blendable_dataset", "dataset_utils", "bert_dataset")


@pytest.fixture(scope="module")
def M():
    mods = {n: importlib.import_module(f"fengshen.data.megatron_dataloader.{n}") for n in NAMES}
    for m in mods.values():
        assert os.path.join("fengshen-lm_b200", "compat") in m.__file__, m.__file__
    return mods


@pytest.fixture(scope="module")
def H():
    return importlib.import_module("fengshen.data.megatron_dataloader.helpers")


def test_files_arithmetic_and_samples_match_the_reference(M, H, tmp_path):
    golden = np.load(os.path.join(ROOT, "tests", "golden", "megatron_dataset.npz"))
    with quiet_stdout():
        ours = C.run_cases(M, H, str(tmp_path))
    assert set(ours) == set(golden.files)
    for k in golden.files:
        a, b = np.asarray(ours[k]), golden[k]
        assert a.shape == b.shape and a.dtype == b.dtype, (k, a.shape, b.shape, a.dtype, b.dtype)
        assert np.array_equal(a, b), k
    # the fixture is not vacuous
    assert golden["idx_u16"][:9].tobytes() == b"MMIDIDX\x00\x00" and golden["bin_u16"].size * 2 == golden["bin_i32"].size
    assert golden["sop_nsl"].sum() > 0 and (golden["sop_labels"] != -100).any() and (golden["mlm_token_type_ids"] == 0).all()


def test_reference_written_files_are_read_back(M, tmp_path):
    """The golden file holds the exact bytes the reference wrote: dump them and read them with the compat reader."""
    golden = np.load(os.path.join(ROOT, "tests", "golden", "megatron_dataset.npz"))
    prefix = str(tmp_path / "from_reference")
    golden["idx_u16"].tofile(prefix + ".idx")
    golden["bin_u16"].tofile(prefix + ".bin")
    with quiet_stdout():
        ds = M["indexed_dataset"].make_dataset(prefix, "infer", skip_warmup=True)
    assert np.array_equal(ds.sizes, golden["sizes"]) and np.array_equal(ds.doc_idx, golden["doc_idx"])
    assert np.array_equal(ds[7], golden["item_7"]) and ds[7].dtype == np.uint16
    docs = C.corpus_sentences(1, 60)
    flat = [s for d in docs for s in d]
    assert len(ds) == len(flat) and all(np.array_equal(ds[i], flat[i].astype(np.uint16)) for i in range(len(flat)))
    import pickle
    clone = pickle.loads(pickle.dumps(ds))                 # DataLoader workers receive the dataset pickled (path only)
    assert np.array_equal(clone[11], ds[11])
    with pytest.raises(ValueError):
        ds[0:10:2]
    bad = str(tmp_path / "bad")
    open(bad + ".idx", "wb").write(b"TNTIDX\x00\x00" + b"\x00" * 64)
    open(bad + ".bin", "wb").write(b"")
    assert M["indexed_dataset"].infer_dataset_impl(bad) == "cached"
    with pytest.raises(NotImplementedError):
        M["indexed_dataset"].make_dataset(bad, "infer")


def test_missing_index_map_is_built_and_the_three_way_split(M, H, tmp_path):
    docs = C.corpus_sentences(3, 200)
    prefix = str(tmp_path / "corpus")
    C.write_dataset(M, prefix, docs, vocab_size=30000)
    tok = C.ordered_tokenizer()
    with quiet_stdout():
        train, valid, test = M["dataset_utils"].build_train_valid_test_datasets(
            [prefix], "mmap", "8,1,1", [300, 40, 40], 128, 0.15, 0.1, 1234, tok, True, binary_head=True)
    assert len(train) >= 300 and len(valid) >= 40 and len(test) >= 40
    cached = sorted(f for f in os.listdir(tmp_path) if f.endswith(".npy"))
    assert cached == ["corpus_test_indexmap_40mns_125msl_0.10ssp_1234s.npy", "corpus_train_indexmap_300mns_125msl_0.10ssp_1234s.npy",
                      "corpus_valid_indexmap_40mns_125msl_0.10ssp_1234s.npy"]
    with quiet_stdout():
        ds = M["indexed_dataset"].make_dataset(prefix, "mmap", skip_warmup=True)
    bounds = M["dataset_utils"].get_train_valid_test_split_("8,1,1", 200)
    # every split's spans stay inside its own documents; the map equals a direct helpers.build_mapping call on the split's view
    for part, (lo, hi) in zip((train, valid, test), zip(bounds[:-1], bounds[1:])):
        first, last = ds.doc_idx[lo], ds.doc_idx[hi]
        m = np.asarray(part.samples_mapping)
        assert (m[:, 0] >= first).all() and (m[:, 1] <= last).all()
    want = H.build_mapping(np.array(ds.doc_idx[bounds[0]:bounds[1] + 1]), np.array(ds.sizes), np.iinfo(np.int32).max - 1, 300, 125,
                           0.1, 1234, False, 2)
    assert np.array_equal(np.asarray(train.samples_mapping), want)
    s1, s2 = train[5], train[5]
    assert all(np.array_equal(s1[k], s2[k]) for k in ("input_ids", "labels"))       # a sample depends on (seed, index) only
    assert s1["input_ids"].shape == (128,) and s1["attention_mask"].sum() >= 3
    # second construction loads the cached maps instead of rebuilding
    before = {f: os.path.getmtime(tmp_path / f) for f in cached}
    with quiet_stdout():
        M["dataset_utils"].build_train_valid_test_datasets([prefix], "mmap", "8,1,1", [300, 40, 40], 128, 0.15, 0.1, 1234, tok,
                                                           True, binary_head=True)
    assert before == {f: os.path.getmtime(tmp_path / f) for f in cached}
    # blends of two prefixes
    other = str(tmp_path / "other")
    C.write_dataset(M, other, C.corpus_sentences(4, 120), vocab_size=30000)
    with quiet_stdout():
        btrain, bvalid, btest = M["dataset_utils"].build_train_valid_test_datasets(
            ["0.75", prefix, "0.25", other], "mmap", "8,1,1", [200, 20, 20], 128, 0.15, 0.0, 99, tok, True, binary_head=False)
    assert len(btrain) >= 200 and btrain[0]["next_sentence_label"] == 0
    share = np.bincount(btrain.dataset_index[:1000], minlength=2) / min(1000, len(btrain))
    assert abs(share[0] - 0.75) < 0.02
    with pytest.raises(NotImplementedError):
        M["dataset_utils"].build_train_valid_test_datasets([prefix], "mmap", "1", [10, 0, 0], 64, 0.15, 0.0, 1, tok, True,
                                                           dataset_type="bart")
