"""TEST INFRASTRUCTURE — golden vectors for the Megatron BERT data path from the UNMODIFIED reference: its Python modules
fengshen/data/megatron_dataloader/{utils,indexed_dataset,blendable_dataset,dataset_utils,bert_dataset}.py loaded by path and
registered under their real names, with `helpers` = its own C++ compiled into oracle/_ref/ (oracle/build_ref.sh). Two shims, both
environment repairs rather than code changes: `np.float` (removed from numpy 1.24+, named in indexed_dataset.py:101) is restored
for the import, and torch.distributed is initialised single-process because blendable_dataset.py calls get_rank() unconditionally.
Run in the authoring container:  sh oracle/build_ref.sh && python oracle/make_golden_megatron_dataset.py"""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
NAMES = ("utils", "indexed_dataset", "blendable_dataset", "dataset_utils", "bert_dataset")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    sys.path.insert(0, HERE)
    from make_golden_index_builders import load_reference_helpers, quiet_stdout
    import megatron_dataset_cases as C
    helpers = load_reference_helpers()
    if helpers is None:
        raise SystemExit("oracle/_ref/helpers*.so missing: run `sh oracle/build_ref.sh` first")
    if not hasattr(np, "float"):
        np.float = float
    for pkg in ("fengshen", "fengshen.data", "fengshen.data.megatron_dataloader"):
        m = types.ModuleType(pkg)
        m.__path__ = []
        sys.modules[pkg] = m
    sys.modules["fengshen.data.megatron_dataloader.helpers"] = helpers
    sys.modules["fengshen.data.megatron_dataloader"].helpers = helpers
    M = {n: _load(f"fengshen.data.megatron_dataloader.{n}", os.path.join(REF, "fengshen/data/megatron_dataloader", n + ".py"))
         for n in NAMES}
    for m in M.values():
        assert m.__file__.startswith(REF)
    with tempfile.TemporaryDirectory() as tmp, quiet_stdout():
        out = C.run_cases(M, helpers, tmp)
    path = os.path.join(ROOT, "tests", "golden", "megatron_dataset.npz")
    np.savez_compressed(path, **out)
    print(len(out), "arrays,", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
