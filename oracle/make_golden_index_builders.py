"""TEST INFRASTRUCTURE — golden vectors for the index builders, produced by the REFERENCE's own C++
(fengshen/data/megatron_dataloader/helpers.cpp compiled by oracle/build_ref.sh into oracle/_ref/). Run in the authoring
container:  sh oracle/build_ref.sh && python oracle/make_golden_index_builders.py"""
import contextlib
import importlib.util
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def load_reference_helpers():
    """The pybind11 module built from the reference source, or None when oracle/_ref has not been built."""
    hits = glob.glob(os.path.join(HERE, "_ref", "helpers*.so"))
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location("helpers", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@contextlib.contextmanager
def quiet_stdout():
    """helpers.cpp prints progress from C++ (std::cout) regardless of `verbose` in build_sample_idx."""
    sys.stdout.flush()
    saved = os.dup(1)
    with open(os.devnull, "w") as null:
        os.dup2(null.fileno(), 1)
        try:
            yield
        finally:
            os.dup2(saved, 1)
            os.close(saved)


def main():
    sys.path.insert(0, HERE)
    import index_builder_cases as C
    ref = load_reference_helpers()
    if ref is None:
        raise SystemExit("oracle/_ref/helpers*.so missing: run `sh oracle/build_ref.sh` first")
    with quiet_stdout():
        out = C.run_cases(ref)
    path = os.path.join(ROOT, "tests", "golden", "index_builders.npz")
    np.savez_compressed(path, **out)
    print(len(out), "arrays,", os.path.getsize(path), "bytes;", {k: v.shape for k, v in list(out.items())[:4]})


if __name__ == "__main__":
    main()
