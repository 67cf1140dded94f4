"""Host-side timing of the index builders: fsb_index_* (libfsb200.so) next to the reference's own C++ compiled into oracle/_ref/
(test infrastructure; skipped when absent). Single thread both; prints rows/s. Not part of bench.py: the builders run once per
dataset, offline — this is the 'measured beside the reference' line for SURVEY §8f rank 3."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fengshen-lm_b200", "compat"), os.path.join(ROOT, "fengshen-lm_b200")]
import index_builder_cases as C  # noqa: E402
from make_golden_index_builders import load_reference_helpers, quiet_stdout  # noqa: E402
from fengshen.data.megatron_dataloader import helpers as ours  # noqa: E402


def best(f, n=3):
    ts = []
    for _ in range(n):
        t = time.perf_counter(); r = f(); ts.append(time.perf_counter() - t)
    return min(ts), r


def main():
    ref = load_reference_helpers()
    rs = np.random.RandomState(0)
    docs, sizes = C.corpus(rs, 2_000_000)
    titles = rs.randint(0, 20, size=2_000_000).astype(np.int32)
    tok = rs.randint(0, 3000, size=3_000_000).astype(np.int32)
    order = np.concatenate([rs.permutation(tok.shape[0]) for _ in range(2)]).astype(np.int32)
    w = rs.dirichlet(np.ones(8))
    n_blend = 20_000_000
    jobs = {
        "build_mapping (2 M docs, 9 M sentences, 3 epochs)": lambda H: H.build_mapping(docs, sizes, 3, 10 ** 12, 512, 0.1, 1234, False, 2),
        "build_blocks_mapping (same corpus, 3 epochs)": lambda H: H.build_blocks_mapping(docs, sizes, titles, 3, 10 ** 12, 512, 1234, False, False),
        "build_sample_idx (9 G tokens over 2 epochs, seq 2048)": lambda H: H.build_sample_idx(tok, order, 2048, 2, int(tok.sum())),
        "build_blending_indices (8 datasets, 20 M samples)": lambda H: H.build_blending_indices(np.zeros(n_blend, np.uint8), np.zeros(n_blend, np.int64), w, 8, n_blend, False) or np.zeros(n_blend),
    }
    for name, job in jobs.items():
        t_ours, r = best(lambda: job(ours))
        line = f"{name}: fsb200 {t_ours * 1e3:8.1f} ms ({r.shape[0] / t_ours / 1e6:6.1f} M rows/s)"
        if ref is not None:
            with quiet_stdout():
                t_ref, r2 = best(lambda: job(ref))
            line += f" | reference C++ {t_ref * 1e3:8.1f} ms  -> x{t_ref / t_ours:.2f}"
        print(line, flush=True)


if __name__ == "__main__":
    main()
