/* A plain-C99 client of include/fsb200.h: proves the header is C (no C++ types leak through the boundary) and that the shared
 * library is usable without Python or torch. It only calls HOST entry points, so it runs on a machine without a GPU:
 * fsb_version, fsb_last_error, the error path of a device entry point (argument validation happens before any CUDA call), and the
 * index builders (checked against hand-computed values). Built and run by tests/test_abi.py. */
#include <stdio.h>
#include <string.h>

#include "fsb200.h"

static int fail(const char* what) {
  fprintf(stderr, "c_abi_client: %s (last error: %s)\n", what, fsb_last_error());
  return 1;
}

int main(void) {
  if (fsb_version() < 1) return fail("fsb_version");
  if (fsb_last_error() == NULL) return fail("fsb_last_error returned NULL");

  /* blending: weights 0.5 / 0.5 over 6 samples alternate 0,1,0,1,... (ties go to the lowest index) */
  uint8_t which[6];
  int64_t sample[6];
  const double w[2] = {0.5, 0.5};
  if (fsb_index_build_blending_indices(which, sample, w, 2, 6) != FSB_OK) return fail("build_blending_indices");
  const uint8_t want_which[6] = {0, 1, 0, 1, 0, 1};
  const int64_t want_sample[6] = {0, 0, 1, 1, 2, 2};
  if (memcmp(which, want_which, sizeof which) || memcmp(sample, want_sample, sizeof sample)) return fail("blending values");

  /* sample index: documents of 5, 0, 7 tokens, one epoch, sequence length 4 -> (12 - 1) / 4 = 2 samples, 3 rows */
  const int32_t sizes[3] = {5, 0, 7};
  const int32_t order[3] = {0, 1, 2};
  int32_t idx[6];
  if (fsb_index_build_sample_idx(sizes, order, 3, 4, 1, 12, idx, 3) != FSB_OK) return fail("build_sample_idx");
  const int32_t want_idx[6] = {0, 0, 0, 4, 2, 3};   /* positions 0, 4 (doc 0), 8 = 5 + 0 + 3 (doc 2, the empty one is skipped) */
  if (memcmp(idx, want_idx, sizeof idx)) return fail("sample_idx values");
  if (fsb_index_build_sample_idx(sizes, order, 3, 4, 1, 12, idx, 2) != FSB_ERR_INVALID) return fail("wrong row count accepted");
  if (strstr(fsb_last_error(), "rows") == NULL) return fail("error text");

  /* span map: two documents (3 + 2 sentences), no short sequences: count pass, then fill; rows are (first, end, target) */
  const int64_t docs[3] = {0, 3, 5};
  const int32_t sent[5] = {10, 10, 10, 4, 4};
  const int64_t rows = fsb_index_build_mapping(docs, 2, sent, 1, 1000, 15, 0.0, 7, 2, FSB_U32, NULL, 0);
  if (rows != 2) return fail("build_mapping row count");   /* doc 0 cannot close early (only one sentence would remain) */
  uint32_t map[6];
  if (fsb_index_build_mapping(docs, 2, sent, 1, 1000, 15, 0.0, 7, 2, FSB_U32, map, rows) != rows) return fail("build_mapping fill");
  int seen_a = 0, seen_b = 0, i;
  for (i = 0; i < 2; ++i) {
    if (map[3 * i] == 0 && map[3 * i + 1] == 3 && map[3 * i + 2] == 15) seen_a = 1;
    if (map[3 * i] == 3 && map[3 * i + 1] == 5 && map[3 * i + 2] == 15) seen_b = 1;
  }
  if (!(seen_a && seen_b)) return fail("build_mapping values");

  /* a device entry point with a bad argument fails cleanly (validation precedes any CUDA call), never crashes */
  if (fsb_rmsnorm_fwd(NULL, NULL, NULL, NULL, NULL, NULL, 4, 6, 1e-6f, NULL) == FSB_OK) return fail("rmsnorm accepted NULL");
  printf("c_abi_client ok (fsb_version %d)\n", fsb_version());
  return 0;
}
