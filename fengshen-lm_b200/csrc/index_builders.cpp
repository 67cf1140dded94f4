// Host-side index builders of the Megatron indexed datasets — the C ABI behind
// fengshen/data/megatron_dataloader/helpers.cpp (pybind11 module `helpers`, :788-793): sample index of the flattened GPT-style
// stream, sentence-span sample maps for BERT-style pretraining (with and without block ids), and the dataset blending schedule.
// Integer work, bit-exact with the reference (which is compiled from its own source into oracle/_ref/ as the test oracle).
//
// Design notes (not a transcription):
//  * every output buffer is caller-owned; the variable-length builders are called twice, first with out == NULL to learn the
//    number of rows, then to fill — both passes re-seed the 32-bit Mersenne twister, so they see the same draws;
//  * the span builders share one walker (`walk_spans`) parameterised by how a target length is chosen and what a row holds;
//  * build_sample_idx works on token POSITIONS: sample k starts at flattened position k*seq_length, found by advancing one cursor
//    over the running document ends (O(samples + documents), 64-bit positions, zero-length documents skipped like the reference).
// The pseudo-random sequence is part of the format (a cached index built by the reference must equal one built here):
// std::mt19937(seed) drives the short-sequence draws, std::mt19937_64(seed + 1) the Fisher-Yates shuffle — both generators are
// fully specified by the C++ standard.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <random>

#include "../../include/fsb200.h"
#include "host_common.h"

namespace {

constexpr int32_t kLongSentence = 512;   // helpers.cpp:32 — documents holding a longer sentence are skipped entirely

template <typename T>
inline void put(void* out, int dtype, int64_t i, T v) {
  if (dtype == FSB_U64) static_cast<uint64_t*>(out)[i] = static_cast<uint64_t>(v);
  else static_cast<uint32_t*>(out)[i] = static_cast<uint32_t>(v);
}

struct TargetDraw {   // helpers.cpp:197-212: one 32-bit draw decides BOTH whether the sample is short and how short
  int32_t ratio, max_len;
  std::mt19937 gen;
  TargetDraw(double short_prob, int32_t max_len_, int32_t seed)
      : ratio(short_prob > 0 ? static_cast<int32_t>(std::round(1.0 / short_prob)) : 0), max_len(max_len_), gen(seed) {}
  int32_t next() {
    if (ratio == 0) return max_len;
    const auto r = gen();
    return (r % ratio) == 0 ? static_cast<int32_t>(2 + r % (max_len - 1)) : max_len;
  }
};

bool has_long_sentence(const int32_t* sizes, int64_t first, int64_t last) {
  for (int64_t s = first; s < last; ++s)
    if (sizes[s] > kLongSentence) return true;
  return false;
}

// Walks epochs x documents x sentences and calls emit(row, first_sentence, end_sentence, doc, block_id, target) for every span.
// `target_of(doc)` gives the first target of a document, `next_target()` the one after each emitted span. `remain_min` is the
// number of sentences that must be left over for a span to close early (2 for build_mapping: "> 1"; min_num_sent for blocks).
template <typename FirstTarget, typename NextTarget, typename Emit>
uint64_t walk_spans(const int64_t* docs, int64_t n_docs, const int32_t* sizes, int32_t num_epochs, uint64_t max_rows,
                    int32_t min_num_sent, int32_t remain_min, bool long_check_needs_min, FirstTarget target_of,
                    NextTarget next_target, Emit emit) {
  uint64_t row = 0;
  for (int32_t epoch = 0; epoch < num_epochs && row < max_rows; ++epoch) {
    int32_t block_id = 0;
    for (int64_t doc = 0; doc < n_docs; ++doc) {
      const int64_t first = docs[doc], last = docs[doc + 1];
      int64_t remain = last - first;
      if (remain < min_num_sent) continue;
      // build_mapping inspects sentence lengths only when the document has more than one sentence (helpers.cpp:338),
      // build_blocks_mapping whenever it has min_num_sent (:633) — the same thing once `remain >= min_num_sent` except for
      // one-sentence documents under min_num_sent == 1
      if ((!long_check_needs_min ? remain > 1 : true) && has_long_sentence(sizes, first, last)) continue;
      int64_t start = first;
      int32_t len = 0, count = 0;
      int32_t target = target_of(doc);
      for (int64_t s = first; s < last; ++s) {
        len += sizes[s];
        ++count;
        --remain;
        if ((len >= target && remain >= remain_min && count >= min_num_sent) || remain == 0) {
          emit(row, start, s + 1, doc, block_id, target);
          ++row;
          ++block_id;
          start = s + 1;
          target = next_target(doc, target);
          len = 0;
          count = 0;
        }
      }
    }
  }
  return row;
}

void shuffle_rows(void* out, int dtype, int64_t rows, int cols, int32_t seed) {   // helpers.cpp:447-459 / :738-750
  std::mt19937_64 gen(static_cast<uint64_t>(seed + 1));
  for (int64_t i = rows - 1; i > 0; --i) {
    const int64_t j = static_cast<int64_t>(gen() % static_cast<uint64_t>(i + 1));
    if (dtype == FSB_U64) {
      auto* p = static_cast<uint64_t*>(out);
      for (int c = 0; c < cols; ++c) std::swap(p[cols * i + c], p[cols * j + c]);
    } else {
      auto* p = static_cast<uint32_t*>(out);
      for (int c = 0; c < cols; ++c) std::swap(p[cols * i + c], p[cols * j + c]);
    }
  }
}

}  // namespace

extern "C" {

int fsb_index_build_sample_idx(const int32_t* sizes, const int32_t* doc_idx, int64_t n_doc_idx, int32_t seq_length,
                               int32_t num_epochs, int64_t tokens_per_epoch, int32_t* out, int64_t out_rows) {
  FSB_REQUIRE(sizes && doc_idx && out, "build_sample_idx: null pointer");
  FSB_REQUIRE(seq_length > 1 && num_epochs > 0 && tokens_per_epoch > 1, "build_sample_idx: seq_length > 1, num_epochs > 0, "
              "tokens_per_epoch > 1 required (got %d, %d, %ld)", seq_length, num_epochs, (long)tokens_per_epoch);
  const int64_t num_samples = (static_cast<int64_t>(num_epochs) * tokens_per_epoch - 1) / seq_length;
  FSB_REQUIRE(out_rows == num_samples + 1, "build_sample_idx: out must hold %ld rows (got %ld)", (long)(num_samples + 1),
              (long)out_rows);
  out[0] = 0;
  out[1] = 0;
  int64_t d = 0;          // cursor into doc_idx
  int64_t doc_begin = 0;  // flattened position of the first token of doc_idx[d]
  for (int64_t k = 1; k <= num_samples; ++k) {
    const int64_t pos = k * static_cast<int64_t>(seq_length);   // sample k reads tokens [pos, pos + seq_length]
    for (;;) {
      FSB_REQUIRE(d < n_doc_idx, "build_sample_idx: doc_idx exhausted at sample %ld (tokens_per_epoch inconsistent with sizes)",
                  (long)k);
      const int64_t len = sizes[doc_idx[d]];
      if (pos < doc_begin + len) break;
      doc_begin += len;
      ++d;
    }
    out[2 * k] = static_cast<int32_t>(d);
    out[2 * k + 1] = static_cast<int32_t>(pos - doc_begin);
  }
  return 0;
}

int64_t fsb_index_build_mapping(const int64_t* docs, int64_t n_docs, const int32_t* sizes, int32_t num_epochs,
                                uint64_t max_num_samples, int32_t max_seq_length, double short_seq_prob, int32_t seed,
                                int32_t min_num_sent, int dtype, void* out, int64_t out_rows) {
  if (!(docs && sizes) || n_docs < 0 || num_epochs <= 0 || max_seq_length <= 1 || short_seq_prob < 0.0 ||
      short_seq_prob > 1.0 || seed <= 0 || (dtype != FSB_U32 && dtype != FSB_U64)) {
    fsb::set_error("build_mapping: need docs/sizes, num_epochs > 0, max_seq_length > 1, 0 <= short_seq_prob <= 1, seed > 0, "
                   "dtype FSB_U32 | FSB_U64");
    return -1;
  }
  TargetDraw draw(short_seq_prob, max_seq_length, seed);
  auto first_target = [&](int64_t) { return draw.next(); };
  auto next_target = [&](int64_t, int32_t) { return draw.next(); };
  int64_t bad = 0;
  auto emit = [&](uint64_t row, int64_t a, int64_t b, int64_t, int32_t, int32_t target) {
    if (!out) return;
    if (static_cast<int64_t>(row) >= out_rows) { ++bad; return; }
    put(out, dtype, 3 * row, a);
    put(out, dtype, 3 * row + 1, b);
    put(out, dtype, 3 * row + 2, target);
  };
  const uint64_t rows = walk_spans(docs, n_docs, sizes, num_epochs, max_num_samples, min_num_sent, /*remain_min=*/2,
                                   /*long_check_needs_min=*/false, first_target, next_target, emit);
  if (out) {
    if (bad || static_cast<int64_t>(rows) != out_rows) {
      fsb::set_error("build_mapping: out holds %ld rows, the mapping has %lu", (long)out_rows, (unsigned long)rows);
      return -1;
    }
    shuffle_rows(out, dtype, out_rows, 3, seed);
  }
  return static_cast<int64_t>(rows);
}

int64_t fsb_index_build_blocks_mapping(const int64_t* docs, int64_t n_docs, const int32_t* sizes, const int32_t* titles_sizes,
                                       int32_t num_epochs, uint64_t max_num_samples, int32_t max_seq_length, int32_t seed,
                                       int use_one_sent_blocks, int dtype, void* out, int64_t out_rows) {
  if (!(docs && sizes && titles_sizes) || n_docs < 0 || num_epochs <= 0 || max_seq_length <= 1 || seed <= 0 ||
      (dtype != FSB_U32 && dtype != FSB_U64)) {
    fsb::set_error("build_blocks_mapping: need docs/sizes/titles_sizes, num_epochs > 0, max_seq_length > 1, seed > 0, "
                   "dtype FSB_U32 | FSB_U64");
    return -1;
  }
  const int32_t min_num_sent = use_one_sent_blocks ? 1 : 2;
  auto first_target = [&](int64_t doc) { return max_seq_length - titles_sizes[doc]; };
  auto next_target = [&](int64_t, int32_t t) { return t; };
  int64_t bad = 0;
  auto emit = [&](uint64_t row, int64_t a, int64_t b, int64_t doc, int32_t block_id, int32_t) {
    if (!out) return;
    if (static_cast<int64_t>(row) >= out_rows) { ++bad; return; }
    put(out, dtype, 4 * row, a);
    put(out, dtype, 4 * row + 1, b);
    put(out, dtype, 4 * row + 2, doc);
    put(out, dtype, 4 * row + 3, block_id);
  };
  const uint64_t rows = walk_spans(docs, n_docs, sizes, num_epochs, max_num_samples, min_num_sent, /*remain_min=*/min_num_sent,
                                   /*long_check_needs_min=*/true, first_target, next_target, emit);
  if (out) {
    if (bad || static_cast<int64_t>(rows) != out_rows) {
      fsb::set_error("build_blocks_mapping: out holds %ld rows, the mapping has %lu", (long)out_rows, (unsigned long)rows);
      return -1;
    }
    shuffle_rows(out, dtype, out_rows, 4, seed);
  }
  return static_cast<int64_t>(rows);
}

int fsb_index_build_blending_indices(uint8_t* dataset_index, int64_t* dataset_sample_index, const double* weights,
                                     int32_t num_datasets, int64_t size) {
  FSB_REQUIRE(dataset_index && dataset_sample_index && weights, "build_blending_indices: null pointer");
  FSB_REQUIRE(num_datasets > 0 && num_datasets <= 256 && size >= 0, "build_blending_indices: 1..256 datasets (uint8 index), "
              "size >= 0 (got %d, %ld)", num_datasets, (long)size);
  // counts are kept as doubles next to the integers: every value is an integer below 2^53, so the double is exactly the
  // int64 -> double conversion the reference performs in its inner loop, without performing it num_datasets times per sample
  int64_t taken[256] = {0};
  double taken_f[256] = {0.0};
  for (int64_t i = 0; i < size; ++i) {
    // the dataset that lags its quota i * w the most gets the next sample; ties go to the lowest index (helpers.cpp:62-78)
    const double at = i > 0 ? static_cast<double>(i) : 1.0;
    int32_t pick = 0;
    double worst = weights[0] * at - taken_f[0];
    for (int32_t d = 1; d < num_datasets; ++d) {
      const double lag = weights[d] * at - taken_f[d];
      if (lag > worst) { worst = lag; pick = d; }
    }
    dataset_index[i] = static_cast<uint8_t>(pick);
    dataset_sample_index[i] = taken[pick]++;
    taken_f[pick] += 1.0;
  }
  return 0;
}

}  // extern "C"
