"""Host-side sample assembly for the MegatronBERT / BERT pretraining recipes (C1 / C3): the helpers the reference's
`ErLangShenCollator` calls (fengshen/examples/pretrain_erlangshen_bert/pretrain_erlangshen.py:35-123). Restated with the reference's
exact random-number consumption (one numpy RandomState drives segment split, truncation and masking), pinned by
tests/golden/bert_collator.npz which oracle/make_golden_bert_collator.py generates from the unmodified reference."""
