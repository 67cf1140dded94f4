from .t5_datasets import (T5SpanCorruptionCollator, compute_input_and_target_lengths)  # noqa: F401
