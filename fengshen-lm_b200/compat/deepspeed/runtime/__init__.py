from . import activation_checkpointing, pipe  # noqa: F401
