class PrecisionPlugin:  # imported for type annotations only (megatron_deepspeed.py:29)
    pass
