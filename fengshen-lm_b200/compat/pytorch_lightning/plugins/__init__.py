class ClusterEnvironment:  # imported for type annotations only (megatron_deepspeed.py:25)
    pass
