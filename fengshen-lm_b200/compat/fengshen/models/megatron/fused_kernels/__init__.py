"""The reference's load_fused_kernels() (fengshen/models/megatron/fused_kernels/__init__.py:32-44) imports its pybind
softmax extensions and exit()s when they are missing. Here it loads libfsb200.so and raises (no fallback)."""


def load_fused_kernels():
    from fsb200 import lib
    lib.load()
