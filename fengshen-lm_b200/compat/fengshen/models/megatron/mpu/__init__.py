"""Model-parallel unit at TP = PP = 1 (every BASELINE config): the query functions the hot-path scripts call
(fengshen/models/megatron/mpu/initialize.py). Tensor parallelism itself is SURVEY.md §8(f) rank 1 — not built."""
import torch.distributed as dist

_MP_WORLD, _MP_RANK = 1, 0
_INIT_PARAMS_IN_CUDA = True


def initialize_model_parallel(model_parallel_size=1, pipe_model_parallel_size=1, topology=None, fp32_allreduce=False):
    if model_parallel_size != 1 or pipe_model_parallel_size != 1:
        raise NotImplementedError("fsb200: tensor / pipeline model parallelism are outside the data-parallel hot path")


def model_parallel_is_initialized():
    return True


def set_model_parallel_world_size(n):
    global _MP_WORLD
    if n != 1:
        raise NotImplementedError("fsb200: model_parallel_size must be 1")
    _MP_WORLD = n


def set_model_parallel_rank(r):
    global _MP_RANK
    _MP_RANK = r


def set_init_params_in_cuda(flag):
    global _INIT_PARAMS_IN_CUDA
    _INIT_PARAMS_IN_CUDA = flag


def get_model_parallel_world_size():
    return _MP_WORLD


def get_model_parallel_rank():
    return _MP_RANK


def get_data_parallel_world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def get_data_parallel_rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def get_data_parallel_group():
    return None


def divide(a, b):
    assert a % b == 0, f"{a} is not divisible by {b}"
    return a // b
