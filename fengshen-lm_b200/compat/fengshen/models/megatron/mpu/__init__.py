"""Model-parallel unit: the group bookkeeping of fengshen/models/megatron/mpu/initialize.py:24-145 (tensor-parallel ranks are
CONSECUTIVE global ranks, data-parallel ranks stride by the tensor-parallel size — the Megatron grid the reference inherits) plus
the query functions the hot-path scripts call. Pipeline parallelism stays at 1. The tensor-parallel arithmetic itself lives in
fsb200.models.llama (column / row-parallel GEMMs, vocabulary-parallel embedding, gathered logits) and fsb200.engine."""
import torch.distributed as dist

_MP_WORLD, _MP_RANK = 1, 0
_MP_GROUP, _DP_GROUP = None, None
_INIT_PARAMS_IN_CUDA = True


def initialize_model_parallel(model_parallel_size=1, pipe_model_parallel_size=1, topology=None, fp32_allreduce=False):
    """mpu/initialize.py:37-118: every rank creates every group; rank r belongs to tensor-parallel group r // t and to the
    data-parallel group of the ranks with the same r % t."""
    global _MP_WORLD, _MP_RANK, _MP_GROUP, _DP_GROUP
    if pipe_model_parallel_size != 1:
        raise NotImplementedError("fsb200: pipeline model parallelism is outside the data-parallel hot path")
    t = int(model_parallel_size)
    if t == 1:
        _MP_WORLD, _MP_RANK, _MP_GROUP, _DP_GROUP = 1, 0, None, None
        return
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("fsb200 mpu: tensor parallelism needs an initialised torch.distributed process group")
    world, rank = dist.get_world_size(), dist.get_rank()
    if world % t:
        raise ValueError(f"world size {world} is not divisible by model_parallel_size {t}")
    for g in range(world // t):
        ranks = list(range(g * t, (g + 1) * t))
        grp = dist.new_group(ranks)
        if rank in ranks:
            _MP_GROUP = grp
    for r in range(t):
        ranks = list(range(r, world, t))
        grp = dist.new_group(ranks)
        if rank in ranks:
            _DP_GROUP = grp
    _MP_WORLD, _MP_RANK = t, rank % t


def model_parallel_is_initialized():
    return True


def set_model_parallel_world_size(n):
    global _MP_WORLD
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


def get_model_parallel_group():
    return _MP_GROUP


def get_data_parallel_world_size():
    if _DP_GROUP is not None:
        return dist.get_world_size(_DP_GROUP)
    w = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    return w // max(1, _MP_WORLD)


def get_data_parallel_rank():
    if _DP_GROUP is not None:
        return dist.get_rank(_DP_GROUP)
    r = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    return r // max(1, _MP_WORLD)


def get_data_parallel_group():
    return _DP_GROUP


def divide(a, b):
    assert a % b == 0, f"{a} is not divisible by {b}"
    return a // b
