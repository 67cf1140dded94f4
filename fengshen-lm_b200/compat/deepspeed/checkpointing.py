"""deepspeed.checkpointing symbols re-exported by fengshen/models/megatron/mpu/random.py:18-37. The reference never
recomputes activations, so `checkpoint` simply calls the function; the RNG tracker is inert at TP=1."""
import contextlib

_MODEL_PARALLEL_RNG_TRACKER_NAME = "model-parallel-rng"


class CudaRNGStatesTracker:
    def reset(self):
        pass

    def add(self, name, seed):
        pass

    @contextlib.contextmanager
    def fork(self, name=_MODEL_PARALLEL_RNG_TRACKER_NAME):
        yield


_CUDA_RNG_STATE_TRACKER = CudaRNGStatesTracker()


def get_cuda_rng_tracker():
    return _CUDA_RNG_STATE_TRACKER


def model_parallel_cuda_manual_seed(seed):
    import torch
    torch.manual_seed(seed)


def _set_cuda_rng_state(new_state, device=-1):
    pass


def checkpoint(function, *args):
    return function(*args)


def configure(mpu_, deepspeed_config=None, partition_activations=None, contiguous_checkpointing=None,
              num_checkpoints=None, checkpoint_in_cpu=None, synchronize=None, profile=None):
    return None
