"""NCCL communicator behind the C ABI's fsb_comm_* entry points (include/fsb200.h): what a host that is not PyTorch binds for
the ZeRO data path — bucketed gradient reduce-scatter, the fp32 norm all-reduce, the in-place parameter all-gather (the
collectives the reference delegates to DeepSpeed, fengshen/strategies/megatron_deepspeed.py:302-320). Here torch.distributed is
used ONCE, to hand rank 0's 128-byte NCCL id to the other ranks; every collective then goes through libfsb200.so on the
caller's current CUDA stream."""
import ctypes

import torch
import torch.distributed as dist

from . import lib as L


def _dt(t):
    if t.dtype == torch.float32:
        return L.F32
    if t.dtype == torch.bfloat16:
        return L.BF16
    raise RuntimeError(f"fsb200 comm: dtype {t.dtype} not supported (bf16 / fp32)")


class Communicator:
    def __init__(self, world, rank, process_group=None, device=None):
        self.world, self.rank = world, rank
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        ident = (ctypes.c_char * 128)()
        if rank == 0:
            L.call("fsb_comm_unique_id", ctypes.addressof(ident))
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                                   group=process_group)
        ident = (ctypes.c_char * 128).from_buffer_copy(box[0])
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            L.call("fsb_comm_init", ctypes.addressof(handle), ctypes.addressof(ident), world, rank)
        self.handle = handle

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def reduce_scatter(self, out, full):
        """out (numel n) = slice `rank` of the element-wise SUM over ranks of full (numel n * world)."""
        if full.numel() != out.numel() * self.world or full.dtype != out.dtype:
            raise RuntimeError("fsb200 comm.reduce_scatter: full must be world x out elements of the same dtype")
        L.call("fsb_comm_reduce_scatter", self.handle, full.data_ptr(), out.data_ptr(), out.numel(), _dt(out), self._stream())

    def all_gather(self, full, mine):
        if full.numel() != mine.numel() * self.world or full.dtype != mine.dtype:
            raise RuntimeError("fsb200 comm.all_gather: full must be world x mine elements of the same dtype")
        L.call("fsb_comm_all_gather", self.handle, mine.data_ptr(), full.data_ptr(), mine.numel(), _dt(mine), self._stream())

    def all_reduce(self, t):
        L.call("fsb_comm_all_reduce", self.handle, t.data_ptr(), t.data_ptr(), t.numel(), _dt(t), self._stream())

    def destroy(self):
        if self.handle:
            L.call("fsb_comm_destroy", self.handle)
            self.handle = ctypes.c_void_p()
