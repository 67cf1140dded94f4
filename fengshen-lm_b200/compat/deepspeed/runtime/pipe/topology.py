class PipeModelDataParallelTopology:
    """Axes ('pipe', 'data', 'model') -> rank mapping (megatron_deepspeed.py:347-351). Only pp = mp = 1 is supported."""

    def __init__(self, num_pp, num_mp, num_dp):
        if num_pp != 1 or num_mp != 1:
            raise NotImplementedError("fsb200: pipeline / tensor parallelism are outside the data-parallel hot path")
        self.num_pp, self.num_mp, self.num_dp = num_pp, num_mp, num_dp

    def get_dim(self, axis):
        return {"pipe": self.num_pp, "model": self.num_mp, "data": self.num_dp}[axis]

    def world_size(self):
        return self.num_pp * self.num_mp * self.num_dp
