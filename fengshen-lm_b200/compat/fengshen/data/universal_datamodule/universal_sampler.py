"""Resumable Megatron-style batch samplers keyed by `consumed_samples` — behaviour of
fengshen/data/universal_datamodule/universal_sampler.py:22-125 restated (index arithmetic must match exactly: it decides
which samples each data-parallel rank sees and where a resumed run continues)."""
import torch


class PretrainingSampler:
    """Sequential: global batches of micro_batch*dp consecutive indices starting at consumed_samples; rank r takes the
    r-th micro-batch slice (universal_sampler.py:22-68)."""

    def __init__(self, total_samples, consumed_samples, micro_batch_size, data_parallel_rank, data_parallel_size,
                 drop_last=True):
        assert total_samples > 0, f'no sample to consume: {total_samples}'
        assert consumed_samples < total_samples, f'no samples left to consume: {consumed_samples}, {total_samples}'
        assert micro_batch_size > 0 and data_parallel_size > 0
        assert data_parallel_rank < data_parallel_size
        self.total_samples, self.consumed_samples = total_samples, consumed_samples
        self.micro_batch_size, self.data_parallel_rank = micro_batch_size, data_parallel_rank
        self.global_batch = micro_batch_size * data_parallel_size
        self.drop_last = drop_last

    def __len__(self):
        return self.total_samples // self.global_batch

    def __iter__(self):
        lo = self.data_parallel_rank * self.micro_batch_size
        hi = lo + self.micro_batch_size
        batch = []
        for idx in range(self.consumed_samples, self.total_samples):
            batch.append(idx)
            if len(batch) == self.global_batch:
                yield batch[lo:hi]
                batch = []
        if batch and not self.drop_last:
            yield batch[lo:hi]


class PretrainingRandomSampler:
    """Rank r owns the bucket [r*bucket, (r+1)*bucket) of the dataset, permuted by randperm(seed = epoch); a resumed
    run skips the consumed prefix of the permutation (universal_sampler.py:71-125)."""

    def __init__(self, total_samples, consumed_samples, micro_batch_size, data_parallel_rank, data_parallel_size, epoch):
        assert total_samples > 0 and micro_batch_size > 0 and data_parallel_size > 0
        assert data_parallel_rank < data_parallel_size
        self.total_samples, self.consumed_samples = total_samples, consumed_samples
        self.micro_batch_size = micro_batch_size
        self.data_parallel_rank, self.data_parallel_size = data_parallel_rank, data_parallel_size
        self.global_batch = micro_batch_size * data_parallel_size
        self.last_batch_size = total_samples % self.global_batch
        self.epoch = epoch

    def __len__(self):
        return self.total_samples // self.global_batch

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __iter__(self):
        active = self.total_samples - self.last_batch_size
        current_epoch_samples = self.consumed_samples % active
        assert current_epoch_samples % self.global_batch == 0
        bucket_size = (self.total_samples // self.global_batch) * self.micro_batch_size
        bucket_offset = current_epoch_samples // self.data_parallel_size
        start_idx = self.data_parallel_rank * bucket_size
        g = torch.Generator()
        g.manual_seed(self.epoch)
        perm = torch.randperm(bucket_size, generator=g).tolist()
        batch = []
        for x in perm[bucket_offset:]:
            batch.append(start_idx + x)
            if len(batch) == self.micro_batch_size:
                self.consumed_samples += self.global_batch
                yield batch
                batch = []
