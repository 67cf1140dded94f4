"""`BlendableDataset` — fengshen/data/megatron_dataloader/blendable_dataset.py:26-64: a weighted interleaving of datasets whose
schedule comes from `helpers.build_blending_indices` (here: fsb_index_build_blending_indices of libfsb200.so). Unlike the
reference it also works before torch.distributed is initialised (the reference calls get_rank() unconditionally)."""
import time

import numpy as np
import torch

from fengshen.data.megatron_dataloader import helpers
from fengshen.data.megatron_dataloader.utils import print_rank_0


class BlendableDataset(torch.utils.data.Dataset):
    def __init__(self, datasets, weights):
        self.datasets = datasets
        if len(datasets) != len(weights):
            raise ValueError("BlendableDataset: one weight per dataset")
        if not 0 < len(datasets) < 255:
            raise ValueError("BlendableDataset: 1..254 datasets (the index is uint8)")
        self.size = sum(len(d) for d in datasets)
        w = np.array(weights, dtype=np.float64)
        total = np.sum(w)
        if not total > 0.0:
            raise ValueError("BlendableDataset: weights must sum to a positive number")
        w /= total
        t0 = time.time()
        self.dataset_index = np.zeros(self.size, dtype=np.uint8)
        self.dataset_sample_index = np.zeros(self.size, dtype=np.int64)
        helpers.build_blending_indices(self.dataset_index, self.dataset_sample_index, w, len(datasets), self.size, False)
        print_rank_0('> elapsed time for building blendable dataset indices: {:.2f} (sec)'.format(time.time() - t0))

    def __len__(self):
        return self.size

    def __getitem__(self, idx):
        return self.datasets[self.dataset_index[idx]][self.dataset_sample_index[idx]]
