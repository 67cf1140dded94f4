from .universal_datamodule import UniversalDataModule  # noqa: F401
from .universal_sampler import PretrainingRandomSampler, PretrainingSampler  # noqa: F401
