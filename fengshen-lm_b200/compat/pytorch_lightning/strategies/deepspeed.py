"""`pytorch_lightning.strategies.DeepSpeedStrategy` as the hot path sees it: a bag of ZeRO settings with a `.config`
dict (read by fengshen/models/model_utils.py:62-64: `'offload_optimizer' in strategy.config['zero_optimization']`),
discoverable from $PL_DEEPSPEED_CONFIG_PATH (megatron_deepspeed.py:53) and from the registry strings
`deepspeed_stage_{1,2,3}[_offload]`. Execution is delegated to fsb200.engine.ZeroEngine by the Trainer."""
import json
import os

_DEEPSPEED_AVAILABLE = True  # the `deepspeed` shim in fengshen-lm_b200/compat provides the symbols that are imported


class DeepSpeedStrategy:
    strategy_name = "deepspeed"
    DEEPSPEED_ENV_VAR = "PL_DEEPSPEED_CONFIG_PATH"

    def __init__(self, accelerator=None, zero_optimization=True, stage=2, remote_device="cpu", offload_optimizer=False,
                 offload_parameters=False, offload_params_device="cpu", nvme_path="/local_nvme",
                 params_buffer_count=5, params_buffer_size=100_000_000, max_in_cpu=1_000_000_000,
                 offload_optimizer_device="cpu", optimizer_buffer_count=4, block_size=1048576, queue_depth=8,
                 single_submit=False, overlap_events=True, thread_count=1, pin_memory=False, sub_group_size=1e12,
                 contiguous_gradients=True, overlap_comm=True, allgather_partitions=True, reduce_scatter=True,
                 allgather_bucket_size=200_000_000, reduce_bucket_size=200_000_000, zero_allow_untested_optimizer=True,
                 logging_batch_size_per_gpu="auto", config=None, logging_level=None, parallel_devices=None,
                 cluster_environment=None, loss_scale=0, initial_scale_power=16, loss_scale_window=1000, hysteresis=2,
                 min_loss_scale=1, partition_activations=False, cpu_checkpointing=False,
                 contiguous_memory_optimization=False, synchronize_checkpoint_boundary=False,
                 load_full_weights=False, precision_plugin=None, process_group_backend=None, **_):
        if offload_optimizer or offload_parameters:
            raise NotImplementedError("fsb200: ZeRO-offload is out of scope (SURVEY.md §8): optimizer state lives in HBM")
        if stage not in (1, 2):
            raise NotImplementedError(f"fsb200: ZeRO stage {stage} is not implemented (stages 1 and 2 are)")
        self.stage = stage
        self.overlap_comm = overlap_comm
        self.process_group_backend = process_group_backend
        self.config = self._load_config(config)
        if self.config is None:
            self.config = {"zero_optimization": {"stage": stage, "contiguous_gradients": contiguous_gradients,
                                                 "overlap_comm": overlap_comm, "reduce_scatter": reduce_scatter,
                                                 "allgather_bucket_size": allgather_bucket_size,
                                                 "reduce_bucket_size": reduce_bucket_size}}
        zo = self.config.setdefault("zero_optimization", {"stage": stage})
        if "offload_optimizer" in zo or "offload_param" in zo:
            raise NotImplementedError("fsb200: ZeRO-offload is out of scope: optimizer state lives in HBM")
        self.stage = int(zo.get("stage", stage))
        if self.stage not in (1, 2):
            raise NotImplementedError(f"fsb200: ZeRO stage {self.stage} is not implemented (stages 1 and 2 are)")

    def _load_config(self, config):
        if config is None and self.DEEPSPEED_ENV_VAR in os.environ:
            config = os.environ[self.DEEPSPEED_ENV_VAR]
        if isinstance(config, (str, os.PathLike)):
            if not os.path.isfile(config):
                raise FileNotFoundError(f"You passed in a path to a DeepSpeed config but the path does not exist: {config}")
            with open(config) as f:
                config = json.load(f)
        return config

    @property
    def gradient_clipping(self):
        return float(self.config.get("gradient_clipping", 0.0) or 0.0)

    @property
    def precision(self):
        if self.config.get("bf16", {}).get("enabled"):
            return "bf16"
        if self.config.get("fp16", {}).get("enabled"):
            return "16"
        return None


class DDPStrategy:
    """`--strategy ddp`: plain data parallelism; executed by the same engine (sharding the optimizer state changes no
    arithmetic — SURVEY.md Appendix D)."""
    strategy_name = "ddp"

    def __init__(self, **_):
        self.config = {"zero_optimization": {"stage": 1}}
        self.stage = 1
        self.overlap_comm = True
        self.gradient_clipping = 0.0
        self.precision = None


def strategy_from_string(name):
    if name is None or name in ("ddp", "auto", "ddp_find_unused_parameters_false"):
        return DDPStrategy()
    if name == "deepspeed":
        return DeepSpeedStrategy()
    if name.startswith("deepspeed_stage_"):
        rest = name[len("deepspeed_stage_"):]
        if "offload" in rest:
            raise NotImplementedError("fsb200: ZeRO-offload strategies are out of scope")
        return DeepSpeedStrategy(stage=int(rest[0]))
    raise ValueError(f"unknown strategy {name!r}")
