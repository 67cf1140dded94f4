"""`fengshen.models.llama.modeling_llama.LlamaForCausalLM` -> fsb200.models.llama.LlamaForCausalLM, plus the
`from_pretrained(path, torch_dtype=...)` entry the scripts use (examples/ziya_llama/finetune_ziya_llama.py:102-107)."""
import json
import os

import torch

from fsb200.models.llama import LlamaForCausalLM as _FsbLlama
from .configuration_llama import LlamaConfig


class LlamaForCausalLM(_FsbLlama):
    config_class = LlamaConfig

    def __init__(self, config, **kw):
        for k, want in (("rotary_pct", 1), ("pos_emb", "rotary"), ("norm", "rmsnorm"), ("mlp_type", "llama"),
                        ("hidden_dropout", 0), ("attention_dropout", 0), ("use_bias_in_attn_linear", False)):
            have = getattr(config, k, want)
            if have != want:
                raise NotImplementedError(f"fsb200 LlamaForCausalLM: config.{k}={have!r} is outside the Ziya-LLaMA "
                                          f"hot path (only {want!r} is implemented)")
        if "tp_group" not in kw:   # built inside LightningModule.setup(), after the strategy initialised mpu (as the reference)
            from fengshen.models.megatron import mpu
            if mpu.get_model_parallel_world_size() > 1:
                kw["tp_group"] = mpu.get_model_parallel_group()
        super().__init__(config, **kw)

    @classmethod
    def from_pretrained(cls, path, torch_dtype=None, **kw):
        """Loads config.json + pytorch_model.bin (or the sharded index) written by the reference's save_pretrained /
        hf_to_fs.py. torch_dtype is accepted for signature parity; parameters are stored in bf16."""
        with open(os.path.join(path, "config.json")) as f:
            raw = json.load(f)
        raw.pop("torch_dtype", None); raw.pop("architectures", None); raw.pop("model_type", None)
        model = cls(LlamaConfig(**raw), **kw)
        idx = os.path.join(path, "pytorch_model.bin.index.json")
        files = [os.path.join(path, "pytorch_model.bin")]
        if os.path.exists(idx):
            with open(idx) as f:
                files = sorted({os.path.join(path, v) for v in json.load(f)["weight_map"].values()})
        sd = {}
        for fn in files:
            sd.update(torch.load(fn, map_location="cpu", weights_only=True))
        model.load_reference_state_dict(sd)
        return model

    def save_pretrained(self, path, **_):
        """HF-style export (what the scripts call after training, e.g. examples/pretrain_t5/pretrain_t5.py:105-112 for its
        model): config.json + pytorch_model.bin in the reference's key layout; `from_pretrained(path)` reads it back, and
        `fengshen.utils.llama_convert.fs_to_hf_state_dict` turns it into a transformers LLaMA checkpoint."""
        from fengshen.utils.llama_convert import save_pretrained_fs
        hook = getattr(self, "param_hook", None)
        eng = getattr(hook, "__self__", None)
        if eng is not None and hasattr(eng, "wait_params"):
            eng.wait_params()
        cfg = self.config.to_dict() if hasattr(self.config, "to_dict") else vars(self.config)
        save_pretrained_fs({k: v for k, v in self.state_dict().items()}, cfg, path)
