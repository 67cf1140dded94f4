"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference LLaMA (fengshen/models/llama + fengshen/models/megatron) from
/root/reference on CPU, so golden vectors can be generated and the CPU restatement in oracle/ can be pinned.

Only usable in the authoring container (the GPU box has no /root/reference). Nothing on the product path imports this.
The three workarounds are exactly those verified in SURVEY.md §8c:
  (1) `import fengshen` executes fengshen/__init__.py:16-19, which imports model families that break on
      transformers 5.x  ->  register bare namespace modules for `fengshen` and `fengshen.models` instead;
  (2) fengshen/models/megatron/mpu/random.py:18-19 hard-imports deepspeed  ->  a stub exposing the six symbols used;
  (3) fengshen/models/megatron/layers/transformer.py:334 passes device=torch.cuda.current_device()  ->  patch it to 'cpu'.
"""
import contextlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("FSB_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "fengshen", "models", "llama"))


def _install_shims():
    if "fengshen.models.llama.modeling_llama" in sys.modules:
        return
    # (1) namespace packages that skip fengshen/__init__.py and fengshen/models/__init__.py
    for name, rel in (("fengshen", "fengshen"), ("fengshen.models", "fengshen/models")):
        if name not in sys.modules or getattr(sys.modules[name], "_fsb_oracle_shim", False) is False:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
            m._fsb_oracle_shim = True
            sys.modules[name] = m
    # (2) deepspeed stub: mpu/random.py:18-37 re-exports these from deepspeed.runtime.activation_checkpointing.checkpointing
    if "deepspeed" not in sys.modules:
        ds = types.ModuleType("deepspeed")
        ck = types.ModuleType("deepspeed.checkpointing")

        class _Tracker:
            @contextlib.contextmanager
            def fork(self, name=None):
                yield

            def add(self, *a, **k):
                pass

            def reset(self):
                pass

        _tracker = _Tracker()
        ck._MODEL_PARALLEL_RNG_TRACKER_NAME = "model-parallel-rng"
        ck._CUDA_RNG_STATE_TRACKER = _tracker
        ck._set_cuda_rng_state = lambda *a, **k: None
        ck.checkpoint = lambda fn, *a: fn(*a)
        ck.model_parallel_cuda_manual_seed = lambda seed: None
        ck.get_cuda_rng_tracker = lambda: _tracker
        rt = types.ModuleType("deepspeed.runtime")
        ac = types.ModuleType("deepspeed.runtime.activation_checkpointing")
        ac.checkpointing = ck
        rt.activation_checkpointing = ac
        ds.checkpointing = ck
        ds.runtime = rt
        ds._fsb_oracle_stub = True
        sys.modules["deepspeed"] = ds
        sys.modules["deepspeed.checkpointing"] = ck
        sys.modules["deepspeed.runtime"] = rt
        sys.modules["deepspeed.runtime.activation_checkpointing"] = ac
        sys.modules["deepspeed.runtime.activation_checkpointing.checkpointing"] = ck
    # (3) CPU "current device"
    torch.cuda.current_device = lambda: "cpu"


def load_reference_llama():
    """Returns (LlamaForCausalLM class, LlamaConfig class, mpu module) of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    _install_shims()
    from fengshen.models.megatron import mpu
    from fengshen.models.llama.configuration_llama import LlamaConfig
    from fengshen.models.llama.modeling_llama import LlamaForCausalLM
    mpu.set_model_parallel_world_size(1)
    mpu.set_model_parallel_rank(0)
    mpu.set_init_params_in_cuda(False)  # pattern of utils/llama_convert/hf_to_fs.py:84-86
    return LlamaForCausalLM, LlamaConfig, mpu


def make_reference_config(LlamaConfig, vocab_size, hidden_size, num_layers, num_heads, max_pos=2048, eps=1e-6,
                          attention="global", dtype=torch.float32):
    """LlamaConfig + the ad-hoc fields of utils/llama_convert/hf_to_fs.py:31-53 (Ziya recipe). `attention="global"`
    selects the baddbmm/softmax/bmm path (the flash path needs the 3P flash_attn_cuda v1 extension)."""
    ff = int(2 * hidden_size * 4 / 3)
    ff = 256 * ((ff + 255) // 256)  # transformer.py:589-590
    cfg = LlamaConfig(vocab_size=vocab_size, hidden_size=hidden_size, num_hidden_layers=num_layers,
                      num_attention_heads=num_heads, intermediate_size=ff, hidden_act="silu", rotary_pct=1,
                      rotary_emb_base=10000, max_position_embeddings=max_pos, initializer_range=0.02,
                      rms_norm_epsilon=eps, torch_dtype=dtype, use_cache=False, pad_token_id=0, bos_token_id=1,
                      eos_token_id=2, tie_word_embeddings=False, use_parallel_residual=False)
    cfg.llama_mlp_multiple_of = 256
    cfg.init_method = "small_init"
    cfg.hidden_dropout = 0
    cfg.output_layer_init_method = "wang_init"
    cfg.pos_emb = "rotary"
    cfg.norm = "rmsnorm"
    cfg.gpt_j_residual = False
    cfg.gpt_j_tied = False
    cfg.apply_query_key_layer_scaling = False
    cfg.attention_softmax_in_fp32 = False
    cfg.scaled_masked_softmax_fusion = False
    cfg.scaled_upper_triang_masked_softmax_fusion = False
    cfg.bias_gelu_fusion = False
    cfg.attention_dropout = 0
    cfg.output_layer_parallelism = "column"
    cfg.eod_mask_loss = False
    cfg.bias_dropout_fusion = False
    cfg.attention_config = [[[attention], "all"]]
    cfg.mlp_type = "llama"
    cfg.use_bias_in_attn_linear = False
    cfg.lora = False
    return cfg
