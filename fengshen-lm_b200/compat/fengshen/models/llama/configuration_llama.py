"""LlamaConfig with the constructor of fengshen/models/llama/configuration_llama.py:68-109 plus the ad-hoc fields that
utils/llama_convert/hf_to_fs.py:31-53 writes into released Ziya config.json files (read as plain attributes)."""
from transformers.configuration_utils import PretrainedConfig


class LlamaConfig(PretrainedConfig):
    model_type = "llama"

    def __init__(self, vocab_size=32000, hidden_size=4096, num_hidden_layers=32, num_attention_heads=32,
                 intermediate_size=11008, hidden_act="silu", rotary_pct=1, rotary_emb_base=10000,
                 max_position_embeddings=2048, initializer_range=0.02, rms_norm_epsilon=1.0e-6, use_cache=True,
                 pad_token_id=0, bos_token_id=1, eos_token_id=2, tie_word_embeddings=False,
                 use_parallel_residual=True, **kwargs):
        self.vocab_size = vocab_size
        self.max_position_embeddings = max_position_embeddings
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.hidden_act = hidden_act
        self.rotary_pct = rotary_pct
        self.rotary_emb_base = rotary_emb_base
        self.initializer_range = initializer_range
        self.rms_norm_epsilon = rms_norm_epsilon
        self.use_cache = use_cache
        self.use_parallel_residual = use_parallel_residual
        # Ziya recipe defaults (hf_to_fs.py:31-53); a config.json may override them through **kwargs
        for k, v in dict(llama_mlp_multiple_of=256, hidden_dropout=0, attention_dropout=0, pos_emb="rotary",
                         norm="rmsnorm", mlp_type="llama", use_bias_in_attn_linear=False,
                         attention_config=[[["flash"], "all"]]).items():
            setattr(self, k, kwargs.pop(k, v))
        super().__init__(pad_token_id=pad_token_id, bos_token_id=bos_token_id, eos_token_id=eos_token_id,
                         tie_word_embeddings=tie_word_embeddings, **kwargs)
