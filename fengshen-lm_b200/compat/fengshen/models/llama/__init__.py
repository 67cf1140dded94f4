from .configuration_llama import LlamaConfig  # noqa: F401
from .modeling_llama import LlamaForCausalLM  # noqa: F401
