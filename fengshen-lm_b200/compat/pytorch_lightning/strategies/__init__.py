from .deepspeed import DeepSpeedStrategy, DDPStrategy, _DEEPSPEED_AVAILABLE, strategy_from_string  # noqa: F401
