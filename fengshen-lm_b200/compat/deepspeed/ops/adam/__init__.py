"""FusedAdam / DeepSpeedCPUAdam as hyper-parameter carriers: `configure_optimizers` (model_utils.py:62-72) builds them,
the LR scheduler mutates their param_groups, and fsb200's ZeroEngine executes the update (fsb_adamw_flat on the shard)."""
import torch


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, bias_correction=True, betas=(0.9, 0.999), eps=1e-8, adam_w_mode=True,
                 weight_decay=0.0, amsgrad=False, set_grad_none=True):
        if amsgrad:
            raise RuntimeError("FusedAdam does not support the AMSGrad variant.")
        if not adam_w_mode:
            raise NotImplementedError("fsb200: only adam_w_mode=True (decoupled weight decay) is implemented")
        super().__init__(params, dict(lr=lr, bias_correction=bias_correction, betas=betas, eps=eps,
                                      weight_decay=weight_decay))
        self.adam_w_mode = 1

    def step(self, closure=None):
        """No-op: the sharded update runs in fsb200.engine.ZeroEngine.step() (one fused kernel per bucket slice)."""
        return None

    def zero_grad(self, set_to_none=True):
        return None


class DeepSpeedCPUAdam(FusedAdam):
    def __init__(self, model_params, lr=1e-3, bias_correction=True, betas=(0.9, 0.999), eps=1e-8, weight_decay=0,
                 amsgrad=False, adamw_mode=True, fp32_optimizer_states=True):
        raise NotImplementedError("fsb200: ZeRO-offload (DeepSpeedCPUAdam) is out of scope; optimizer state lives in HBM")
