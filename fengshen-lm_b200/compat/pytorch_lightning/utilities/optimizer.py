def optimizers_to_device(optimizers, device):
    return optimizers
