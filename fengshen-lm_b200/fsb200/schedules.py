"""Host-side learning-rate schedules (fp64 Python, like the reference's): the two the hot-path scripts select.

polynomial_lr : transformers.get_polynomial_decay_schedule_with_warmup, chosen by `get_scheduler('polynomial', ...)`
                at fengshen/models/model_utils.py:94-96,250-252 (power 1, lr_end = --min_learning_rate).
linear_lr     : transformers.get_linear_schedule_with_warmup, examples/wenzhong_qa/finetune_wenzhong.py:102-104 and
                examples/pretrain_bert/pretrain_bert.py:148-173.
Both are evaluated at the scheduler's step counter, which starts at 0 for the first optimizer step.
"""


def polynomial_lr(step, base_lr, warmup, total, lr_end=1e-7, power=1.0):
    if step < warmup:
        return base_lr * float(step) / float(max(1, warmup))
    if step > total:
        return lr_end
    decay_steps = total - warmup
    if decay_steps <= 0:
        return lr_end
    remaining = 1 - (step - warmup) / decay_steps
    return (base_lr - lr_end) * remaining ** power + lr_end


def linear_lr(step, base_lr, warmup, total):
    if step < warmup:
        return base_lr * float(step) / float(max(1, warmup))
    return base_lr * max(0.0, float(total - step) / float(max(1, total - warmup)))
