"""GPU: an example script written purely against the reference's import surface (examples/pretrain_ziya_llama.py, the
structure of fengshen/examples/ziya_llama/finetune_ziya_llama.py) runs through the compat Trainer on the fsb200 engine:
loss goes down, the DeepSpeed-layout checkpoint directory is written, and a resumed run continues from it."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "fengshen-lm_b200", "compat"), os.path.join(ROOT, "examples")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _args(tmp, extra=()):
    return ["--hidden_size", "256", "--num_layers", "2", "--num_heads", "4", "--vocab_size", "512",
            "--max_seq_length", "64", "--num_samples", "8", "--train_batchsize", "4", "--max_steps", "12",
            "--max_epochs", "-1", "--learning_rate", "1e-3", "--adam_beta2", "0.95", "--warmup_steps", "2",
            "--strategy", "deepspeed_stage_2", "--default_root_dir", str(tmp), "--save_ckpt_path", str(tmp / "ckpt"),
            "--load_ckpt_path", str(tmp / "ckpt" / "last.ckpt"), "--every_n_train_steps", "6", "--save_last",
            "--log_every_n_steps", "1", "--dataloader_workers", "0", *extra]


def test_example_script_trains_checkpoints_and_resumes(tmp_path):
    import pretrain_ziya_llama as ex
    trainer, module = ex.main(_args(tmp_path))
    assert trainer.global_step == 12
    import json
    losses = [json.loads(l)["train/loss"] for l in open(os.path.join(trainer.logger.save_dir, "metrics.jsonl"))]
    assert losses[-1] < losses[0] - 0.2, losses
    ck = tmp_path / "ckpt" / "last.ckpt" / "checkpoint"
    assert (ck / "mp_rank_00_model_states.pt").exists() and (ck / "zero_pp_rank_0_mp_rank_00_optim_states.pt").exists()
    state = torch.load(ck / "mp_rank_00_model_states.pt", map_location="cpu", weights_only=False)
    assert "module" in state and any(k.endswith("attention.query_key_value.weight") for k in state["module"])
    assert state["global_samples"] == 12 * 4  # steps x micro-batch x world x GA
    w_before = module.model.flat.params.clone()
    # resume: continues at step 12 and trains 4 more steps from the saved weights / optimizer shard / LR schedule
    trainer2, module2 = ex.main(_args(tmp_path, ("--max_steps", "16")))
    assert trainer2.global_step == 16
    assert getattr(module2, "consumed_samples", None) == 48          # on_load_checkpoint hook (finetune_ziya_llama.py:180-183)
    assert not torch.equal(module2.model.flat.params, w_before)


def test_resumed_run_continues_the_sample_stream(tmp_path, monkeypatch):
    """the loader built for get_total_steps() (before the checkpoint is read) must not survive the resume —
    the first batch of a resumed run is the one that follows the last consumed sample of the permutation
    (PretrainingRandomSampler, universal_sampler.py:104-125), not sample 0 again."""
    import pretrain_ziya_llama as ex
    seen = []
    orig = ex.SyntheticCollator.__call__

    def spy(self, samples):
        out = orig(self, samples)
        seen.append(out["input_ids"][:, :4].clone())
        return out
    monkeypatch.setattr(ex.SyntheticCollator, "__call__", spy)
    common = ("--replace_sampler_ddp", "False", "--num_samples", "64", "--every_n_train_steps", "4")
    straight = tmp_path / "a"; straight.mkdir()
    ex.main(_args(straight, common + ("--max_steps", "8")))
    want = [t.clone() for t in seen]
    seen.clear()
    resumed = tmp_path / "b"; resumed.mkdir()
    ex.main(_args(resumed, common + ("--max_steps", "4")))
    first_leg = len(seen)
    trainer2, module2 = ex.main(_args(resumed, common + ("--max_steps", "8")))
    assert trainer2.global_step == 8 and module2.consumed_samples == 16
    # the 4 batches consumed after the resume are batches 5..8 of the uninterrupted run
    resumed_batches = seen[first_leg:]
    assert len(resumed_batches) >= 4
    for a, b in zip(resumed_batches[:4], want[4:8]):
        assert torch.equal(a, b), "resumed run replayed already-consumed samples"
    # and the first resumed step used the restored learning rate, not the schedule's step-0 value
    assert trainer2.optimizers[0].param_groups[0]["lr"] > 0
