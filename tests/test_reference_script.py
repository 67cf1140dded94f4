"""The UNMODIFIED reference example script against the fsb200 backend (north_star: "an existing example script runs unchanged").

`/root/reference/fengshen/examples/ziya_llama/finetune_ziya_llama.py` is executed AS A SCRIPT (runpy, `__main__`) with only
`fengshen-lm_b200/compat` put in front of `sys.path` — its `fengshen.*`, `pytorch_lightning`, `deepspeed` imports resolve to the
compat surface, its `Llama` LightningModule, `LlamaSFTCollator`, argument parser, `UniversalDataModule(--train_file ...)`,
`DeepSpeedStrategy(...)`, `UniversalCheckpoint`, `Trainer.fit` / `Trainer.predict` run as written. Fixtures: a tiny saved model
directory (config.json + pytorch_model.bin in the reference's key layout), a JSON-lines SFT corpus, and a stub tokenizer
(LlamaTokenizer.from_pretrained needs a sentencepiece model file that cannot be fetched offline).

The reference tree is not shipped to the GPU box (and must not be copied into this repository), so:
  * the CPU half (import, parser, collator, data module — everything before the model touches CUDA) runs wherever
    /root/reference exists (the authoring container, `-m "not gpu"`);
  * the GPU half (fit 10 steps with a falling loss + a resumable checkpoint, then predict_step -> generate) runs wherever BOTH
    a GPU and the reference tree exist, and skips otherwise (FSB_REFERENCE_ROOT overrides the path).
tests/test_compat_gpu.py drives the same Trainer path on the GPU box with a script of identical structure."""
import json
import os
import runpy
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FSB_REFERENCE_ROOT", "/root/reference")
SCRIPT = os.path.join(REF, "fengshen", "examples", "ziya_llama", "finetune_ziya_llama.py")
needs_ref = pytest.mark.skipif(not os.path.exists(SCRIPT), reason="reference tree not present on this machine")


class StubTokenizer:
    """What the script uses of LlamaTokenizer: __call__(text, add_special_tokens=...).input_ids, eos_token_id,
    convert_ids_to_tokens / convert_tokens_to_string, decode. Byte-level, ids 3..258; 2 = </s>."""
    eos_token_id, pad_token_id, bos_token_id = 2, 0, 1

    def __call__(self, text, add_special_tokens=True, **_):
        ids = [3 + b for b in text.encode("utf-8")]
        from types import SimpleNamespace
        return SimpleNamespace(input_ids=([self.bos_token_id] if add_special_tokens else []) + ids)

    def convert_ids_to_tokens(self, ids):
        return [int(i) for i in (ids.tolist() if hasattr(ids, "tolist") else ids)]

    def convert_tokens_to_string(self, toks):
        return bytes([t - 3 for t in toks if 3 <= t < 259]).decode("utf-8", "replace")

    def decode(self, ids, skip_special_tokens=False):
        return self.convert_tokens_to_string(self.convert_ids_to_tokens(ids))


def _fixtures(tmp):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import llama_oracle as O
    V, h, L, nh = 264, 256, 2, 4
    mdir = tmp / "model"
    mdir.mkdir()
    json.dump({"vocab_size": V, "hidden_size": h, "num_hidden_layers": L, "num_attention_heads": nh,
               "max_position_embeddings": 2048, "rms_norm_epsilon": 1e-6, "torch_dtype": "float16"},
              open(mdir / "config.json", "w"))
    torch.save(O.make_weights(V, h, L, seed=0), mdir / "pytorch_model.bin")
    data = tmp / "train.json"
    with open(data, "w") as f:
        for i in range(64):
            f.write(json.dumps({"task": "qa", "prompt": [f"question {i} about the sea?"],
                                "output": [f"answer {i}: the sea is wide and salty."]}, ensure_ascii=False) + "\n")
    return mdir, data


def _argv(tmp, mdir, data, extra=()):
    return [SCRIPT, "--model_path", str(mdir), "--tokenizer_path", "stub", "--train_file", str(data), "--val_file", str(data),
            "--test_file", str(data), "--train_batchsize", "4", "--val_batchsize", "4", "--test_batchsize", "2",
            "--max_seq_length", "128", "--max_steps", "10", "--max_epochs", "-1", "--learning_rate", "1e-3",
            "--adam_beta2", "0.95", "--warmup_steps", "2", "--strategy", "deepspeed_stage_2",
            "--default_root_dir", str(tmp), "--save_ckpt_path", str(tmp / "ckpt"), "--load_ckpt_path", str(tmp / "ckpt" / "last.ckpt"),
            "--every_n_train_steps", "5", "--save_last", "--log_every_n_steps", "1", "--dataloader_workers", "0",
            "--precision", "bf16", *extra]


def _run_script(argv, monkeypatch):
    compat = os.path.join(ROOT, "fengshen-lm_b200", "compat")
    for p in (os.path.dirname(SCRIPT), os.path.join(ROOT, "fengshen-lm_b200"), compat):
        monkeypatch.syspath_prepend(p)
    for k in [k for k in sys.modules if k.split(".")[0] in ("fengshen", "pytorch_lightning", "deepspeed", "llama_generate")]:
        monkeypatch.delitem(sys.modules, k)
    import transformers
    monkeypatch.setattr(transformers.LlamaTokenizer, "from_pretrained", classmethod(lambda cls, *a, **k: StubTokenizer()))
    monkeypatch.setattr(sys, "argv", argv)
    return runpy.run_path(SCRIPT, run_name="__main__")


@needs_ref
def test_reference_script_imports_and_its_collator_runs_against_compat(tmp_path, monkeypatch):
    """CPU: everything the script does before the model touches CUDA."""
    compat = os.path.join(ROOT, "fengshen-lm_b200", "compat")
    for p in (os.path.dirname(SCRIPT), os.path.join(ROOT, "fengshen-lm_b200"), compat):
        monkeypatch.syspath_prepend(p)
    for k in [k for k in sys.modules if k.split(".")[0] in ("fengshen", "pytorch_lightning", "deepspeed", "llama_generate")]:
        monkeypatch.delitem(sys.modules, k)
    mod = runpy.run_path(SCRIPT, run_name="not_main")                     # the module body: imports + class definitions
    assert mod["LlamaForCausalLM"].__module__.startswith("fengshen.models.llama")
    import argparse
    import pytorch_lightning as pl
    p = argparse.ArgumentParser()
    p = mod["add_module_args"](p)
    p = pl.Trainer.add_argparse_args(p)
    p = mod["UniversalDataModule"].add_data_specific_args(p)
    p = mod["Llama"].add_module_specific_args(p)
    p = mod["UniversalCheckpoint"].add_argparse_args(p)
    mdir, data = _fixtures(tmp_path)
    args = p.parse_args(_argv(tmp_path, mdir, data)[1:])
    tok = StubTokenizer()
    coll = mod["LlamaSFTCollator"](tokenizer=tok, max_seq_length=args.max_seq_length)
    dm = mod["UniversalDataModule"](tokenizer=tok, args=args, collate_fn=coll)
    ds = dm.datasets[args.train_datasets_field]
    batch = coll([ds[0], ds[1], ds[2]])
    assert set(batch) == {"input_ids", "attention_mask", "position_ids", "labels"}
    B, S = batch["input_ids"].shape
    assert B == 3 and S <= 128 and batch["labels"].shape == (B, S)
    assert (batch["labels"][0] != -100).sum() > 0 and (batch["labels"][0] == -100).sum() > 0      # prompt masked, answer kept
    assert torch.equal(batch["position_ids"][0], torch.arange(S))
    module = mod["Llama"](args, tokenizer=tok)
    assert module.hparams.model_path == str(mdir)
    strat = mod["DeepSpeedStrategy"](tensor_model_parallel_size=1, pipe_model_parallel_size=1, mpu_seed=42)
    assert strat.stage == 2


@needs_ref
@pytest.mark.gpu
def test_reference_script_unchanged_fit_resume_predict(tmp_path, monkeypatch, capsys):
    mdir, data = _fixtures(tmp_path)
    ns = _run_script(_argv(tmp_path, mdir, data), monkeypatch)
    trainer = ns["trainer"]
    assert trainer.global_step == 10
    losses = [json.loads(l)["train/loss"] for l in open(os.path.join(trainer.logger.save_dir, "metrics.jsonl"))]
    assert losses[-1] < losses[0] - 0.3, losses
    ck = tmp_path / "ckpt" / "last.ckpt" / "checkpoint"
    assert (ck / "mp_rank_00_model_states.pt").exists()
    ns2 = _run_script(_argv(tmp_path, mdir, data, ("--max_steps", "14")), monkeypatch)      # resumes from last.ckpt
    assert ns2["trainer"].global_step == 14 and ns2["model"].consumed_samples == 40
    _run_script(_argv(tmp_path, mdir, data, ("--do_eval_only",)), monkeypatch)               # predict_step -> generate
    assert "ans:" in capsys.readouterr().out
