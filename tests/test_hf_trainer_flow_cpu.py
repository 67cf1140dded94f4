"""CPU: the host side of the HF-backed recipes end to end — fsb200.launch, transformers rebinding, example script, compat
Trainer, ZeroEngine (ZeRO-1 + clipping / default strategy), DeepSpeed-layout checkpoints, resume (step / sample / LR-schedule
position), metric-formatted checkpoint names, HF-style export — with the model classes replaced by tests/toy_models.py and the
engine kernels by tests/cpu_kernels.py. The bodies are the ones tests/test_hf_scripts_gpu.py runs on the real models."""
import os
import sys

import pytest

import hf_fixtures as F
import hf_recipes as R

for p in ("fengshen-lm_b200",):
    if os.path.join(F.ROOT, p) not in sys.path:
        sys.path.insert(0, os.path.join(F.ROOT, p))


@pytest.fixture
def launched_with_doubles(monkeypatch):
    saved_path = list(sys.path)
    for k in [k for k in sys.modules if k.split(".")[0] in ("fengshen", "pytorch_lightning", "deepspeed")]:
        monkeypatch.delitem(sys.modules, k)
    import cpu_kernels
    import fsb200.engine as engine
    import fsb200.hf as hf
    import fsb200.launch as launch
    import toy_models as T

    class CpuEngine(engine.ZeroEngine):
        def __init__(self, model, **kw):
            kw.setdefault("kernels", cpu_kernels)
            super().__init__(model, **kw)

    monkeypatch.setattr(engine, "ZeroEngine", CpuEngine)
    monkeypatch.setattr(hf, "MegatronBertForPreTraining",
                        type("MegatronBertForPreTraining", (hf._HFSurface, T.ToyMegatronBert), {"config_name": "MegatronBertConfig"}))
    monkeypatch.setattr(hf, "GPT2LMHeadModel",
                        type("GPT2LMHeadModel", (hf._HFSurface, T.ToyGPT2), {"config_name": "GPT2Config"}))
    launch.prepare(R.EXAMPLE)
    yield hf
    hf.uninstall()
    sys.path[:] = saved_path


def test_erlangshen_recipe_host_flow(launched_with_doubles, tmp_path, monkeypatch):
    trainer, module = R.erlangshen_recipe(tmp_path, monkeypatch, min_drop=0.05, lr="2e-2")
    assert not module.model.flat.params.is_cuda


def test_wenzhong_recipe_host_flow(launched_with_doubles, tmp_path):
    R.wenzhong_recipe(tmp_path, min_drop=0.02, device="cpu")
