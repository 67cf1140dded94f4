"""GPU half of the HF-backed recipes (C3 Erlangshen-MegatronBERT, C2 Wenzhong-GPT2): scripts written against the reference's
import surface — `from transformers import MegatronBertForPreTraining / GPT2LMHeadModel`, pytorch_lightning, fengshen.* — run
through `fsb200.launch` (compat packages + fsb200.hf.install()) on the fsb200 engine and libfsb200.so. The reference tree is not
on the GPU box, so the scripts are examples/pretrain_erlangshen_bert.py (structure of pretrain_erlangshen.py:126-240) and a
LightningModule with the structure of finetune_wenzhong.py:37-113 (tests/hf_recipes.py); tests/test_hf_scripts_cpu.py runs the
CPU half of the UNMODIFIED scripts where the reference exists, tests/test_hf_trainer_flow_cpu.py the same bodies on test doubles."""
import os
import sys

import pytest

import hf_fixtures as F
import hf_recipes as R

pytestmark = pytest.mark.gpu


@pytest.fixture
def launched(monkeypatch):
    monkeypatch.syspath_prepend(os.path.join(F.ROOT, "fengshen-lm_b200"))
    saved_path = list(sys.path)
    import fsb200.hf as hf
    import fsb200.launch as launch
    launch.prepare(R.EXAMPLE)
    yield hf
    hf.uninstall()
    sys.path[:] = saved_path


def test_erlangshen_recipe_trains_checkpoints_and_resumes(launched, tmp_path, monkeypatch):
    # transformers' MegatronBertForPreTraining on CPU with the same data / lr: 6.9 -> ~5.0 over these 16 steps
    trainer, module = R.erlangshen_recipe(tmp_path, monkeypatch, min_drop=0.5)
    assert type(module.model).__module__ == "fsb200.hf" and module.model.flat.params.is_cuda


def test_wenzhong_recipe_structure_trains(launched, tmp_path):
    trainer, module = R.wenzhong_recipe(tmp_path, min_drop=1.0, device="cuda")     # 48 near-identical rows: memorised quickly
    assert type(module.model).__module__ == "fsb200.hf" and module.model.flat.params.is_cuda


def test_t5_recipe_structure_trains_exports_and_resumes_mid_epoch(launched, tmp_path):
    # transformers' MT5ForConditionalGeneration on CPU with the same data / lr: 135 -> ~36 over the first 6 steps
    trainer, module = R.t5_recipe(tmp_path, min_drop=10.0)
    assert type(module.model).__module__ == "fsb200.hf" and module.model.flat.params.is_cuda
