"""TEST INFRASTRUCTURE — writes tests/golden/gpt2_small.npz from transformers' own GPT2LMHeadModel (CPU, fp32, eager),
the implementation the reference's GPT-2 script calls (examples/wenzhong_qa/finetune_wenzhong.py:56).
Run:  python oracle/make_golden_hf.py      (records the transformers version it was generated with)"""
import os
import sys

import numpy as np
import torch
import transformers

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import hf_oracle as H  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
STEPS, LR, WD = 20, 1e-3, 0.1


def main():
    torch.set_num_threads(os.cpu_count())
    V = H.GPT2_SMALL["vocab_size"]
    model = H.build_gpt2(H.GPT2_SMALL)
    batch = H.make_lm_batch(V, 2, 96, seed=1234)
    out = model(input_ids=batch["input_ids"], labels=batch["labels"])
    out.loss.backward()
    rec = {"loss": np.array(out.loss.item()), "logits_slice": out.logits.detach()[:, :, :64].numpy(),
           "transformers_version": np.array(transformers.__version__)}
    for n, p in model.named_parameters():
        rec["gradnorm/" + n] = np.array(p.grad.norm().item())
    model = H.build_gpt2(H.GPT2_SMALL)
    batches = [H.make_lm_batch(V, 2, 96, seed=1234 + i) for i in range(4)]
    rec["loss_curve"] = np.array(H.train_gpt2(model, batches, STEPS, lr=LR, weight_decay=WD, warmup=2))
    rec["train_hparams"] = np.array([LR, 0.9, 0.999, 1e-8, WD])
    np.savez_compressed(os.path.join(OUT, "gpt2_small.npz"), **rec)
    print("gpt2_small: loss", out.loss.item(), "curve", rec["loss_curve"][0], "->", rec["loss_curve"][-1])
    # BERT-family golden losses (same seeds as tests/test_bert_gpu.py)
    Vb = H.BERT_SMALL["vocab_size"]
    lb = H.build_bert(H.BERT_SMALL)(**H.make_mlm_batch(Vb, 3, 96, seed=5, pad_tail=20)).loss.item()
    lm = H.build_megatron_bert(H.BERT_SMALL)(**H.make_mlm_batch(Vb, 3, 96, seed=6, nsp=True, pad_tail=11)).loss.item()
    np.savez_compressed(os.path.join(OUT, "bert_small.npz"), bert_loss=np.array(lb), megatron_loss=np.array(lm),
                        transformers_version=np.array(transformers.__version__))
    print("bert_small: bert", lb, "megatron", lm)
    # C5 mT5 (oracle row for round 2): loss, logits slice, gradient norms, and the relative-position bucket tables
    Vt = H.MT5_SMALL["vocab_size"]
    m5 = H.build_mt5(H.MT5_SMALL)
    tb = H.make_t5_batch(Vt, 2, 96, 48, seed=7, pad_tail=13)
    o5 = m5(**tb)
    o5.loss.backward()
    rec5 = {"loss": np.array(o5.loss.item()), "logits_slice": o5.logits.detach()[:, :, :64].numpy(),
            "transformers_version": np.array(transformers.__version__)}
    for n, p in m5.named_parameters():
        if p.grad is not None:
            rec5["gradnorm/" + n] = np.array(p.grad.norm().item())
    from transformers.models.t5.modeling_t5 import T5Attention
    rel = torch.arange(-300, 301)
    rec5["bucket_bidirectional"] = T5Attention._relative_position_bucket(rel, bidirectional=True, num_buckets=32,
                                                                         max_distance=128).numpy()
    rec5["bucket_causal"] = T5Attention._relative_position_bucket(rel, bidirectional=False, num_buckets=32,
                                                                  max_distance=128).numpy()
    np.savez_compressed(os.path.join(OUT, "mt5_small.npz"), **rec5)
    print("mt5_small: loss", o5.loss.item())


if __name__ == "__main__":
    main()
