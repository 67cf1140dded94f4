"""ZeRO-1/2 step engine: what the reference delegates to DeepSpeed through
`fengshen/strategies/megatron_deepspeed.py:302-320` (`deepspeed.initialize`) and `models/model_utils.py:62-72`
(`FusedAdam(adam_w_mode=True)`), re-designed for one NVSwitch domain:

  backward  : the model reports each bucket (one transformer layer) as soon as its gradients are final; the engine
              reduce-scatters that bucket's bf16 gradients on a side stream (NCCL over NVLink), overlapping the rest of
              backward. The 1/(world*GA) average is folded into dlogits upstream, so the wire op is a plain SUM.
  GA        : reduced shards are accumulated into an fp32 shard buffer per micro-step (ZeRO-2 semantics: full-size
              gradients never outlive a micro-step's bucket).
  step      : local sum-of-squares of the owned shard -> scalar all-reduce -> clip coefficient (device side, no host
              sync) -> fused AdamW on the fp32 {master, m, v} shard writing the bf16 parameters of the owned slice in
              place -> per-bucket in-place all-gather.
Sharding is per bucket (fsb200/flat.py): rank r owns slice r of every bucket. With world_size == 1 the same code runs
with the collectives compiled out of the data path (config 2, single GPU).

`kernels` is the object providing accumulate / sumsq / clip_coef / adamw_flat (default: fsb200.ops, i.e. the CUDA
library; it raises if the library or a GPU is missing). CPU unit tests of the sharding logic inject a test double.
"""
import torch
import torch.distributed as dist


class ZeroEngine:
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1, grad_clip=0.0, ga_steps=1,
                 process_group=None, stage=2, kernels=None, overlap_comm=True):
        self.model = model
        self.flat = model.flat
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.grad_clip, self.ga_steps, self.stage = float(grad_clip), int(ga_steps), int(stage)
        if stage not in (1, 2):
            raise ValueError("ZeroEngine supports ZeRO stage 1 and 2 (optimizer-state / gradient sharding)")
        self.pg = process_group
        self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
        self.world = dist.get_world_size(process_group) if self.distributed else 1
        self.rank = dist.get_rank(process_group) if self.distributed else 0
        if self.world != self.flat.world_size:
            raise ValueError(f"model was laid out for world_size={self.flat.world_size}, process group has {self.world}")
        if kernels is None:
            from . import ops as kernels  # CUDA library; no fallback
        self.k = kernels
        dev = self.flat.params.device
        self.device = dev
        n = self.flat.shard_numel
        self.master = torch.empty(n, dtype=torch.float32, device=dev)
        for i in range(len(self.flat.buckets)):
            self._seg(self.master, i).copy_(self.flat.bucket_slice(i, self.rank).float())
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.acc32 = torch.zeros(n, dtype=torch.float32, device=dev) if self.ga_steps > 1 else None
        self.recv16 = torch.zeros(n, dtype=torch.bfloat16, device=dev) if self.world > 1 else None
        self.sumsq = torch.zeros((), dtype=torch.float32, device=dev)
        self.coef = torch.ones((), dtype=torch.float32, device=dev)
        self.grad_norm = torch.zeros((), dtype=torch.float32, device=dev)
        self.use_streams = dev.type == "cuda" and overlap_comm and self.world > 1
        self.comm_stream = torch.cuda.Stream(device=dev) if self.use_streams else None
        self.micro = 0
        self.step_count = 0
        self.comm_bytes = 0
        model.grad_hook = self._on_bucket
        model.loss_scale = 1.0 / (self.ga_steps * self.world)

    # rank-local segment of bucket i inside a shard-sized buffer
    def _seg(self, buf, i):
        off = self.flat.shard_offsets[i]
        per = self.flat.buckets[i][2] // self.world
        return buf[off:off + per]

    # ---- backward side -----------------------------------------------------------------------------------------
    def _on_bucket(self, name):
        i = self.flat.bucket_index[name]
        first = self.micro == 0
        if self.world > 1:
            full = self.flat.bucket_view(i, grad=True)
            out = self._seg(self.recv16, i)
            if self.use_streams:
                self.comm_stream.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(self.comm_stream):
                    dist.reduce_scatter_tensor(out, full, op=dist.ReduceOp.SUM, group=self.pg)
                    if self.acc32 is not None:
                        self.k.accumulate(self._seg(self.acc32, i), out, 1.0, overwrite=first)
            else:
                dist.reduce_scatter_tensor(out, full, op=dist.ReduceOp.SUM, group=self.pg)
                if self.acc32 is not None:
                    self.k.accumulate(self._seg(self.acc32, i), out, 1.0, overwrite=first)
            self.comm_bytes += full.numel() * 2 * (self.world - 1) // self.world
        elif self.acc32 is not None:
            self.k.accumulate(self._seg(self.acc32, i), self.flat.bucket_view(i, grad=True), 1.0, overwrite=first)

    def backward_done(self):
        """Call once after each micro-batch's loss.backward()."""
        self.micro += 1

    def _grad_seg(self, i):
        if self.acc32 is not None:
            return self._seg(self.acc32, i)
        if self.world > 1:
            return self._seg(self.recv16, i)
        return self.flat.bucket_view(i, grad=True)

    # ---- optimizer step ----------------------------------------------------------------------------------------
    def step(self, lr=None, weight_decay=None):
        if self.micro != self.ga_steps:
            raise RuntimeError(f"ZeroEngine.step() after {self.micro} micro-batches, expected {self.ga_steps}")
        lr = self.lr if lr is None else lr
        wd = self.weight_decay if weight_decay is None else weight_decay
        if self.use_streams:
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)
        nb = len(self.flat.buckets)
        coef = None
        if self.grad_clip > 0.0:
            for i in range(nb):
                self.k.sumsq(self._grad_seg(i), self.sumsq, accumulate=(i > 0))
            if self.world > 1:
                dist.all_reduce(self.sumsq, op=dist.ReduceOp.SUM, group=self.pg)
            self.k.clip_coef(self.sumsq, self.grad_clip, self.coef, self.grad_norm)
            coef = self.coef
        self.step_count += 1
        for i in range(nb):
            decay_on = self.flat.buckets[i][3]
            self.k.adamw_flat(self._seg(self.master, i), self._seg(self.exp_avg, i), self._seg(self.exp_avg_sq, i),
                              self._grad_seg(i), self.flat.bucket_slice(i, self.rank), lr, self.betas[0], self.betas[1],
                              self.eps, wd if decay_on else 0.0, self.step_count, coef)
        if self.world > 1:
            for i in range(nb):
                dist.all_gather_into_tensor(self.flat.bucket_view(i), self.flat.bucket_slice(i, self.rank), group=self.pg)
                self.comm_bytes += self.flat.buckets[i][2] * 2 * (self.world - 1) // self.world
        self.micro = 0

    # ---- checkpoint (rank-local optimizer shard, the analogue of DeepSpeed's zero_pp_rank_*_optim_states.pt) ----
    def state_dict(self):
        return {"master": self.master, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "step": self.step_count, "world": self.world, "rank": self.rank}

    def load_state_dict(self, sd):
        if sd["world"] != self.world or sd["rank"] != self.rank:
            raise ValueError("optimizer shard was saved for a different (world, rank)")
        self.master.copy_(sd["master"]); self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_count = int(sd["step"])
