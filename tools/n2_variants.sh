run() { tag=$1; shift; timeout 300 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 10 --warmup 3 --no-parity --no-e2e $EXTRA > gpurun_out/r2_n2_$tag.json 2> gpurun_out/r2_n2_$tag.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_n2_$tag.json")); print("$tag", round(d["value"]), round(d["ms_per_step"],2), round(d["roofline"]["achieved"],1))
except Exception as e: print("$tag failed", e)
PY
}
EXTRA="" run base X=1
EXTRA="" run ctas4 NCCL_MAX_CTAS=4
EXTRA="--comm-sms 4" run ctas4_sms4 NCCL_MAX_CTAS=4
EXTRA="--comm-sms 8" run sms8 X=1
EXTRA="" run ctas2 NCCL_MAX_CTAS=2
