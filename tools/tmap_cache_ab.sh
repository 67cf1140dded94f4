#!/bin/sh
# A/B of the tensor-map memo on the launch-bound config (C1 BERT-base, batch 8 x 128) and on C2: eager step, no extra blocks.
for w in bert-base gpt2-110m; do
  for c in 0 1; do
    FSB_TMAP_CACHE=$c timeout 120 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-graph-block 2>/dev/null \
      | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w cache=$c', round(d['value']), 'tok/s', round(d['ms_per_step'],3), 'ms/step; e2e', round(d['e2e']['value']))"
  done
done
