#!/bin/bash
# 8-GPU weak-scaling points (run with: gpurun --gpus 8 -- bash scripts/scale_n8.sh)
out=gpurun_out/r01_scale; mkdir -p $out
for w in din_tc dcn; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 8 --workload $w --steps 20 --warmup 5 2> $out/${w}_n8.err | tail -1 > $out/${w}_n8.json
  python -c "import json;d=json.load(open('$out/${w}_n8.json'));print('$w', d['n_gpus'], round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']))"
done
