#!/bin/bash
# Scaling run on N GPUs of one box (gpurun --gpus N):  bash scripts/scale_run.sh N [tag]
# Default workload (DIN, tensor-core unit) + DeepFM under torchrun, and the row-sharded BST configuration.
N=${1:-8}
tag=${2:-r02}
out=gpurun_out/${tag}_scale
mkdir -p $out
port=29540
for w in din_tc deepfm; do
  port=$((port + 1))
  if [ "$N" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --workload $w --steps 100 --warmup 10 --no-others --no-aten --no-cpu-baseline \
        2> $out/${w}_n$N.err | tail -1 > $out/${w}_n$N.json
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $N --workload $w --steps 100 --warmup 10 2> $out/${w}_n$N.err | tail -1 > $out/${w}_n$N.json
  fi
  python - <<EOF
import json
try:
    d = json.loads(open("$out/${w}_n$N.json").read())
    print("$w N=$N", "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "ms/step %.4f" % d["ms_per_step"])
except Exception as e:
    print("$w N=$N failed:", e)
EOF
done
if [ "$N" != 1 ]; then
  rows=$((12500000 * N))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29560 \
      scripts/sharded_bst.py --rows $rows --steps 30 2> $out/sharded_bst_n$N.err | tail -1 > $out/sharded_bst_n$N.json
  cat $out/sharded_bst_n$N.json
fi
