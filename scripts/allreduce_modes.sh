#!/bin/bash
# The step's gradient all-reduce through NCCL vs torch symmetric-memory kernels (gpurun --gpus N -- bash scripts/allreduce_modes.sh N)
N=${1:-2}
out=gpurun_out/r02_allreduce; mkdir -p $out
for mode in nccl two_shot multimem auto; do
  RANK_B200_ALLREDUCE=$mode timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29593 bench.py --gpus $N --workload din_tc --steps 60 --warmup 10 2> $out/${mode}_n$N.err | tail -1 > $out/${mode}_n$N.json
  python - <<EOF
import json
try:
    d = json.loads(open("$out/${mode}_n$N.json").read())
    print("$mode N=$N ms/step %.4f value %.4g allreduce=%s loss %.6f" % (d["ms_per_step"], d["value"], d["config"].get("allreduce"), d["loss"]))
except Exception as e:
    print("$mode failed", e)
EOF
  grep -i "rank_b200:\|Error" $out/${mode}_n$N.err | head -3
done
