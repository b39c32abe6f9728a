#!/bin/bash
# NCCL settings for the 13-17 MB gradient all-reduce of a step (run with gpurun --gpus N -- bash scripts/nccl_tune.sh N)
N=${1:-2}
out=gpurun_out/r02_nccl; mkdir -p $out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29590 \
      bench.py --gpus $N --workload din_tc --steps 60 --warmup 10 2> $out/${tag}_n$N.err | tail -1 > $out/${tag}_n$N.json
  python - <<EOF
import json
try:
    d = json.loads(open("$out/${tag}_n$N.json").read())
    print("$tag N=$N ms/step %.4f value %.4g" % (d["ms_per_step"], d["value"]))
except Exception as e:
    print("$tag failed", e)
EOF
}
run default NCCL_DEBUG=WARN
run ch32 NCCL_MIN_NCHANNELS=32
run simple_ch32 NCCL_PROTO=Simple NCCL_MIN_NCHANNELS=32
run ll128 NCCL_PROTO=LL128
run tree NCCL_ALGO=Tree
