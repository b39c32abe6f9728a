export RANK_B200_TRACE=1
out=gpurun_out/r02_scale; mkdir -p $out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 scripts/sharded_bst.py --rows 50000000 --steps 30 2> $out/sharded_bst_n4.err | tail -1 > $out/sharded_bst_n4.json
cat $out/sharded_bst_n4.json; grep "rank 0" $out/sharded_bst_n4.err | tail -3
for w in din_tc deepfm; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 4 --workload $w --steps 100 --warmup 10 2> $out/${w}_n4.err | tail -1 > $out/${w}_n4.json
  python -c "
import json
d=json.loads(open('$out/${w}_n4.json').read()); print('$w N=4', d['value'], d['e2e']['value'], d['ms_per_step'])"
done
