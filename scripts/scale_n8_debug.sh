export RANK_B200_TRACE=1 RANK_B200_TRACE_AFTER=50 NCCL_DEBUG=WARN
out=gpurun_out/r02_scale; mkdir -p $out
timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 scripts/sharded_bst.py --rows 100000000 --steps 30 2> $out/sharded_bst_n8.err | tail -1 > $out/sharded_bst_n8.json
cat $out/sharded_bst_n8.json; grep "rank 0\]" $out/sharded_bst_n8.err | tail -4; grep -A14 "most recent call first" $out/sharded_bst_n8.err | head -40
