timeout 200 python -m pytest tests/test_sharded.py -q -m gpu -k "exchange" -x 2>&1 | tail -5
out=gpurun_out/r02_scale; mkdir -p $out
for ex in 1 0; do
  RANK_B200_OCCURRENCE_EXCHANGE=$ex timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29597 bench.py --gpus 2 --workload deepfm --steps 100 --warmup 10 2> $out/deepfm_ex${ex}_n2.err | tail -1 > $out/deepfm_ex${ex}_n2.json
  python -c "
import json
d=json.loads(open('$out/deepfm_ex${ex}_n2.json').read()); print('deepfm exchange=$ex N=2 ms/step %.4f value %.4g loss %.6f' % (d['ms_per_step'], d['value'], d['loss']))"
  tail -2 $out/deepfm_ex${ex}_n2.err | cut -c1-300
done
