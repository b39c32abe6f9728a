#!/bin/bash
# Last measurement pass of a round on the final tree (one gpurun call): GPU test suite, the default bench line,
# ncu full captures + launch lists of the kernels that changed.  Writes under gpurun_out/<tag>/.
tag=${1:-r02}
out=gpurun_out/$tag
mkdir -p $out/bench $out/ncu
timeout 300 python -m pytest tests -q -m gpu 2>&1 | tail -300 > $out/pytest_gpu.txt
tail -1 $out/pytest_gpu.txt
timeout 420 python bench.py 2> $out/bench/default.err | tail -1 > $out/bench/default.json
python scripts/show_bench.py $out/bench/default.json 2>&1 | head -3
cmd() { echo "python bench.py --workload $1 --no-graph --steps 2 --warmup 3 --no-cpu-baseline --no-others --no-aten"; }
full() {  # workload, kernel regex, skip, count
  timeout 200 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $out/ncu/full_$1 $(cmd $1) > $out/ncu/ncu_full_$1.log 2>&1
}
launches() {
  timeout 100 $(cmd $1) > $out/ncu/plain_$1.log 2>&1 || return
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/ncu/launches_$1.csv $(cmd $1) > $out/ncu/ncu_launch_$1.log 2>&1
}
full din_tc 'din_(fwd|bwd)_tc_kernel|din_weight_tiles' 6 3
launches din_tc
launches afm
full afm 'afm_(fwd|bwd)_tc_kernel|afm_weight_tiles' 6 3
launches bst_tc
ls -la $out/ncu | tail -20
