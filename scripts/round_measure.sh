#!/bin/bash
# Round measurement pass, run on the GPU box through gpurun:  bash scripts/round_measure.sh [tag]
# Writes everything under gpurun_out/<tag>/ : bench JSON lines, ncu launch lists and full captures.
tag=${1:-r01}
out=gpurun_out/$tag
mkdir -p $out/bench $out/ncu
WORKLOADS="dcn deepfm fwfm afm afm_fp32 din din_softmax din_tc din_softmax_tc bst deepcrossing"
for w in $WORKLOADS; do
  timeout 300 python bench.py --workload $w --steps 30 --warmup 5 2> $out/bench/$w.err | tail -1 > $out/bench/$w.json
done
timeout 300 python bench.py --impl reference --workload dcn --steps 5 --warmup 3 2> $out/bench/reference_dcn.err | tail -1 > $out/bench/reference_dcn.json

# ncu: launch lists of one eager step (after the same command ran clean above), then full captures
for w in dcn din_tc afm; do
  cmd="python bench.py --workload $w --no-graph --steps 2 --warmup 3 --no-cpu-baseline"
  timeout 300 $cmd > $out/ncu/plain_$w.log 2>&1 || continue
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/ncu/launches_$w.csv $cmd > $out/ncu/ncu_launch_$w.log 2>&1
done
full() {  # workload, kernel regex, skip, count
  cmd="python bench.py --workload $1 --no-graph --steps 2 --warmup 3 --no-cpu-baseline"
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $out/ncu/full_$1 $cmd > $out/ncu/ncu_full_$1.log 2>&1
}
full afm 'afm_(fwd|bwd)_tc_kernel' 6 2
full din_tc 'din_(fwd|bwd)_tc_kernel' 6 2
full fwfm 'fwfm_(fwd|bwd)_kernel' 6 2
full dcn 'crossnet_(fwd|bwd)_kernel|small_field_sort|segment_' 20 5
ls -la $out/bench $out/ncu
