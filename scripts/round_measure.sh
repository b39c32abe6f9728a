#!/bin/bash
# Round measurement pass, run on the GPU box through gpurun:  bash scripts/round_measure.sh [tag] [part]
# Writes everything under gpurun_out/<tag>/ : bench JSON lines, ncu launch lists and full captures.
# Parts (each a few minutes; every command under its own timeout):  bench | ncu | full | all
tag=${1:-r02}
part=${2:-all}
out=gpurun_out/$tag
mkdir -p $out/bench $out/ncu
if [ "$part" = bench ] || [ "$part" = all ]; then
  # the default line: DIN (tensor-core activation unit), all legs + the other workloads' summaries
  timeout 900 python bench.py 2> $out/bench/default.err | tail -1 > $out/bench/default.json
  timeout 300 python bench.py --impl reference --steps 5 --warmup 3 2> $out/bench/reference.err | tail -1 > $out/bench/reference.json
  for w in dcn deepfm afm bst bst_tc; do
    timeout 300 python bench.py --workload $w --no-others --steps 100 2> $out/bench/$w.err | tail -1 > $out/bench/$w.json
  done
fi
if [ "$part" = ncu ] || [ "$part" = all ]; then
  # ncu: launch lists of eager steps (after the same command ran clean), kernel SHARES of a step
  for w in din_tc dcn bst_tc afm; do
    cmd="python bench.py --workload $w --no-graph --steps 2 --warmup 3 --no-cpu-baseline --no-others --no-aten"
    timeout 300 $cmd > $out/ncu/plain_$w.log 2>&1 || continue
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/ncu/launches_$w.csv $cmd > $out/ncu/ncu_launch_$w.log 2>&1
  done
fi
full() {  # workload, kernel regex, skip, count
  cmd="python bench.py --workload $1 --no-graph --steps 2 --warmup 3 --no-cpu-baseline --no-others --no-aten"
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $out/ncu/full_$1 $cmd > $out/ncu/ncu_full_$1.log 2>&1
}
if [ "$part" = full ] || [ "$part" = all ]; then
  full din_tc 'din_(fwd|bwd)_tc_kernel|din_weight_tiles' 6 3
  full bst_tc 'bst_(fwd|bwd)_tc_kernel' 4 2
  full dcn 'crossnet_(fwd|bwd)_kernel|direct_reduce' 9 3
fi
ls -la $out/bench $out/ncu
