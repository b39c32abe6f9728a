#!/bin/bash
# profiles/<tag>_ncu_summary.md + raw csv + traffic json from gpurun_out/<tag>/ncu (see round_measure.sh)
tag=${1:-r02}
src=gpurun_out/$tag/ncu
out=profiles/${tag}_ncu_summary.md
S=scripts/summarize_ncu.py
{
cat <<HDR
# Round-${tag#r} profiles (B200, ncu from CUDA 12.9, \`--clock-control none\`)

Commands (\`scripts/round_measure.sh\`, run through \`gpurun\`, each after the same command exited 0 without ncu):
\`\`\`
python bench.py --workload W --no-graph --steps 2 --warmup 3 --no-cpu-baseline --no-others --no-aten
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file launches_W.csv <same>
ncu --set full --clock-control none --import-source on -k regex:<kernels> -s N -c M -o full_W <same>
\`\`\`
Raw data: \`${tag}_launches_*.csv\` (every launch of the run), \`${tag}_full_*.raw.csv\` (\`ncu --page raw --csv\` of the
full captures).  Summaries below are produced by \`scripts/summarize_ncu.py\` / \`scripts/make_ncu_report.sh\`.  ncu
times are cold-cache and serialised: compare shares, not absolutes; the bench numbers in \`${tag}_bench.md\` are the
measured ones.  Phase profiles (last section) come from \`scripts/phase_profile.py\` (clock64 per phase, thread 0).

HDR
echo "## Launch list — DCN step (eager, B = 8192; one step between two forward launches)"; echo
python $S launches $src/launches_dcn.csv crossnet_fwd_kernel
echo; echo "## Launch list — DIN step, tensor-core activation unit (eager, B = 8192, T = 50)"; echo
python $S launches $src/launches_din_tc.csv din_fwd_tc_kernel
echo; echo "## Launch list — AFM step, tensor-core attention (eager, B = 8192, F = 10, D = 32, A = 128)"; echo
python $S launches $src/launches_afm.csv afm_fwd_tc_kernel
echo; echo "## Launch list — BST step, tensor-core block (eager, B = 8192, T = 20, 4 heads)"; echo
python $S launches $src/launches_bst_tc.csv bst_fwd_tc_kernel
for w in din_tc afm bst_tc dcn; do
  echo; echo "## Full capture — $w"
  python $S full $src/full_$w.ncu-rep
  ncu -i $src/full_$w.ncu-rep --page raw --csv > profiles/${tag}_full_$w.raw.csv 2>/dev/null
done
for w in dcn din_tc afm bst_tc; do cp $src/launches_$w.csv profiles/${tag}_launches_$w.csv; done
if [ -f gpurun_out/${tag}_phase_bst_tc_bwd.txt ]; then
  echo; echo "## Phase profile — \`tc::bst_bwd_tc_kernel\` (\`scripts/phase_profile.py run bst_tc bwd\`, clock64 of thread 0 per phase, B = 8192)"; echo
  echo '```'; grep -v Warning gpurun_out/${tag}_phase_bst_tc_bwd.txt | tail -12; echo '```'
fi
# per-source-line attribution of the warp-state samples (scripts/ncu_by_line.py): needs the cubins of the library
# the capture was made with -- only done when the library in the tree still is that one
if [ -n "$NCU_BY_LINE" ]; then
  tmp=$(mktemp -d); (cd $tmp && cuobjdump -xelf all $OLDPWD/implementation-*_b200/librank_b200.so >/dev/null 2>&1; for c in *.sm_100a.cubin; do nvdisasm -g -c $c > ${c%.sm_100a.cubin}.sass 2>/dev/null; done)
  byline() {  # capture, ncu kernel regex, sass file, mangled substring, source, title
    ncu -i $src/full_$1.ncu-rep --page source --csv --kernel-name regex:$2 > $tmp/k.csv 2>/dev/null
    echo; echo "### \`$2\` by source line (warp-state samples, top 14)"; echo; echo '```'
    python scripts/ncu_by_line.py $tmp/k.csv $tmp/$3.sass $4 implementation-*_b200/csrc/$5 14 | cut -c1-200
    echo '```'
  }
  echo; echo "## Where the warps wait — per-source-line attribution of the full captures"
  byline din_tc din_fwd_tc_kernel din din_fwd_tc_kernel din.cu
  byline din_tc din_bwd_tc_kernel din din_bwd_tc_kernel din.cu
  byline afm afm_fwd_tc_kernel afm_tc afm_fwd_tc_kernelILi32E afm_tc.cu
  byline afm afm_bwd_tc_kernel afm_tc afm_bwd_tc_kernelILi32E afm_tc.cu
  rm -rf $tmp
fi
echo; echo "## SASS evidence of the Blackwell paths in the shipped library (\`cuobjdump -sass librank_b200.so\`)"; echo
echo '```'
cuobjdump -sass implementation-*_b200/librank_b200.so 2>/dev/null | grep -oE "UTCHMMA|LDTM|UTCBAR|UBLKCP|UTMALDG|SYNCS\.[A-Z.]+|LDGSTS" | sort | uniq -c
echo '```'
echo "(\`UTCHMMA\` = tcgen05.mma kind::f16 — AFM, DIN and BST kernels; \`LDTM\` = tcgen05.ld; \`UTCBAR\` = tcgen05.commit; \`UBLKCP\` = cp.async.bulk"
echo "global -> shared on an mbarrier, the TMA bulk copy of DIN's weight tiles; \`SYNCS.ARRIVE.TRANS\` = mbarrier.arrive.expect_tx;"
echo "\`SYNCS.PHASECHK\` = mbarrier.try_wait; \`LDGSTS\` = cp.async row / index staging.)"
} > $out.new
echo "wrote $out.new"
