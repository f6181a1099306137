CMD="python tools/microbench.py tails --flush --iters 2 --only convblock_tail"
$CMD > gpurun_out/p1_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:convblock_tail_bwd -c 1 -o gpurun_out/p1_tail_bwd -f $CMD > gpurun_out/p1_ncu_bwd.log 2>&1
tail -3 gpurun_out/p1_ncu_bwd.log
ncu --set full --clock-control none --import-source on -k regex:convblock_tail_fwd -c 1 -o gpurun_out/p1_tail_fwd -f $CMD > gpurun_out/p1_ncu_fwd.log 2>&1
tail -3 gpurun_out/p1_ncu_fwd.log
cat gpurun_out/p1_plain.log | head -12
