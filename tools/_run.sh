python -m pytest tests/test_gpu_parity.py tests/test_gpu_trainer.py -x -q > gpurun_out/g12_pytest.log 2>&1; tail -3 gpurun_out/g12_pytest.log
B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 30 --windows 5"
run() { tag=$1; shift; env "$@" $B --trace-file gpurun_out/g12_trace_$tag.txt > gpurun_out/g12_bench_$tag.json 2> gpurun_out/g12_bench_$tag.err; python - <<P
import json
d=json.load(open('gpurun_out/g12_bench_$tag.json'))
print('$tag', round(d['ms_per_step'],4), round(d['value']), d['loss_rel_err'], d['launches_per_step'])
P
}
run old PCM_B200_LIB=$PWD/tools/_libold.so
run new A=1
run oldb PCM_B200_LIB=$PWD/tools/_libold.so
run newb A=1
grep -E "tail_bwd|gn_silu_img_bwd" gpurun_out/g12_trace_old.txt | awk '{print $1, $NF}' > /tmp/a.txt; grep -E "tail_bwd|gn_silu_img_bwd" gpurun_out/g12_trace_new.txt | awk '{print $1}' > /tmp/b.txt; paste /tmp/b.txt /tmp/a.txt
