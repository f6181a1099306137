python -m pytest tests/test_gpu_trainer.py tests/test_gpu_parity.py -x -q > gpurun_out/g16_pytest.log 2>&1; tail -3 gpurun_out/g16_pytest.log
B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 30 --windows 5"
run() { tag=$1; shift; env "$@" $B --trace-file gpurun_out/g16_trace_$tag.txt > gpurun_out/g16_bench_$tag.json 2> gpurun_out/g16_bench_$tag.err; python - <<P
import json
d=json.load(open('gpurun_out/g16_bench_$tag.json'))
print('$tag', round(d['ms_per_step'],4), round(d['value']), d['loss_rel_err'], d['launches_per_step'], d['final_loss'])
P
}
run a A=1
run b A=1
