python -m pytest tests -m gpu -x -q > gpurun_out/g15_pytest_all.log 2>&1; tail -3 gpurun_out/g15_pytest_all.log
B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 30 --windows 5"
run() { tag=$1; shift; env "$@" $B --trace-file gpurun_out/g15_trace_$tag.txt > gpurun_out/g15_bench_$tag.json 2> gpurun_out/g15_bench_$tag.err; python - <<P
import json
d=json.load(open('gpurun_out/g15_bench_$tag.json'))
print('$tag', round(d['ms_per_step'],4), round(d['value']), d['loss_rel_err'], d['launches_per_step'], d['final_loss'])
P
}
run h0 PCM_HEAD_MSE=0
run h1 PCM_HEAD_MSE=1
run h0b PCM_HEAD_MSE=0
run h1b PCM_HEAD_MSE=1
grep -E "head|mse" gpurun_out/g15_trace_h1.txt gpurun_out/g15_trace_h0.txt | awk '{print $1, $2, $NF}'
