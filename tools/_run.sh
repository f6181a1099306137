B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 30 --windows 5"
run() { tag=$1; shift; env "$@" $B --trace-file gpurun_out/g11_trace_$tag.txt > gpurun_out/g11_bench_$tag.json 2> gpurun_out/g11_bench_$tag.err; python - <<P
import json
d=json.load(open('gpurun_out/g11_bench_$tag.json'))
print('$tag', round(d['ms_per_step'],4), round(d['value']), d['loss_rel_err'], d['launches_per_step'])
P
}
run w0 PCM_TAIL_WAVES=0
run w1 PCM_TAIL_WAVES=1
run w0b PCM_TAIL_WAVES=0
run w1b PCM_TAIL_WAVES=1
grep -E "tail|gn_silu" gpurun_out/g11_trace_w0.txt | awk '{print $1, $NF}' > /tmp/a.txt; grep -E "tail|gn_silu" gpurun_out/g11_trace_w1.txt | awk '{print $1}' > /tmp/b.txt; paste /tmp/b.txt /tmp/a.txt
python -m pytest tests -m gpu -x -q > gpurun_out/g11_pytest.log 2>&1; tail -3 gpurun_out/g11_pytest.log
