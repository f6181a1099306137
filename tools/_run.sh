python -m pytest tests/test_gpu_conv_tc.py -q -x -k "wgrad or convT" > gpurun_out/g8_pytest_wg.log 2>&1; tail -5 gpurun_out/g8_pytest_wg.log
B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 30 --windows 5"
run() { tag=$1; shift; env "$@" $B --trace-file gpurun_out/g8_trace_$tag.txt > gpurun_out/g8_bench_$tag.json 2> gpurun_out/g8_bench_$tag.err; python - <<P
import json
d=json.load(open('gpurun_out/g8_bench_$tag.json'))
print('$tag', round(d['ms_per_step'],4), round(d['value']), d['loss_rel_err'], d['launches_per_step'], d['kernel_breakdown_ms'].get('pcm_wgrad3x3_tc'), d['kernel_breakdown_ms'].get('pcm_wgrad3x3_tc_grouped'), d['kernel_breakdown_ms'].get('pcm_unpack_grads_batched'))
P
}
run default A=1
run default2 A=1
run noatomic PCM_WGRAD_NOATOMIC=1
python -m pytest tests -m gpu -x -q > gpurun_out/g8_pytest.log 2>&1; tail -3 gpurun_out/g8_pytest.log
