B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 30 --windows 5"
run() { tag=$1; shift; env "$@" $B --trace-file gpurun_out/g6_trace_$tag.txt > gpurun_out/g6_bench_$tag.json 2> gpurun_out/g6_bench_$tag.err; python - <<P
import json
d=json.load(open('gpurun_out/g6_bench_$tag.json'))
print('$tag', round(d['ms_per_step'],4), round(d['value']), d['loss_rel_err'], d['launches_per_step'], d['kernel_breakdown_ms'].get('pcm_wgrad3x3_tc'), d['kernel_breakdown_ms'].get('pcm_wgrad3x3_tc_grouped'))
P
}
run base PCM_WGRAD_GROUP=0 PCM_CONV_GROUP=0
run wg4 PCM_CONV_GROUP=0
run wg2 PCM_CONV_GROUP=0 PCM_WGRAD_GROUP=2
run wg4cg2 PCM_CONV_GROUP=2
run wg4cg4 PCM_CONV_GROUP=4
run wg4cg4c32 PCM_CONV_GROUP=4 PCM_CONV_GROUP_MAXC=32
run noatomic PCM_WGRAD_NOATOMIC=1
run noatomic_g0 PCM_WGRAD_NOATOMIC=1 PCM_WGRAD_GROUP=0 PCM_CONV_GROUP=0
