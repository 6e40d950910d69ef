set -x
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python tools/gpu_small_check.py --perf > gpurun_out/small_check14.log 2>&1; tail -2 gpurun_out/small_check14.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/small_check.json'))
for k,v in d.items():
    if not v.get('ok') or k.startswith('perf'): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
PY
timeout 900 python -m pytest tests/test_gpu.py -x -q -m gpu -k "small or swapab or wgrad or graph or engine" 2>&1 | tail -5
Q="--steps 20 --warmup 5 --no-saturated --no-parity --no-nccl-baseline --no-extra-configs"
timeout 300 python bench.py $Q > gpurun_out/n1_14.json 2> gpurun_out/n1_14.err
python -c "
import json;d=json.loads(open('gpurun_out/n1_14.json').read().strip().splitlines()[-1]);print('N1', round(d['ms_per_step'],3), round(d['value']), d['exposed_comm_wait_ms_per_rank'], d['hottest_expert_rows_per_layer'])" || tail -5 gpurun_out/n1_14.err
