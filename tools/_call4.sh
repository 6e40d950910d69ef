set -x
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 300 python tools/gpu_small_check.py --perf > gpurun_out/small_check.log 2>&1; grep -E "wgrad_adam|perf_|ALL_OK|SOME_FAILED|Error" gpurun_out/small_check.log | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-saturated --no-parity --no-nccl-baseline > gpurun_out/r2_named_n1_dyn.json 2> gpurun_out/r2_named_n1_dyn.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_named_n1_dyn.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline'],d['e2e']['ms_per_step'])"; tail -3 gpurun_out/r2_named_n1_dyn.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 300 --csv --log-file gpurun_out/launches_named.csv python bench.py --gpus 1 --steps 3 --warmup 3 --no-saturated --no-parity --no-nccl-baseline --no-e2e --no-graph > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches_named.csv
