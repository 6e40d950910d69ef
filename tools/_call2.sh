set -x
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 300 python tools/gpu_small_check.py --perf > gpurun_out/small_check.log 2>&1; tail -40 gpurun_out/small_check.log
timeout 300 python tools/gpu_layer_check.py > gpurun_out/layer_check.log 2>&1; grep -E "layer_16|ALL_OK|SOME_FAILED|Error|error" gpurun_out/layer_check.log | cut -c1-700
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-saturated > gpurun_out/r2_named_n1.json 2> gpurun_out/r2_named_n1.err; tail -c 2500 gpurun_out/r2_named_n1.json; tail -5 gpurun_out/r2_named_n1.err
