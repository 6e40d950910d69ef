set -x
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
export CUDA_VISIBLE_DEVICES=0
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_layer_check.py check_layer check_layer_small > gpurun_out/san_memcheck_layer.log 2>&1; grep -E "layer_16|ERROR SUMMARY|ALL_OK|SOME_FAILED" gpurun_out/san_memcheck_layer.log | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/gpu_layer_check.py check_layer check_layer_small > gpurun_out/san_racecheck_layer.log 2>&1; grep -E "layer_16|RACECHECK SUMMARY|ERROR SUMMARY|ALL_OK|SOME_FAILED" gpurun_out/san_racecheck_layer.log | cut -c1-200
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python tools/gpu_layer_check.py check_layer check_layer_small > gpurun_out/san_synccheck_layer.log 2>&1; grep -E "layer_16|ERROR SUMMARY|ALL_OK|SOME_FAILED" gpurun_out/san_synccheck_layer.log | cut -c1-200
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_small_check.py > gpurun_out/san_memcheck_small.log 2>&1; grep -E "ERROR SUMMARY|ALL_OK|SOME_FAILED" gpurun_out/san_memcheck_small.log
unset CUDA_VISIBLE_DEVICES
timeout 900 compute-sanitizer --tool memcheck --target-processes all --print-limit 20 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/multi_gpu_check.py --force-shadow > gpurun_out/san_memcheck_2gpu.log 2>&1; grep -E "MULTI_GPU|ERROR SUMMARY" gpurun_out/san_memcheck_2gpu.log | head
