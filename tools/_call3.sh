set -x
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
export CUDA_VISIBLE_DEVICES=0
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_full_n1.json 2> gpurun_out/r2_full_n1.err; tail -c 600 gpurun_out/r2_full_n1.json; tail -5 gpurun_out/r2_full_n1.err
ncu --set full --clock-control none --import-source on -k regex:wgrad_adam -c 1 -o gpurun_out/prof_wgrad_adam python tools/gpu_small_check.py --perf-only > gpurun_out/ncu_wa.log 2>&1; tail -3 gpurun_out/ncu_wa.log
ncu --set full --clock-control none --import-source on -k regex:swapab -s 4 -c 1 -o gpurun_out/prof_swapab python tools/gpu_small_check.py --perf-only > gpurun_out/ncu_sab.log 2>&1; tail -3 gpurun_out/ncu_sab.log
unset CUDA_VISIBLE_DEVICES
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-saturated > gpurun_out/r2_named_n2.json 2> gpurun_out/r2_named_n2.err; tail -c 1500 gpurun_out/r2_named_n2.json; tail -8 gpurun_out/r2_named_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/multi_gpu_check.py --force-shadow 2>&1 | tail -4
