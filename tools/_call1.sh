set -x
nvidia-smi -L
nvidia-smi topo -m | head -20
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
CUDA_VISIBLE_DEVICES=0 python bench.py --gpus 1 --steps 20 --warmup 5 --batch-per-gpu 256 > gpurun_out/r2_b256_n1_old.json 2> gpurun_out/r2_b256_n1_old.err
tail -c 3000 gpurun_out/r2_b256_n1_old.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/symm_probe.py > gpurun_out/symm_probe.log 2>&1
tail -5 gpurun_out/symm_probe.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --batch-per-gpu 256 > gpurun_out/r2_b256_n2_old.json 2> gpurun_out/r2_b256_n2_old.err
tail -c 2500 gpurun_out/r2_b256_n2_old.json
CUDA_VISIBLE_DEVICES=0 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 --ref-batch 256 > gpurun_out/r2_ref_b256.json 2>&1
tail -c 1500 gpurun_out/r2_ref_b256.json
