set -x
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/fault_check.py 2>&1 | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/fault_check.py --graph 2>&1 | tail -4
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-saturated --no-nccl-baseline > gpurun_out/r2_named_n2_dyn.json 2> gpurun_out/r2_named_n2_dyn.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_named_n2_dyn.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline'],d['e2e']['ms_per_step'],d['exposed_comm_wait_ms_per_step'],d['parity']['ok'])"; tail -3 gpurun_out/r2_named_n2_dyn.err
export CUDA_VISIBLE_DEVICES=0
LAH_CUDA_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_named.csv python bench.py --gpus 1 --steps 2 --warmup 3 --no-saturated --no-parity --no-nccl-baseline --no-e2e --no-graph > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches_named.csv
