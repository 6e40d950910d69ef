set -x
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/final_n8.json 2> gpurun_out/final_n8.err; python -c "
import json;d=json.loads(open('gpurun_out/final_n8.json').read().strip().splitlines()[-1])
print('N8', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['exposed_comm_wait_ms_per_step'], d.get('vs_nccl_baseline'), d['parity']['ok'], d.get('saturated',{}).get('ms_per_step'), d.get('saturated',{}).get('value'), d.get('config5_failure01',{}).get('ms_per_step'), d.get('extras_timed_out_after_s'), d['config']['expert_path'][:5])"; tail -3 gpurun_out/final_n8.err
Q="--steps 20 --warmup 5 --no-saturated --no-parity --no-nccl-baseline --no-extra-configs"
for C in 64 120; do
  LAH_OPTIMIZER_CTAS=$C timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 $Q > gpurun_out/ov_n8_$C.json 2> gpurun_out/ov_n8_$C.err
  python -c "
import json,sys;d=json.loads(open('gpurun_out/ov_n8_$C.json').read().strip().splitlines()[-1]);print('N8 ctas=$C', round(d['ms_per_step'],3), round(d['value']), d['exposed_comm_wait_ms_per_step'])" || tail -3 gpurun_out/ov_n8_$C.err
done
