set -x
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-saturated > gpurun_out/r2_named_n8.json 2> gpurun_out/r2_named_n8.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_named_n8.json').read().strip().splitlines()[-1])
print(d['value'],d['ms_per_step'],d['roofline'],d['e2e']['ms_per_step'],d['exposed_comm_wait_ms_per_step'],d['parity'],d.get('vs_nccl_baseline'))
print(d['stage_ms_rank0']); print({k:(v.get('value'),v.get('ms_per_step'),v.get('error')) for k,v in d.items() if k.startswith('config')}); print(d['nccl_baseline'].get('ms_per_step'))"; tail -5 gpurun_out/r2_named_n8.err
