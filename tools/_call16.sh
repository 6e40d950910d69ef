set -x
N=$1
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
S=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/final_n$N.json 2> gpurun_out/final_n$N.err
echo "wall $(( $(date +%s) - S )) s"
python -c "
import json;d=json.loads(open('gpurun_out/final_n$N.json').read().strip().splitlines()[-1])
print('N$N', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['gpu_launches'], d['exposed_comm_wait_ms_per_rank'], d['hottest_expert_rows_per_layer'], d.get('vs_nccl_baseline'), d['parity']['ok'], d.get('saturated',{}).get('ms_per_step'), d.get('strong_scaling_saturated'), d.get('config5_failure01',{}).get('ms_per_step'), d.get('extras_timed_out_after_s'))"; tail -3 gpurun_out/final_n$N.err
