set -x
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for B in 32768 65536; do
for V in 0 1; do
Q="--steps 8 --warmup 3 --no-saturated --no-nccl-baseline --no-extra-configs --batch-per-gpu $B --expert-path big"
LAH_BIG_ADAM_OVERLAP=$V timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 $Q > gpurun_out/sat2_ov$V.json 2> gpurun_out/sat2_ov$V.err
python -c "
import json;d=json.loads(open('gpurun_out/sat2_ov$V.json').read().strip().splitlines()[-1]);print('N2 B=$B overlap=$V', round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['ms_per_step'],3), d['exposed_comm_wait_ms_per_rank'], d['parity']['ok'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -5 gpurun_out/sat2_ov$V.err
done
done
timeout 600 python -m pytest tests/test_gpu.py -x -q -m gpu -k "overlap or two_gpu or rank_failure" 2>&1 | tail -3
