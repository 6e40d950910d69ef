set -x
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
Q="--steps 20 --warmup 5 --no-saturated --no-parity --no-nccl-baseline --no-extra-configs"
for C in 124 132 140; do
  LAH_OPTIMIZER_CTAS=$C CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --gpus 1 $Q > gpurun_out/ov_n1_$C.json 2> gpurun_out/ov_n1_$C.err
  python -c "
import json,sys;d=json.loads(open('gpurun_out/ov_n1_$C.json').read().strip().splitlines()[-1]);print('N1 ctas=$C', round(d['ms_per_step'],3), round(d['value']), d['loss_first_last'])" || tail -3 gpurun_out/ov_n1_$C.err
done
for C in 0 96 124 136; do
  LAH_OPTIMIZER_CTAS=$C timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 $Q > gpurun_out/ov_n4_$C.json 2> gpurun_out/ov_n4_$C.err
  python -c "
import json,sys;d=json.loads(open('gpurun_out/ov_n4_$C.json').read().strip().splitlines()[-1]);print('N4 ctas=$C', round(d['ms_per_step'],3), round(d['value']), d['loss_first_last'], d['exposed_comm_wait_ms_per_step'], d['config']['expert_path'][:5])" || tail -3 gpurun_out/ov_n4_$C.err
done
