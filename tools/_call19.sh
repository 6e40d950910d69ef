set -x
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu.py -x -q -m gpu -k "graph or kernels_and_fused or smoke or baseline_trainer or update_every or microbatch or public_api or native_library" 2>&1 | tail -4
Q="--steps 8 --warmup 3 --no-saturated --no-parity --no-nccl-baseline --no-extra-configs --batch-per-gpu 65536 --expert-path big"
for V in 0 1; do
LAH_BIG_ADAM_OVERLAP=$V timeout 300 python bench.py $Q > gpurun_out/sat_ov$V.json 2> gpurun_out/sat_ov$V.err
python -c "
import json;d=json.loads(open('gpurun_out/sat_ov$V.json').read().strip().splitlines()[-1]);print('SAT overlap=$V', round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['ms_per_step'],3), d['clocks'], d['stage_ms_rank0'])" || tail -5 gpurun_out/sat_ov$V.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-saturated --no-nccl-baseline --no-extra-configs > gpurun_out/n1_19.json 2> gpurun_out/n1_19.err
python -c "
import json;d=json.loads(open('gpurun_out/n1_19.json').read().strip().splitlines()[-1]);print('N1', round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['ms_per_step'],3), d['parity']['ok'], d['clocks'])" || tail -5 gpurun_out/n1_19.err
