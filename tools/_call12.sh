set -x
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
export CUDA_VISIBLE_DEVICES=0
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err; python -c "
import json;d=json.loads(open('gpurun_out/final_n1.json').read().strip().splitlines()[-1])
print('N1', d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['roofline'], d.get('vs_nccl_baseline'), d['parity']['ok'], d.get('saturated',{}).get('ms_per_step'), d.get('config5_failure01',{}).get('ms_per_step'), d.get('extras_timed_out_after_s'))"; tail -3 gpurun_out/final_n1.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_ref_n1.json 2> gpurun_out/final_ref_n1.err; python -c "
import json;d=json.loads(open('gpurun_out/final_ref_n1.json').read().strip().splitlines()[-1]);print('REF', d.get('value'), d.get('ms_per_step'), d.get('e2e'), d.get('config',{}).get('global_batch'))"
ncu --set full --clock-control none --import-source on -k regex:wgrad_adam -c 1 -o gpurun_out/prof_wgrad_adam_dyn python tools/gpu_small_check.py --perf-only > gpurun_out/ncu_wa2.log 2>&1; tail -2 gpurun_out/ncu_wa2.log
ncu --set full --clock-control none --import-source on -k regex:swapab -s 6 -c 2 -o gpurun_out/prof_swapab_dyn python tools/gpu_small_check.py --perf-only > gpurun_out/ncu_sab2.log 2>&1; tail -2 gpurun_out/ncu_sab2.log
ncu --set full --clock-control none --import-source on -k regex:attention_bwd -c 1 -o gpurun_out/prof_attention_bwd python -c "
import sys; sys.path.insert(0,'.')
from tools import gpu_attention_check as A
A.check_attention_bwd()" > gpurun_out/ncu_attb.log 2>&1; tail -2 gpurun_out/ncu_attb.log
timeout 300 python tools/gpu_layer_check.py > gpurun_out/layer_check.log 2>&1; tail -1 gpurun_out/layer_check.log
timeout 600 python tools/gpu_attention_check.py > gpurun_out/native_layers_check.log 2>&1; tail -1 gpurun_out/native_layers_check.log
timeout 300 python tools/gpu_small_check.py --perf > gpurun_out/small_check.log 2>&1; tail -1 gpurun_out/small_check.log
