set -x
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/bench_sweep.py --gpus 8 --steps 20 --warmup 5 --points small:0 small:32 small:64 small:96 big > gpurun_out/sweep8.log 2> gpurun_out/sweep8.err
grep '^{' gpurun_out/sweep8.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d.get('point'), d.get('ms_per_step'), d.get('samples_per_s'), d.get('exposed_wait_ms_per_rank'), d.get('hottest_expert_rows_per_layer'), d.get('graph'), d.get('error'))"
tail -3 gpurun_out/sweep8.err
