set -x
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
export CUDA_VISIBLE_DEVICES=0
timeout 900 python - <<'PY' 2>&1 | tail -30
import sys; sys.path.insert(0, '.')
from tools import gpu_attention_check as A
for fn in (A.check_attention, A.check_attention_bwd, A.check_transformer_train):
    try:
        fn()
    except Exception as e:
        import traceback; traceback.print_exc()
print({k: v.get('ok') for k, v in A.results.items()})
PY
timeout 600 compute-sanitizer --tool synccheck --print-limit 5 python tools/gpu_small_check.py > gpurun_out/san_synccheck_small.log 2>&1; grep -E "ERROR SUMMARY|ALL_OK|SOME_FAILED|Barrier error" gpurun_out/san_synccheck_small.log | sort | uniq -c | head
timeout 600 python tools/gpu_small_check.py 2>&1 | grep -E "ALL_OK|SOME_FAILED"
