set -x
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12
CUDA_VISIBLE_DEVICES=0 timeout 300 python - <<'PY' 2>&1 | grep -E "attention_bwd|transformer_train"
import sys; sys.path.insert(0, '.')
from tools import gpu_attention_check as A
A.check_attention_bwd(); A.check_transformer_train()
PY
