#!/usr/bin/env bash
# Installs the UNMODIFIED reference (mryab/learning-at-home) into baseline/_ref for `bench.py --impl reference`.
# The reference tree has no setup.py / pyproject.toml, so `pip install /root/reference` fails with
#   "Neither 'setup.py' nor 'pyproject.toml' found".
# We therefore install from a copy under /tmp that adds packaging metadata only (setup.py + empty __init__.py files for
# the experiments/ directories); no reference source file is changed (verified with diff -r below).
set -euo pipefail
REPO="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
REF="${1:-/root/reference}"
TMP="$(mktemp -d /tmp/refcopy.XXXXXX)"
cp -r "$REF"/. "$TMP"/
cat > "$TMP/setup.py" <<'PY'
from setuptools import setup
setup(name="learning_at_home_reference", version="0.0.0",
      packages=["lib", "lib.client", "lib.network", "lib.runtime", "lib.server", "lib.utils",
                "experiments", "experiments.convergence", "experiments.throughput"])
PY
touch "$TMP/experiments/__init__.py" "$TMP/experiments/convergence/__init__.py" "$TMP/experiments/throughput/__init__.py"
rm -rf "$REPO/baseline/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$REPO/baseline/_ref" "$TMP" 2>&1 | tail -2
diff -r -x __pycache__ "$REF/lib" "$REPO/baseline/_ref/lib" && echo "lib/: identical to reference"
for f in convergence/dmoe_emulator.py convergence/faulty_dmoe_emulator.py throughput/layers.py; do
  diff "$REF/experiments/$f" "$REPO/baseline/_ref/experiments/$f" && echo "experiments/$f: identical"
done
rm -rf "$TMP"
