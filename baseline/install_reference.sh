#!/bin/bash
# Installs the UNMODIFIED reference (usnistgov/optbayesexpt v1.2.0, pure Python) into baseline/_ref so that
# `bench.py --impl reference` and the cpu_baseline legs can import it on the GPU box (where /root/reference
# does not exist).  baseline/_ref is git-ignored (never committed) but travels with gpurun.
#   1. pip --target from a scratch copy (the source tree is read-only and setup.py reads README.md from cwd);
#   2. pip fails here because setup.py lists setup_requires=['pytest-runner'], which is not in the offline
#      wheelhouse; the package is pure Python, so the fallback installs exactly what pip would have: a verbatim
#      copy of the `optbayesexpt/` package directory (no file is edited).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
[ -d "$SRC/optbayesexpt" ] || { echo "no reference at $SRC"; exit 0; }
rm -rf /tmp/obe_ref_src "$HERE/_ref"
cp -r "$SRC" /tmp/obe_ref_src
if (cd /tmp/obe_ref_src && python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse \
        --no-deps --target "$HERE/_ref" /tmp/obe_ref_src > /tmp/obe_ref_pip.log 2>&1); then
    echo "pip install into baseline/_ref: ok"
else
    echo "pip install failed (setup_requires pytest-runner unavailable offline): copying the pure-Python package verbatim"
    mkdir -p "$HERE/_ref"
    cp -r "$SRC/optbayesexpt" "$HERE/_ref/optbayesexpt"
    find "$HERE/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
fi
python - <<PY
import sys
sys.path.insert(0, "$HERE/_ref")
import optbayesexpt
print("reference importable:", optbayesexpt.__file__)
PY
