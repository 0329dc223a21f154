#!/bin/bash
# Installs the UNMODIFIED reference (AnswerDotAI/cold-compress, mounted read-only at /root/reference) into
# baseline/_ref so that `bench.py --impl reference` can run its own CPU path on the GPU box (baseline/_ref is
# git-ignored but travels with the gpurun snapshot).  The reference's setup.py calls find_packages() on a flat
# directory of modules, so a plain `pip install` produces an empty wheel; the copy under /tmp gets a setup.py that
# lists the top-level modules (py_modules) -- the .py files themselves are installed byte for byte.
set -e
SRC=${1:-/root/reference}
DST=$(cd "$(dirname "$0")/.." && pwd)/baseline/_ref
TMP=$(mktemp -d)
cp -r "$SRC"/. "$TMP"/
cat > "$TMP"/setup.py <<'PY'
from setuptools import setup
import glob, os
mods = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob("*.py") if os.path.basename(p) != "setup.py")
setup(name="cold-compress-reference", version="0.1", py_modules=mods,
      data_files=[("cache_configs", glob.glob("cache_configs/*"))])
PY
rm -rf "$DST"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$DST" "$TMP" 2>&1 | tail -2
# the YAML presets are looked up next to generation_utils.py (generation_utils.py:79-91)
mkdir -p "$DST"/cache_configs && cp "$SRC"/cache_configs/* "$DST"/cache_configs/
for f in cache.py model.py generation_utils.py attention_utils.py prompt_compression.py tp.py; do cmp "$SRC/$f" "$DST/$f"; done
echo "installed $(ls "$DST"/*.py | wc -l) reference modules into $DST"
rm -rf "$TMP"
