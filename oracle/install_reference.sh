#!/bin/sh
# Installs the UNMODIFIED reference sources into baseline/_ref (git-ignored; travels to the GPU box) with pip, offline.
# The reference's pyproject.toml lists only the top-level package (`include = ["kvae"]`: the wheel would contain
# kvae/__init__.py and nothing else, its sub-packages have no __init__.py), so the COPY under /tmp that pip builds from
# gets `include = ["kvae", "kvae.*"]`; no source file is touched.  /root/reference itself is read-only.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
rm -rf /tmp/kvae_refcopy && cp -r "$SRC" /tmp/kvae_refcopy
sed -i 's/include = \["kvae"\]/include = ["kvae", "kvae.*"]/' /tmp/kvae_refcopy/pyproject.toml
rm -rf "$ROOT/baseline/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$ROOT/baseline/_ref" /tmp/kvae_refcopy
diff -r "$SRC/kvae" "$ROOT/baseline/_ref/kvae" | grep -v "__pycache__\|Only in $SRC" && echo "installed copy differs from the reference" && exit 1
echo "baseline/_ref ready"
