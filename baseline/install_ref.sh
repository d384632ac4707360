#!/bin/bash
# Installs the UNMODIFIED reference package `npf` (pure Python) into baseline/_ref/ (git-ignored; shipped to the GPU box by
# gpurun) so that `bench.py --impl reference` can time the reference's own CPU path.  The reference ships no
# setup.py/pyproject, so the install runs from a scratch copy under /tmp that gets a 6-line setup.py (the only added file;
# nothing of the reference is edited).  No reference source enters the git history.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
[ -d "$REF/npf" ] || { echo "install_ref: $REF/npf not found (nothing to do)"; exit 0; }
TMP="$(mktemp -d /tmp/npf_ref_XXXX)"
cp -r "$REF/npf" "$TMP/npf"
cat > "$TMP/setup.py" <<'PY'
from setuptools import setup, find_packages
setup(name="npf", version="0.0.0+reference", packages=find_packages(include=["npf", "npf.*"]))
PY
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$TMP" -q
rm -rf "$TMP"
python - <<PY
import sys; sys.path.insert(0, "$HERE/_ref")
import npf; print("installed reference npf at", npf.__file__)
PY
