#!/bin/bash
# Installs the UNMODIFIED reference into baseline/_ref (git-ignored; it travels to the GPU box with the snapshot).
# The reference's own setup.py names no packages, so a plain `pip install --target` flattens src/* into top-level
# `models/ masks/ utils/ datasets/` -- unusable, because every reference module imports `src.…` / `app.…`.
# We therefore install from a /tmp copy whose ONLY change is the packaging stanza (setup.py lists the
# namespace packages src*, app*, evals*); no reference .py file under those packages is touched.
set -euo pipefail
REF=${1:-/root/reference}
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
TMP=$(mktemp -d)
cp -r "$REF" "$TMP/ref"
cat > "$TMP/ref/setup.py" <<'PY'
from setuptools import setup, find_namespace_packages
setup(name="jepa", version="0.0.1", description="JEPA research code (packaging shim: namespace packages listed)",
      packages=find_namespace_packages(include=["src", "src.*", "app", "app.*", "evals", "evals.*"]))
PY
rm -rf "$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline"
cd "$TMP/ref"
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target "$ROOT/baseline/_ref" . 2>&1 | tail -3
rm -rf "$TMP"
ls "$ROOT/baseline/_ref"
