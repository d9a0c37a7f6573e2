#!/usr/bin/env bash
# The one offline install of the UNMODIFIED reference (build container only: /root/reference is not on the GPU box).
# Output: baseline/_ref/ (git-ignored, travels to the GPU box with gpurun) with the importable packages
#   nicr_mt_scene_analysis, nicr_scene_analysis_datasets, emsanet   and the scripts main.py / inference_*.py as modules.
# The reference's two libraries carry their own pyproject.toml; its top level (emsanet/ + scripts) has none, so it is
# installed from a copy under /tmp to which a 12-line packaging file is added (the sources themselves are untouched).
# Nothing from baseline/_ref is imported by emsanet_b200/ — it is what `python -m emsanet_b200.run main.py ...`,
# `bench.py --impl reference` (kind "reference") and tests/test_run_reference_gpu.py drive.
set -euo pipefail
REF=${1:-/root/reference}
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
DST="$ROOT/baseline/_ref"
TMP=/tmp/emsanet_ref_install
rm -rf "$TMP" "$DST"
mkdir -p "$TMP" "$DST"
cp -r "$REF/lib/nicr-multitask-scene-analysis" "$REF/lib/nicr-scene-analysis-datasets" "$TMP/"
mkdir -p "$TMP/top"
cp -r "$REF/emsanet" "$REF"/main.py "$REF"/inference_*.py "$TMP/top/"
cat > "$TMP/top/pyproject.toml" <<'PY'
[build-system]
build-backend = "setuptools.build_meta"
requires = ["setuptools>=61.0"]
[project]
name = "emsanet-reference"
version = "0"
[tool.setuptools]
packages = ["emsanet"]
py-modules = ["main", "inference_samples", "inference_dataset", "inference_time_whole_model"]
PY
for pkg in "$TMP/nicr-scene-analysis-datasets" "$TMP/nicr-multitask-scene-analysis" "$TMP/top"; do
  python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$DST" "$pkg"
done
rm -rf "$TMP"
echo "installed: $(ls "$DST" | tr '\n' ' ')"
