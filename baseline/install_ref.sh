#!/bin/bash
# Put the UNMODIFIED reference package under baseline/_ref/ (git-ignored, travels to the GPU box).
# `pip install --no-index --no-deps --target baseline/_ref /root/reference` fails in metadata
# generation: the reference's setup.py lists "tqdm~=4.67.1" "lightning>=2.3.0" without a comma
# (one malformed requirement string).  The package is pure Python, so what pip would have installed
# is the hbird/ directory itself: copy it verbatim.  Used by tests/test_gpu_reference_engine.py
# (the reference's own engine driving the b200 plugin).  Nothing in the product reads it.
set -e
REF="${HBIRD_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
[ -d "$REF/hbird" ] || { echo "no reference tree at $REF"; exit 0; }
rm -rf "$HERE/_ref"
mkdir -p "$HERE/_ref"
cp -r "$REF/hbird" "$HERE/_ref/hbird"
find "$HERE/_ref" -name "__pycache__" -type d -prune -exec rm -rf {} +
echo "reference package copied to $HERE/_ref/hbird"
