#!/bin/bash
# Snapshot of the UNMODIFIED reference for end-to-end drop-in checks on the GPU box (tests/test_gpu_reference_scripts.py).
# The reference has no setup.py / pyproject.toml, so "install" = copy the tree as it is.  baseline/_ref is git-ignored (never part
# of the history) but travels to the GPU box with gpurun.  Run in the build container, where /root/reference exists.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
rm -rf "$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline/_ref"
cp -r "$SRC"/. "$ROOT/baseline/_ref/"
find "$ROOT/baseline/_ref" -name "__pycache__" -type d -exec rm -rf {} +
echo "reference snapshot: $(find "$ROOT/baseline/_ref" -type f | wc -l) files under baseline/_ref"
