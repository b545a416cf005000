#!/bin/sh
# Copy the UNMODIFIED reference hot-path sources into the git-ignored baseline/_ref/ (it travels to the GPU box with gpurun; the
# reference has no setup.py, so there is nothing to pip-install).  Never committed.
set -e
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
mkdir -p "$ROOT/baseline/_ref"
cp -r /root/reference/quant /root/reference/utils "$ROOT/baseline/_ref/"
find "$ROOT/baseline/_ref" -name __pycache__ -prune -exec rm -rf {} + 2>/dev/null || true
echo "reference copied to $ROOT/baseline/_ref (git-ignored)"
