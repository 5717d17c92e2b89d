#!/bin/bash
# Build container only: copy the reference's package and tests into the git-ignored baseline/_ref/ so that
# run_conformance.sh can run them on the GPU box (the reference is never committed; remove with --clean).
REPO=$(cd "$(dirname "$0")/../.." && pwd)
if [ "$1" == "--clean" ]; then rm -rf "$REPO/baseline/_ref/reference"; exit 0; fi
mkdir -p "$REPO/baseline/_ref/reference"
cp -r /root/reference/quantumflow /root/reference/tests "$REPO/baseline/_ref/reference/"
find "$REPO/baseline/_ref/reference" -name "__pycache__" -prune -exec rm -rf {} \; 2>/dev/null
echo staged: $(du -sh "$REPO/baseline/_ref/reference" | cut -f1)
