#!/bin/bash
# Runs the reference's own hot-path test files, unmodified, against the engine on a GPU (two ways: the reference's
# classes on the b200 backend plugin, and the mirror package under the reference's name). The reference tree is not
# part of this repository: stage_reference.sh copies it (build container only) into the git-ignored baseline/_ref/,
# which travels to the GPU box with the snapshot. Output: gpurun_out/conformance_*.log
REPO=$(cd "$(dirname "$0")/../.." && pwd)
REF=${QF_REFERENCE_ROOT:-$REPO/baseline/_ref/reference}
mkdir -p "$REPO/gpurun_out"
if [ ! -d "$REF/tests" ]; then echo "no reference tree at $REF (run tools/conformance/stage_reference.sh first)"; exit 1; fi
FILES="tests/test_backend.py tests/test_states.py tests/test_gates.py tests/test_stdgates.py tests/test_circuits.py tests/test_channels.py tests/test_qubits.py tests/test_qaoa.py tests/test_dagcircuit.py tests/test_stdops.py"
cd "$REF"
for mode in backend mirror; do
  QF_REFERENCE_ROOT="$REF" PYTHONPATH="$REPO/tools/conformance:$REPO" timeout 1500 python -m pytest -p qfplug_$mode -p no:cacheprovider -q $FILES \
    > "$REPO/gpurun_out/conformance_$mode.log" 2>&1
  echo "== $mode"; tail -n 4 "$REPO/gpurun_out/conformance_$mode.log"
done
