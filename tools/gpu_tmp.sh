#!/bin/bash
for cfg in "QFB_PLAN_SINK=0,QFB_PLAN_TABLES=0,QFB_PLAN_CARRY=0,QFB_PLAN_ALLIN=0" "QFB_PLAN_ALLIN=0"; do
echo "== $cfg"; env $(echo $cfg | tr ',' ' ') timeout 600 python -m pytest tests/test_gpu_states.py -m gpu -q -p no:cacheprovider -k sampling_is_bit_exact 2>&1 | tail -n 1
done
git stash -q; git checkout -q 8450c77 -- quantumflow_b200 tests/plan_emulator.py 2>/dev/null; python -c "
from quantumflow_b200 import _build; _build.build_library()"; echo "== round-1 tree"; timeout 600 python -m pytest tests/test_gpu_states.py -m gpu -q -p no:cacheprovider -k sampling_is_bit_exact 2>&1 | tail -n 1
