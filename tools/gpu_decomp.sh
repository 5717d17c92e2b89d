#!/bin/bash
mkdir -p gpurun_out
python tools/decompose.py 30 12 3 2>&1 | tee gpurun_out/decompose_12_3.jsonl
