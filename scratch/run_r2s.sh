#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_alt.py -x -q -m gpu > gpurun_out/r2s_alt.log 2>&1; echo "alt rc=$?"
tail -5 gpurun_out/r2s_alt.log
timeout 300 python scratch/idq_split.py > gpurun_out/r2s_idq_split.json 2> gpurun_out/r2s_idq_split.err; echo "split rc=$?"
cat gpurun_out/r2s_idq_split.json; tail -3 gpurun_out/r2s_idq_split.err
