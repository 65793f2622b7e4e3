#!/bin/bash
O=gpurun_out
T=r2hh
MDX_LIB_PATH=$PWD/scratch/ab/libmdx_hoist4.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py -x -q -m gpu > $O/${T}_tests.log 2>&1; echo "tests(hoist4 lib) rc=$?"; tail -2 $O/${T}_tests.log
MDX_JIT=1 MDX_HOIST=4 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/${T}_tests_jit.log 2>&1; echo "tests(jit hoist4) rc=$?"; tail -2 $O/${T}_tests_jit.log
run() { # name, env..., workload
  local name=$1; shift; local wl=$1; shift
  env "$@" timeout 300 python bench.py --workload $wl --steps 60 --warmup 5 --no-cpu --no-secondary --sustained-episodes 0 --no-feat-host > $O/${T}_$name.json 2> $O/${T}_$name.err
  python - <<PY
import json
d=json.loads(open("$O/${T}_$name.json").read().strip().splitlines()[-1])
print("$name", round(d["value"]/1e6,2), "ms", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), "single", round(d["roofline"]["single_launch"]["frac"],4))
PY
}
for rep in 1 2; do
run c3_base c3 X=1
run c3_hoist c3 MDX_LIB_PATH=$PWD/scratch/ab/libmdx_hoist2.so
done
run k100_base c3k100 MDX_HOIST=0
run k100_hoist c3k100 MDX_HOIST=2
run c4_base c4 MDX_HOIST=0
run c4_hoist c4 MDX_HOIST=4
run c5_base c5 MDX_HOIST=0
run c5_hoist c5 MDX_HOIST=4
