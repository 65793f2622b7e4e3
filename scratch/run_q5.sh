#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m pytest tests/test_gpu_next.py -x -q 2>&1 | tail -5 > $O/q5_tests.log
for wl in idq greedy; do
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 8 --csv --log-file $O/q5_launch_$wl.csv python bench.py --workload $wl --steps 30 --warmup 3 --no-cpu > /dev/null 2>&1
done
cat $O/q5_tests.log
for wl in idq greedy; do grep -E "duration|dram__bytes" $O/q5_launch_$wl.csv | cut -d, -f5,13- | cut -c1-60,100- | head -12; done
