#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/q4_tests.log
for rep in 1 2; do
python bench.py --steps 1000 --warmup 20 --no-cpu > $O/q4_keep_$rep.json 2> $O/q4_keep_$rep.err
MDX_LIB_PATH=$PWD/scratch/libmdx_nokeep.so python bench.py --steps 1000 --warmup 20 --no-cpu > $O/q4_nokeep_$rep.json 2> $O/q4_nokeep_$rep.err
done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:mdx_env_kernel -s 100 -c 3 --csv --log-file $O/q4_keep_dram.csv python bench.py --steps 120 --warmup 3 --no-cpu > /dev/null 2>&1
cat $O/q4_tests.log; python - <<PY
import json
for wl in ("keep_1","nokeep_1","keep_2","nokeep_2"):
    try:
        d=json.load(open("$O/q4_%s.json"%wl)); print(wl, round(d["value"]/1e6,2), "M frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]/1e6,2))
    except Exception as ex: print(wl, "failed", ex, open("$O/q4_%s.err"%wl).read()[-800:])
PY
grep -E "dram__bytes|duration|hit_rate" $O/q4_keep_dram.csv | cut -d, -f5,13- | head -12
