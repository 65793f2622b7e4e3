#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
nvidia-smi -L | head -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 600 --warmup 20 > $O/bench_8gpu_r1j.json 2> $O/bench_8gpu_r1j.err
tail -c 1500 $O/bench_8gpu_r1j.json | cut -c1-600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload c5 --steps 200 --warmup 5 > $O/bench_8gpu_r1j_c5.json 2> $O/bench_8gpu_r1j_c5.err
tail -c 1500 $O/bench_8gpu_r1j_c5.json | cut -c1-600
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --impl reference --steps 3 --warmup 3 > $O/bench_8gpu_r1j_ref.json 2> $O/bench_8gpu_r1j_ref.err
cut -c1-300 $O/bench_8gpu_r1j_ref.json
