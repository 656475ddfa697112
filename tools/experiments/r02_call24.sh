#!/bin/bash
# 2 GPUs, config 5: where does the iteration go?  timeline + sensitivity to the number of pushing blocks
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for nb in 0 32 1184; do
CGB200_PUSH_BLOCKS=$nb timeout 900 $TR --master-port 29522 bench.py --workload c5 --gpus 2 --steps 2 --warmup 3 --no-e2e --opt trace=256 > $O/r02_bench_c5_n2_nb$nb.json 2> $O/r02_bench_c5_n2_nb$nb.err; echo "bench c5 n2 nb=$nb rc=$?"; cut -c1-110 $O/r02_bench_c5_n2_nb$nb.json
python tools/trace_report.py $O/trace_c5_n2_r*.npy
done
nvidia-smi topo -m | head -12
