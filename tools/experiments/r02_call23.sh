#!/bin/bash
# 2 GPUs: parity of the sharded path with the linked run lists, config 5 with the scaled halo push, the 38-plane shards
# (what each of 8 GPUs holds) with both dir_spmv kernels and a timeline
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 tools/shard_check.py --parity-only > $O/r02_shard_check_2gpu.json 2> $O/r02_shard_check_2gpu.err; echo "shard_check 2 rc=$?"; tail -c 300 $O/r02_shard_check_2gpu.json; grep -i "error\|assert\|Traceback\|marching" $O/r02_shard_check_2gpu.err | head -8
timeout 900 $TR --master-port 29522 bench.py --workload c5 --gpus 2 --steps 3 --warmup 3 > $O/r02_bench_c5_n2.json 2> $O/r02_bench_c5_n2.err; echo "bench c5 n2 rc=$?"; cut -c1-160 $O/r02_bench_c5_n2.json; tail -2 $O/r02_bench_c5_n2.err | cut -c1-300
for opt in "march=1" "march=2"; do
timeout 900 $TR --master-port 29523 bench.py --workload c4slab4 --gpus 2 --steps 5 --warmup 3 --opt trace=256 --opt $opt > $O/r02_bench_slab4_$opt.json 2> $O/r02_bench_slab4_$opt.err; echo "bench slab4 $opt rc=$?"; cut -c1-120 $O/r02_bench_slab4_$opt.json
python tools/trace_report.py $O/trace_c4slab4_n2_r*.npy | tee $O/r02_trace_slab4_n2_$opt.txt
done
