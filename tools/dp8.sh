# 8-GPU data-parallel parity check + bench, peer exchange vs NCCL (run under gpurun --gpus 8)
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29611 bench.py --gpus $N --check --model pairedattention --batch $((2*N)) > gpurun_out/r2_dpcheck_peer_paired_${N}gpu.json 2> gpurun_out/r2_dpcheck_peer_paired_${N}gpu.err; echo "check rc=$?"; head -c 700 gpurun_out/r2_dpcheck_peer_paired_${N}gpu.json; echo
timeout 600 $TR --master-port 29612 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2_bench_peer_${N}gpu.json 2> gpurun_out/r2_bench_peer_${N}gpu.err; echo "bench peer rc=$?"; head -c 330 gpurun_out/r2_bench_peer_${N}gpu.json; echo
FPG_DDP=nccl timeout 600 $TR --master-port 29613 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2_bench_nccl_${N}gpu.json 2> gpurun_out/r2_bench_nccl_${N}gpu.err; echo "bench nccl rc=$?"; head -c 330 gpurun_out/r2_bench_nccl_${N}gpu.json; echo
timeout 600 python bench.py --steps 30 --warmup 5 --no_unet > gpurun_out/r2_bench_same_box_1gpu.json 2> gpurun_out/r2_bench_same_box_1gpu.err; echo "bench 1 rc=$?"; head -c 330 gpurun_out/r2_bench_same_box_1gpu.json; echo
timeout 600 $TR --master-port 29614 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2_bench_peer_${N}gpu_b.json 2> gpurun_out/r2_bench_peer_${N}gpu_b.err; echo "bench peer rc=$?"; head -c 330 gpurun_out/r2_bench_peer_${N}gpu_b.json; echo
