# 8-GPU: data-parallel parity check + paired bench with the peer exchange and with NCCL + the same box's 1-GPU bench
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d[k] for k in ('ok','vs_one_process_accumulating_the_same_shards') if k in d} or (d['value'],d['ms_per_step'],d['e2e']['value'],d.get('gradient_exchange')))" $1; }
FPG_DDP=peer timeout 300 $TR --master-port 29611 bench.py --gpus $N --check --model pairedattention --batch $((2*N)) --check_size 128 > gpurun_out/r2_dpcheck_peer_paired_${N}gpu.json 2> gpurun_out/r2_dpcheck_peer_paired_${N}gpu.err; echo "check rc=$?"; show gpurun_out/r2_dpcheck_peer_paired_${N}gpu.json || tail -20 gpurun_out/r2_dpcheck_peer_paired_${N}gpu.err
FPG_DDP=peer timeout 300 $TR --master-port 29612 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2_bench_peer_${N}gpu.json 2> gpurun_out/r2_bench_peer_${N}gpu.err; echo "bench peer rc=$?"; show gpurun_out/r2_bench_peer_${N}gpu.json
FPG_DDP=nccl timeout 300 $TR --master-port 29613 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2_bench_nccl_${N}gpu.json 2> gpurun_out/r2_bench_nccl_${N}gpu.err; echo "bench nccl rc=$?"; show gpurun_out/r2_bench_nccl_${N}gpu.json
timeout 300 python bench.py --steps 30 --warmup 5 --no_unet > gpurun_out/r2_bench_same_box_1gpu.json 2> gpurun_out/r2_bench_same_box_1gpu.err; echo "bench 1 rc=$?"; show gpurun_out/r2_bench_same_box_1gpu.json
