# 4-GPU paired bench: default reducer policy (peer exchange) vs NCCL
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print((d['value'],d['ms_per_step'],d['e2e']['value'],d.get('gradient_exchange')))" $1; }
timeout 300 $TR --master-port 29612 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/r2_bench_peer_4gpu.json 2> gpurun_out/r2_bench_peer_4gpu.err; echo "bench default rc=$?"; show gpurun_out/r2_bench_peer_4gpu.json || tail -20 gpurun_out/r2_bench_peer_4gpu.err
FPG_DDP=nccl timeout 300 $TR --master-port 29613 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/r2_bench_nccl_4gpu.json 2> gpurun_out/r2_bench_nccl_4gpu.err; echo "bench nccl rc=$?"; show gpurun_out/r2_bench_nccl_4gpu.json
