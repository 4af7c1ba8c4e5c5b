python tools/micro_peer.py 2>&1 | tee gpurun_out/micro_peer.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_peer_2gpu.json 2> gpurun_out/r2_bench_peer_2gpu.err; echo "bench peer rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench_peer_2gpu.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['gradient_exchange'])"
