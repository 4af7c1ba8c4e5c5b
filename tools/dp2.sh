# 2-GPU data-parallel parity checks + benches (run under gpurun --gpus 2)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d[k] for k in ('ok','vs_one_process_accumulating_the_same_shards') if k in d} or (d['value'],d['ms_per_step'],d['e2e']['value'],d.get('gradient_exchange')))" $1; }
timeout 600 $TR --master-port 29611 bench.py --gpus 2 --check --model pairedattention --batch 16 > gpurun_out/r2_dpcheck_peer_paired_2gpu.json 2> gpurun_out/r2_dpcheck_peer_paired_2gpu.err; echo "check rc=$?"; show gpurun_out/r2_dpcheck_peer_paired_2gpu.json || tail -20 gpurun_out/r2_dpcheck_peer_paired_2gpu.err
timeout 600 $TR --master-port 29615 bench.py --gpus 2 --check --model attentiongan --identity --batch 4 --check_size 128 > gpurun_out/r2_dpcheck_attgan_2gpu.json 2> gpurun_out/r2_dpcheck_attgan_2gpu.err; echo "check rc=$?"; show gpurun_out/r2_dpcheck_attgan_2gpu.json || tail -20 gpurun_out/r2_dpcheck_attgan_2gpu.err
timeout 600 $TR --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 --model pix2pix > gpurun_out/r2_bench_pix2pix_2gpu.json 2> gpurun_out/r2_bench_pix2pix_2gpu.err; echo "bench pix2pix rc=$?"; show gpurun_out/r2_bench_pix2pix_2gpu.json || tail -20 gpurun_out/r2_bench_pix2pix_2gpu.err
timeout 600 $TR --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_peer_2gpu.json 2> gpurun_out/r2_bench_peer_2gpu.err; echo "bench paired rc=$?"; show gpurun_out/r2_bench_peer_2gpu.json || tail -20 gpurun_out/r2_bench_peer_2gpu.err
