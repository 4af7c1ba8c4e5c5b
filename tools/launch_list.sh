# ncu launch list (gpu__time_duration per launch) of one replayed PairedAttention step; $1 = output tag, env passes through
TAG=${1:-x}
python tools/one_step.py > gpurun_out/one_step_$TAG.log 2>&1 || { tail -5 gpurun_out/one_step_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$TAG.csv python tools/one_step.py > gpurun_out/ncu_$TAG.log 2>&1
python tools/agg_launches.py gpurun_out/launches_$TAG.csv | head -${2:-45}
