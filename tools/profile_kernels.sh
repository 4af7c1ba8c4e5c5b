#!/bin/bash
# ncu --set full captures of the dominant kernels on their real layer shapes (one launch each, after the plain run of
# the same command exited 0). Run on a B200: bash tools/profile_kernels.sh ; reports land in gpurun_out/.
set -u
NCU="ncu --set full --clock-control none --import-source on"
RES="16 64 64 256 256 3 1 0 1"
HEAD="16 256 256 64 27 7 1 0 3"
STEM="16 256 256 9 64 7 1 0 3"
python tools/micro_conv.py $RES > gpurun_out/prof_plain_res.log 2>&1 || exit 1
python tools/micro_conv.py $HEAD > gpurun_out/prof_plain_head.log 2>&1 || exit 1
python tools/micro_conv.py $STEM fprop wgrad > gpurun_out/prof_plain_stem.log 2>&1 || exit 1
python tools/micro_in.py > gpurun_out/prof_plain_in.log 2>&1 || exit 1
# micro_conv runs 7 repetitions per kind: launch 3 / 10 / 17 of the igemm kernels = a warm fprop / dgrad / wgrad
$NCU -k regex:igemm -s 3 -c 1 -o gpurun_out/r02_res_fprop python tools/micro_conv.py $RES > /dev/null 2>&1
$NCU -k regex:igemm -s 10 -c 1 -o gpurun_out/r02_res_dgrad python tools/micro_conv.py $RES > /dev/null 2>&1
$NCU -k regex:igemm -s 17 -c 1 -o gpurun_out/r02_res_wgrad python tools/micro_conv.py $RES > /dev/null 2>&1
$NCU -k regex:igemm -s 3 -c 1 -o gpurun_out/r02_head_fprop python tools/micro_conv.py $HEAD > /dev/null 2>&1
$NCU -k regex:igemm -s 10 -c 1 -o gpurun_out/r02_head_dgrad python tools/micro_conv.py $HEAD > /dev/null 2>&1
$NCU -k regex:igemm -s 17 -c 1 -o gpurun_out/r02_head_wgrad python tools/micro_conv.py $HEAD > /dev/null 2>&1
$NCU -k regex:igemm -s 3 -c 1 -o gpurun_out/r02_stem_fprop python tools/micro_conv.py $STEM fprop wgrad > /dev/null 2>&1
$NCU -k regex:igemm -s 10 -c 1 -o gpurun_out/r02_stem_wgrad python tools/micro_conv.py $STEM fprop wgrad > /dev/null 2>&1
# InstanceNorm kernels on the residual-trunk shape: micro_in runs 10 repetitions of stats, apply, bwd(dz2+dres), bwd
$NCU -k regex:"in_stats" -s 3 -c 1 -o gpurun_out/r02_in_stats python tools/micro_in.py > /dev/null 2>&1
$NCU -k regex:"in_apply" -s 3 -c 1 -o gpurun_out/r02_in_apply python tools/micro_in.py > /dev/null 2>&1
$NCU -k regex:"in_bwd_reduce" -s 3 -c 1 -o gpurun_out/r02_in_bwd_reduce python tools/micro_in.py > /dev/null 2>&1
$NCU -k regex:"in_bwd_apply" -s 3 -c 1 -o gpurun_out/r02_in_bwd_apply python tools/micro_in.py > /dev/null 2>&1
# the four output-parity classes of a stride-2 data gradient in one launch, and the 16-channel PatchGAN input layer
S2="16 256 256 64 128 3 2 1 0"
M0="32 256 256 12 64 4 2 1 0"
$NCU -k regex:s2cls -s 3 -c 1 -o gpurun_out/r02_s2cls python tools/micro_conv.py $S2 dgrad > /dev/null 2>&1
$NCU -k regex:igemm -s 3 -c 1 -o gpurun_out/r02_m0_fprop python tools/micro_conv.py $M0 fprop > /dev/null 2>&1
ls -la gpurun_out/r02_*.ncu-rep
