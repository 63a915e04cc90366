#!/bin/bash
# Round profile: launch lists + one `ncu --set full` capture per hot kernel (run on the GPU box through gpurun).
#   bash tools/profile_round.sh <tag> [quick]      -> gpurun_out/<tag>_*.csv / *.ncu-rep
set -u
TAG=${1:-prof}
QUICK=${2:-}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file $OUT/${TAG}_launches_train_s4.csv python tools/profile_step.py --scale 4 --steps 2 > $OUT/${TAG}_l4.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file $OUT/${TAG}_launches_sample_s4.csv python tools/profile_step.py --scale 4 --steps 0 --sample-steps 3 > $OUT/${TAG}_ls4.log 2>&1
# second training step (launch indices continue after the first step's 18 tc_conv / 10 tc_wgrad launches)
$NCU --set full --import-source on -k regex:tc_conv_kernel -s 22 -c 1 -o $OUT/${TAG}_tc_conv -f python tools/profile_step.py --scale 4 --steps 2 > $OUT/${TAG}_p1.log 2>&1
$NCU --set full --import-source on -k regex:tc_wgrad_kernel -s 14 -c 1 -o $OUT/${TAG}_tc_wgrad -f python tools/profile_step.py --scale 4 --steps 2 > $OUT/${TAG}_p2.log 2>&1
$NCU --set full --import-source on -k regex:dw5x5_tma_kernel -s 7 -c 1 -o $OUT/${TAG}_dw5x5 -f python tools/profile_step.py --scale 4 --steps 2 > $OUT/${TAG}_p3.log 2>&1
if [ -z "$QUICK" ]; then
$NCU --metrics gpu__time_duration.sum --csv --log-file $OUT/${TAG}_launches_train_s0.csv python tools/profile_step.py --scale 0 --steps 2 > $OUT/${TAG}_l0.log 2>&1
$NCU --set full --import-source on -k regex:tc_conv_kernel -s 19 -c 1 -o $OUT/${TAG}_tc_conv_n80 -f python tools/profile_step.py --scale 4 --steps 2 > $OUT/${TAG}_p1b.log 2>&1
$NCU --set full --import-source on -k regex:dw5x5_wgrad_tma_kernel -s 3 -c 1 -o $OUT/${TAG}_dw5x5_wgrad -f python tools/profile_step.py --scale 4 --steps 2 > $OUT/${TAG}_p4.log 2>&1
$NCU --set full --import-source on -k regex:fused_allreduce_adam_ema_kernel -s 1 -c 1 -o $OUT/${TAG}_fused_step -f python tools/profile_step.py --scale 0 --steps 3 > $OUT/${TAG}_p5.log 2>&1
fi
ls -la $OUT | grep ${TAG}
