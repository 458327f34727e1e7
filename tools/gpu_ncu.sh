#!/bin/bash
# ncu evidence for one round: launch list of one step (gpu__time_duration of every kernel) and --set full captures of
# selected launches.   usage: bash tools/gpu_ncu.sh TAG [precision] [batch] ["name:skip ..."]
TAG=${1:-r2}; PREC=${2:-fp32}; BATCH=${3:-32}
CAPS=${4:-"first:0 down0:1 res_conv1:14 res_conv2:15 up1:27 final:28"}
OUT=gpurun_out
mkdir -p $OUT
exec </dev/null
echo "== ncu launch list ($PREC, B=$BATCH)"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $OUT/launches_${PREC}_b${BATCH}_$TAG.csv python tools/prof_step.py --precision $PREC --batch $BATCH > $OUT/ncu_list_$TAG.log 2>&1
tail -1 $OUT/ncu_list_$TAG.log | cut -c1-160
for item in $CAPS; do
  name=${item%%:*}; skip=${item##*:}
  echo "== ncu --set full: $name (conv_tc launch $skip)"
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s $skip -c 1 -f \
    -o $OUT/prof_${name}_${PREC}_b${BATCH}_$TAG python tools/prof_step.py --precision $PREC --batch $BATCH > $OUT/ncu_${name}_$TAG.log 2>&1
  tail -1 $OUT/ncu_${name}_$TAG.log | cut -c1-160
done
echo "== ncu --set full: warp_occlude"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:warp_occlude -c 1 -f \
  -o $OUT/prof_warp_${PREC}_b${BATCH}_$TAG python tools/prof_step.py --precision $PREC --batch $BATCH > $OUT/ncu_warp_$TAG.log 2>&1
tail -1 $OUT/ncu_warp_$TAG.log | cut -c1-160
ls -la $OUT/*.ncu-rep | tail -12
