#!/bin/bash
# ncu evidence for one round: launch list of one step (gpu__time_duration of every kernel) and --set full captures of
# selected launches, summarised ON THE BOX (tools/ncu_summary.py, ncu_l2.py, ncu_hot.py) -- the reports are ~12 MB each and
# gpurun brings back at most 64 MiB, so only the text summaries and the reports named in KEEP survive.
#   usage: bash tools/gpu_ncu.sh TAG [precision] [batch] ["name:skip ..."] ["keep names"]
TAG=${1:-r2}; PREC=${2:-fp32}; BATCH=${3:-32}
CAPS=${4:-"first:0 down0:1 res_conv1:14 res_conv2:15 up1:27 final:28"}
KEEP=${5:-}
OUT=gpurun_out
mkdir -p $OUT
exec </dev/null
echo "== ncu launch list ($PREC, B=$BATCH)"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $OUT/launches_${PREC}_b${BATCH}_$TAG.csv python tools/prof_step.py --precision $PREC --batch $BATCH > $OUT/ncu_list_$TAG.log 2>&1
tail -1 $OUT/ncu_list_$TAG.log | cut -c1-160
SUM=$OUT/ncu_${PREC}_b${BATCH}_$TAG.md
(cd /tmp && cuobjdump -xelf all $OLDPWD/eamm_b200/lib/conv_tc.o > /dev/null 2>&1); CUBIN=/tmp/conv_tc.sm_100a.cubin
: > $SUM
summarise() {   # name report
  echo "## $1" >> $SUM
  python tools/ncu_summary.py $2 >> $SUM 2>&1
  echo '```' >> $SUM; python tools/ncu_l2.py $2 >> $SUM 2>&1; python tools/ncu_hot.py $2 14 >> $SUM 2>&1
  case $1 in warp) ;; *) python tools/ncu_lines.py $2 $CUBIN auto 40 >> $SUM 2>&1 ;; esac
  echo '```' >> $SUM
  case " $KEEP " in *" $1 "*) ;; *) rm -f $2 ;; esac
}
for item in $CAPS; do
  name=${item%%:*}; skip=${item##*:}
  echo "== ncu --set full: $name (conv_tc launch $skip)"
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s $skip -c 1 -f \
    -o $OUT/prof_${name}_${PREC}_b${BATCH}_$TAG python tools/prof_step.py --precision $PREC --batch $BATCH > $OUT/ncu_${name}_$TAG.log 2>&1
  summarise $name $OUT/prof_${name}_${PREC}_b${BATCH}_$TAG.ncu-rep
done
echo "== ncu --set full: warp_occlude"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:warp_occlude -c 1 -f \
  -o $OUT/prof_warp_${PREC}_b${BATCH}_$TAG python tools/prof_step.py --precision $PREC --batch $BATCH > $OUT/ncu_warp_$TAG.log 2>&1
summarise warp $OUT/prof_warp_${PREC}_b${BATCH}_$TAG.ncu-rep
grep -h "gpu__time_duration.sum\|tma_ld.sum \|tensor_cycles" $SUM | head -60
du -sh $OUT
