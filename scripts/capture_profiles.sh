#!/bin/bash
# Runs on the GPU box (under gpurun): ncu evidence for profiles/.  Usage: bash scripts/capture_profiles.sh <tag>
# Keeps what it writes well under gpurun's 64 MiB merge limit: raw CSV pages are exported here, reports are deleted.
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
# 1. every launch of the bench command with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph --no-video --no-model --no-kernel-head \
    --no-postprocess > $OUT/${TAG}_launches_bench.log 2>&1
# 2. full captures, one launch of each kind, from the second decode step of a 3-step run
cap() {  # name, kernel regex, skip, count
    ncu --set full --clock-control none -k regex:"$2" -s $3 -c $4 -o $OUT/${TAG}_ncu_full_$1 \
        python scripts/run_stage.py 4 128 256 3 > $OUT/${TAG}_ncu_full_$1.log 2>&1
    ncu -i $OUT/${TAG}_ncu_full_$1.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_full_$1.raw.csv 2>/dev/null
    rm -f $OUT/${TAG}_ncu_full_$1.ncu-rep
    tail -1 $OUT/${TAG}_ncu_full_$1.log
}
# KernelHead tail (pf_kernel_head): launch list + one full capture of its three kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches_kernel_head.csv \
    python scripts/head_timing.py 4 > $OUT/${TAG}_launches_kernel_head.log 2>&1
ncu --set full --clock-control none -k regex:"head_apply|gn_finalize|einsum_kernel" -s 9 -c 3 -o $OUT/${TAG}_ncu_full_kernel_head \
    python scripts/head_timing.py 4 > $OUT/${TAG}_ncu_full_kernel_head.log 2>&1
ncu -i $OUT/${TAG}_ncu_full_kernel_head.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_full_kernel_head.raw.csv 2>/dev/null
rm -f $OUT/${TAG}_ncu_full_kernel_head.ncu-rep
cap stream "pool_kernel|einsum_kernel|upsample2x|binarise" 9 9
# the same streaming kernels inside the step with the caches left alone between replays (L2 warm: x_feats evict-last)
ncu --set full --clock-control none --cache-control none -k regex:"pool_kernel|einsum_kernel|upsample2x|binarise" -s 9 -c 9 \
    -o $OUT/${TAG}_ncu_full_stream_l2warm python scripts/run_stage.py 4 128 256 3 > $OUT/${TAG}_ncu_full_stream_l2warm.log 2>&1
ncu -i $OUT/${TAG}_ncu_full_stream_l2warm.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_full_stream_l2warm.raw.csv 2>/dev/null
rm -f $OUT/${TAG}_ncu_full_stream_l2warm.ncu-rep
cap tcgemm "tcgemm" 27 9
cap helpers "prep_kernel|sumln_kernel|attention_kernel" 9 3
# round 2: the neck (pf_semantic_fpn + pf_fpn_pred), the tracking path and the batched panoptic merge
for what in neck track panoptic; do
    ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches_$what.csv \
        python scripts/neck_track_timing.py $what 4 1 > $OUT/${TAG}_launches_$what.log 2>&1
done
capx() {  # name, kernel regex, skip, count, driver args
    ncu --set full --clock-control none -k regex:"$2" -s $3 -c $4 -o $OUT/${TAG}_$1 \
        python scripts/neck_track_timing.py $5 4 1 > $OUT/${TAG}_$1.log 2>&1
    ncu -i $OUT/${TAG}_$1.ncu-rep --page raw --csv > $OUT/${TAG}_$1.raw.csv 2>/dev/null
    rm -f $OUT/${TAG}_$1.ncu-rep
    tail -1 $OUT/${TAG}_$1.log
}
capx ncu_full_neck "sgemm_conv256|fpn_" 44 22 neck
capx ncu_full_track "sgemm_kernel|roi_align|box_|fc_tail|tracker_match" 20 10 track
capx ncu_full_panoptic "pp_" 8 4 panoptic
ls -la $OUT
