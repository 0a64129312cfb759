#!/bin/bash
# Per-launch ncu metrics of ONE hot-path forward (KITTI shape).  Run on the GPU box:  bash profiles/capture_forward.sh r1f
# The plain run must exit 0 first; the report stays in /tmp (only the CSV comes back through gpurun_out/).
set -e
tag=${1:-cap}
python profiles/run_forward.py --n 2 > gpurun_out/${tag}_plain.log 2>&1
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats \
    --section Occupancy --section WarpStateStats \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum \
    --clock-control none -k regex:"conv_tc|volume_fused|disp_att|avgpool|tap_gather|softmax_reg|convex_up|class_stats" \
    -s 48 -c 48 -o /tmp/${tag}_full python profiles/run_forward.py --n 2 > gpurun_out/${tag}_ncu.log 2>&1
ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv
