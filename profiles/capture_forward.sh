#!/bin/bash
# Per-launch ncu metrics of ONE hot-path forward (KITTI shape).  Run on the GPU box:  bash profiles/capture_forward.sh r2b [regex-of-one-kernel-for-a-full-capture]
# The plain run must exit 0 first; the reports stay in /tmp or gpurun_out (only CSVs / one small .ncu-rep come back).
set -e
tag=${1:-cap}
full=${2:-}
python profiles/run_forward.py --n 2 > gpurun_out/${tag}_plain.log 2>&1
REGEX="conv_tc|volume_fused|disp_att|avgpool|tap_gather|softmax_reg|convex_up|class_stats"
# both forwards are captured; summarize_forward.py keeps the second (warm) half of the launches
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats \
    --section Occupancy --section WarpStateStats \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum \
    --clock-control none -k regex:"$REGEX" \
    -c 400 -o /tmp/${tag}_full python profiles/run_forward.py --n 2 > gpurun_out/${tag}_ncu.log 2>&1
ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv
if [ -n "$full" ]; then
  ncu --set full --import-source on --clock-control none -k regex:"$full" -s 3 -c 1 -o gpurun_out/${tag}_${full}_full \
      python profiles/run_forward.py --n 2 > gpurun_out/${tag}_ncu_full.log 2>&1
  ncu -i gpurun_out/${tag}_${full}_full.ncu-rep --page source --csv > gpurun_out/${tag}_${full}_source.csv || true
fi
