#!/bin/bash
# ncu captures of one pipeline step at bs 256 (run under gpurun, one GPU): usage tools/ncu_round.sh <round-tag>
# 1. light sections for every launch of the step  -> gpurun_out/<tag>_step_light.ncu-rep
# 2. --set full + source for one launch of each kernel of interest -> gpurun_out/<tag>_<name>_full.ncu-rep
# 3. the per-launch duration list of the bench command -> gpurun_out/<tag>_launches.csv
tag=${1:-r2}
set -x
python tools/prof_run.py 256 1 > gpurun_out/${tag}_plain.log 2>&1 || exit 1
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis_Tables --section LaunchStats --section Occupancy --section SchedulerStats \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -f \
    -o gpurun_out/${tag}_step_light python tools/prof_run.py 256 1 > gpurun_out/${tag}_ncu_light.log 2>&1
full() {   # name regex skip
  ncu --set full --import-source on --clock-control none -f --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 \
      -o gpurun_out/${tag}_$1_full python tools/prof_run.py 256 1 > gpurun_out/${tag}_ncu_$1.log 2>&1
}
full b128 'k_umma_bottleneck<\(int\)128, \(int\)32, \(int\)32, \(int\)128, \(int\)2, \(int\)1' 3
full b64 'k_umma_bottleneck<\(int\)64, \(int\)16, \(int\)16, \(int\)64' 1
full init 'k_umma_initial_u8' 0
full head 'k_umma_head' 0
full up4 'k_umma_up<\(int\)128' 0
full up5 'k_umma_up<\(int\)64' 0
full s5 'k_stage5' 0
full occ 'k_occgrid' 0
python bench.py --steps 2 --warmup 3 --skip-cpu --skip-latency --skip-contour --skip-config5 > gpurun_out/${tag}_ncu_plain_bench.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-latency --skip-contour --skip-config5 > gpurun_out/${tag}_ncu_bench.log 2>&1
ls -la gpurun_out/${tag}_*
