#!/bin/bash
# One step of both bench workloads on 20 Mb probes under ncu: per-launch duration, DRAM bytes, issue / occupancy figures (CSV logs
# under gpurun_out/), and one --set full capture of the kernel named in $1 (default: indel_align2).  Run on the GPU box:
#   gpurun --timeout 600 -- 'bash tools/ncu_step_metrics.sh'
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__thread_inst_executed_per_inst_executed.ratio,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed
K=${1:-indel_align2}
mkdir -p gpurun_out
# the windows start after the three warm-up steps and hold at least one whole step (tools/kernel_traffic.py cuts it out)
NC_BENCH_LEN=20000000 NC_BENCH_NO_ALL=1 NC_BENCH_NO_CPU=1 timeout 200 ncu --metrics $M --clock-control none -s 90 -c 30 --csv --log-file gpurun_out/r2f_snp_metrics.csv python bench.py --steps 1 --warmup 3 --from-bam 0 > /dev/null 2>&1
NC_BENCH_ALL_LEN=20000000 NC_BENCH_NO_CPU=1 timeout 250 ncu --metrics $M --clock-control none -s 200 -c 140 --csv --log-file gpurun_out/r2f_all_metrics.csv python bench.py --workload all --steps 1 --warmup 3 --from-bam 0 > /dev/null 2>&1
[ "$K" = none ] || NC_BENCH_ALL_LEN=20000000 NC_BENCH_NO_CPU=1 timeout 150 ncu --set full --import-source on --clock-control none -k regex:$K -s 3 -c 1 -o gpurun_out/r2_$K -f python bench.py --workload all --steps 1 --warmup 3 --from-bam 0 > /dev/null 2>&1
ls -la gpurun_out/
