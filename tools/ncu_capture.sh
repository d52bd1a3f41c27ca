#!/bin/bash
# ncu captures of one round, run on the GPU box (gpurun).  Only small artefacts come back: wide raw CSVs of a metric list over
# every traversal launch of two steps (room / hall), and two `--set full` reports (three room launches; the persistent kernel).
# usage: bash tools/ncu_capture.sh <tag>
T=${1:-r2}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python tools/profile_step.py 0 2 > gpurun_out/${T}_l.log 2>&1
ncu --metrics $M --clock-control none -k regex:k_trace_q -c 34 --csv --page raw --log-file gpurun_out/${T}_room_raw.csv python tools/profile_step.py 0 2 > gpurun_out/${T}_p1.log 2>&1
PS_PATHS=1310720 ncu --metrics $M --clock-control none -k regex:k_trace_q -c 66 --csv --page raw --log-file gpurun_out/${T}_hall_raw.csv python tools/profile_step.py 0 2 concert_hall 32 > gpurun_out/${T}_p2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_trace_q -s 18 -c 3 -o gpurun_out/${T}_room_full -f python tools/profile_step.py 0 2 > gpurun_out/${T}_p3.log 2>&1
FS_TUNE_MEGA=1 PS_NOTIME=1 ncu --set full --import-source on --clock-control none -k regex:k_path_q -s 1 -c 1 -o gpurun_out/${T}_mega -f python tools/profile_step.py 0 2 > gpurun_out/${T}_p4.log 2>&1
ls -la gpurun_out/
