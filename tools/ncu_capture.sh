#!/bin/bash
# ncu captures of one round, run on the GPU box (gpurun).  Only small artefacts come back.
#   <tag>_launches.csv           launch list of the bench command itself (gpu__time_duration, cold-cache, serialised: SHARES)
#   <tag>_{room,hall}_pq_raw.csv  metric list over three launches of the persistent per-batch kernel (the default for these jobs)
#   <tag>_{room,hall}_tq_raw.csv  the same over every per-bounce traversal launch of two updates (FS_TUNE_MEGA=0)
#   <tag>_room_pq_full.ncu-rep, <tag>_hall_pq_full.ncu-rep   `--set full --import-source on`, one launch each
# usage: bash tools/ncu_capture.sh <tag> [all|pq]
T=${1:-r2}; W=${2:-all}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
mkdir -p gpurun_out
# (the BVH builds of the bench's three contexts are > 8000 launches: only the kernels of an IR update are listed)
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:k_(reset_queues|shade_gen|pq_seed|pq_finish|path_q|trace_q|trace_any|trace_closest|connect_gen|connect_all_gen|eval|eval_mis|energy|ir|ir_spectra|lis_store|lis_load|init_ends)' -c 2000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_l.log 2>&1
ncu --metrics $M --clock-control none -k regex:k_path_q -c 3 --csv --page raw --log-file gpurun_out/${T}_room_pq_raw.csv python tools/profile_step.py 0 3 > gpurun_out/${T}_p1.log 2>&1
PS_PATHS=1310720 ncu --metrics $M --clock-control none -k regex:k_path_q -c 3 --csv --page raw --log-file gpurun_out/${T}_hall_pq_raw.csv python tools/profile_step.py 0 3 concert_hall 32 > gpurun_out/${T}_p2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_path_q -s 1 -c 1 -o gpurun_out/${T}_room_pq_full -f python tools/profile_step.py 0 2 > gpurun_out/${T}_p3.log 2>&1
PS_PATHS=1310720 ncu --set full --import-source on --clock-control none -k regex:k_path_q -s 1 -c 1 -o gpurun_out/${T}_hall_pq_full -f python tools/profile_step.py 0 2 concert_hall 32 > gpurun_out/${T}_p4.log 2>&1
if [ "$W" = "all" ]; then
FS_TUNE_MEGA=0 ncu --metrics $M --clock-control none -k regex:k_trace_q -c 34 --csv --page raw --log-file gpurun_out/${T}_room_tq_raw.csv python tools/profile_step.py 0 2 > gpurun_out/${T}_p5.log 2>&1
FS_TUNE_MEGA=0 PS_PATHS=1310720 ncu --metrics $M --clock-control none -k regex:k_trace_q -c 66 --csv --page raw --log-file gpurun_out/${T}_hall_tq_raw.csv python tools/profile_step.py 0 2 concert_hall 32 > gpurun_out/${T}_p6.log 2>&1
fi
ls -la gpurun_out/ | grep ${T}_
