# round-2 profile captures (run under gpurun on one GPU): ncu --set full of K1 and K2 on the text workload,
# and the launch list of a short default bench run. Reports land in gpurun_out/ (summaries -> profiles/).
mkdir -p gpurun_out
B="python bench.py --blocks 8288 --wave 8288 --steps 1 --no-e2e --no-cpu-baseline --no-parity"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_m7 -s 3 -c 1 -f -o gpurun_out/k1_r2 $B > gpurun_out/ncu_k1_r2.log 2>&1; tail -1 gpurun_out/ncu_k1_r2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k2_decompress -s 3 -c 1 -f -o gpurun_out/k2_r2 $B > gpurun_out/ncu_k2_r2.log 2>&1; tail -1 gpurun_out/ncu_k2_r2.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2.csv python bench.py --blocks 16576 --wave 8288 --steps 2 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/ncu_l_r2.log 2>&1; tail -1 gpurun_out/ncu_l_r2.log
