mkdir -p gpurun_out
timeout 150 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 1500 gpurun_out/bench_full.json
B="python bench.py --blocks 8288 --wave 8288 --steps 1 --no-e2e --no-cpu-baseline --no-parity"
timeout 90 ncu --set full --clock-control none --import-source on -k regex:k1_m7 -s 3 -c 1 -f -o gpurun_out/k1_r1e $B > gpurun_out/ncu_k1e.log 2>&1; tail -1 gpurun_out/ncu_k1e.log
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --blocks 16576 --wave 8288 --steps 2 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/ncu_l3.log 2>&1
