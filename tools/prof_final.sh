mkdir -p gpurun_out
B="python bench.py --blocks 8288 --wave 8288 --steps 1 --no-e2e --no-cpu-baseline --no-parity"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_m7 -s 3 -c 1 -f -o gpurun_out/k1_r1d $B > gpurun_out/ncu_k1d.log 2>&1; tail -1 gpurun_out/ncu_k1d.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k2_decompress -s 3 -c 1 -f -o gpurun_out/k2_r1c $B > gpurun_out/ncu_k2c.log 2>&1; tail -1 gpurun_out/ncu_k2c.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --blocks 16576 --wave 8288 --steps 2 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/ncu_l2.log 2>&1
timeout 300 python tools/frame_bench.py --gib 1 --verify 2>&1 | tail -1
timeout 300 python tools/frame_bench.py --gib 8 2>&1 | tail -1
