mkdir -p gpurun_out
B="python bench.py --blocks 4096 --wave 4096 --steps 1 --no-e2e --no-cpu-baseline --no-parity"
SNAPB200_K1_NP=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_g2 -s 3 -c 1 -f -o gpurun_out/k1_np2 $B > gpurun_out/ncu_np2.log 2>&1; tail -1 gpurun_out/ncu_np2.log
