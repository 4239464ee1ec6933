mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in "1 1" "1 2" "0 2"; do
  set -- $v
  SNAPB200_K1_GW=$1 SNAPB200_K1_NP=$2 timeout 200 python bench.py --blocks 32768 --wave 16384 --steps 2 --no-e2e --no-cpu-baseline > gpurun_out/sw_$1_$2.json 2> gpurun_out/sw_$1_$2.err
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/sw_$1_$2.json')); print('GW=$1 NP=$2', round(d['compress_gbs'],2), round(d['decompress_gbs'],2), d['config']['parity'])
except Exception as e:
    print('GW=$1 NP=$2 FAILED', e); print(open('gpurun_out/sw_$1_$2.err').read()[-600:])
"
done
B="python bench.py --blocks 4096 --wave 4096 --steps 1 --no-e2e --no-cpu-baseline --no-parity"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_g1 -s 3 -c 1 -f -o gpurun_out/k1_r1c $B > gpurun_out/ncu_k1c.log 2>&1; tail -2 gpurun_out/ncu_k1c.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k2_decompress -s 3 -c 1 -f -o gpurun_out/k2_r1b $B > gpurun_out/ncu_k2b.log 2>&1; tail -1 gpurun_out/ncu_k2b.log
