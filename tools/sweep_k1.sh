mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in "0 1" "0 2" "0 3" "0 4" "1 1" "1 2" "1 3"; do
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
