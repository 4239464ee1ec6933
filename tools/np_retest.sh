mkdir -p gpurun_out
for v in "1 1 1" "0 1 2" "0 1 3" "0 1 1"; do
  set -- $v
  SNAPB200_K1_MULTI=$1 SNAPB200_K1_GW=$2 SNAPB200_K1_NP=$3 timeout 200 python bench.py --blocks 65536 --wave 32768 --steps 2 --no-e2e --no-cpu-baseline > gpurun_out/np_$1_$2_$3.json 2> gpurun_out/np_$1_$2_$3.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/np_$1_$2_$3.json')); print('MULTI=$1 GW=$2 NP=$3 compress', round(d['compress_gbs'],2), d['config']['parity'])
except Exception as e:
    print('MULTI=$1 GW=$2 NP=$3 FAILED'); print(open('gpurun_out/np_$1_$2_$3.err').read()[-500:])"
done
