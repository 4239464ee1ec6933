mkdir -p gpurun_out
for m in 0 1; do
  SNAPB200_K1_MULTI=$m timeout 200 python bench.py --blocks 65536 --wave 32768 --steps 2 --no-e2e --no-cpu-baseline > gpurun_out/m_$m.json 2> gpurun_out/m_$m.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/m_$m.json')); print('MULTI=$m compress', round(d['compress_gbs'],2), 'decompress', round(d['decompress_gbs'],2), d['config']['parity'])
except Exception as e:
    print('MULTI=$m FAILED'); print(open('gpurun_out/m_$m.err').read()[-700:])"
done
SNAPB200_K1_MULTI=1 timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
