mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --blocks 32768 --wave 16384 --steps 2 --no-e2e --no-cpu-baseline > gpurun_out/q.json 2> gpurun_out/q.err
python -c "
import json
try:
    d=json.load(open('gpurun_out/q.json')); print('compress', round(d['compress_gbs'],2), 'decompress', round(d['decompress_gbs'],2), 'value', round(d['value'],2), d['config']['parity'])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/q.err').read()[-800:])
"
