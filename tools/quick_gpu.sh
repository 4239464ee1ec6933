mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --blocks 65536 --wave 32768 --steps 2 --e2e-blocks 65536 --no-cpu-baseline > gpurun_out/q.json 2> gpurun_out/q.err
python -c "
import json
try:
    d=json.load(open('gpurun_out/q.json')); print('e2e', round(d['e2e']['value'],2), 'compress', round(d['compress_gbs'],2), 'decompress', round(d['decompress_gbs'],2), 'value', round(d['value'],2), d['config']['parity'])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/q.err').read()[-800:])
"
