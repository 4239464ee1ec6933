# sweep the second-generation K1 parser's L2-table chain count (and the legacy parser) on one GPU
mkdir -p gpurun_out
run() {
  label=$1; shift
  env "$@" timeout 200 python bench.py --blocks 131072 --wave 65536 --steps 2 --no-e2e --no-cpu-baseline > gpurun_out/xs_$label.json 2> gpurun_out/xs_$label.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/xs_$label.json')); print('$label compress', round(d['compress_gbs'],2), 'decompress', round(d['decompress_gbs'],2), d['config']['parity'])
except Exception as e:
    print('$label FAILED', e); print(open('gpurun_out/xs_$label.err').read()[-600:])
"
}
run legacy SNAPB200_K1_X=0
for ng in "$@"; do run x_ng$ng SNAPB200_K1_X=1 SNAPB200_K1_NG=$ng; done
