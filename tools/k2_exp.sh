mkdir -p gpurun_out
for occ in 0 1; do
  SNAPB200_K2_OCC=$occ timeout 200 python bench.py --blocks 65536 --wave 32768 --steps 2 --no-e2e --no-cpu-baseline > gpurun_out/k2occ_$occ.json 2> gpurun_out/k2occ_$occ.err
  python -c "
import json
d=json.load(open('gpurun_out/k2occ_$occ.json')); print('OCC=$occ compress', round(d['compress_gbs'],2), 'decompress', round(d['decompress_gbs'],2))"
done
