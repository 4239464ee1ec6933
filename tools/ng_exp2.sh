mkdir -p gpurun_out
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 200 python bench.py --blocks 131072 --wave 65536 --steps 2 --no-e2e --no-cpu-baseline > gpurun_out/x_$label.json 2> gpurun_out/x_$label.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/x_$label.json')); print('$label compress', round(d['compress_gbs'],2), 'decompress', round(d['decompress_gbs'],2), d['config']['parity'])
except Exception as e:
    print('$label FAILED', e); print(open('gpurun_out/x_$label.err').read()[-600:])
"
}
run base_ng4_wide SNAPB200_K1_NG=4 SNAPB200_K1_WIDE=1
for ng in 4 5 6; do run ef_ng$ng SNAPB200_K1_NG=$ng SNAPB200_LIB=$PWD/rust-snappy_b200/libsnapb200_ef.so; done
for ng in 4 5 6 7; do run efel_ng$ng SNAPB200_K1_NG=$ng SNAPB200_LIB=$PWD/rust-snappy_b200/libsnapb200_efel.so; done
run ef_ng0 SNAPB200_K1_NG=0 SNAPB200_LIB=$PWD/rust-snappy_b200/libsnapb200_ef.so
