# A/B the product library against experimental builds on one GPU (run under gpurun; build the variants first, here):
#   bash tools/build_variant.sh mbar -DK1_MBAR      # emitters woken through an mbarrier instead of sleep-polling
#   bash tools/build_variant.sh spec -DK1_GT_SPEC   # L2-table chains: speculative slot reads
#   bash tools/build_variant.sh w64 -DK1_W64        # 64 positions per parser step on the shared-memory-table chains
#   bash tools/build_variant.sh una -DK1_UNALIGNED  # 32-position windows start where the parse stands
#   (the flags combine: build_variant.sh all -DK1_W64 -DK1_UNALIGNED -DK1_GT_SPEC -DK1_MBAR)
# usage: bash tools/ab_variants.sh [suffix ...]      (each suffix = rust-snappy_b200/libsnapb200_<suffix>.so)
mkdir -p gpurun_out
run() {
  label=$1; shift
  env "$@" timeout 200 python bench.py --blocks 131072 --wave 65536 --steps 2 --no-e2e --no-cpu-baseline > gpurun_out/ab_$label.json 2> gpurun_out/ab_$label.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/ab_$label.json')); print('$label compress', round(d['compress_gbs'],2), 'decompress', round(d['decompress_gbs'],2), d['config']['parity'])
except Exception as e:
    print('$label FAILED', e); print(open('gpurun_out/ab_$label.err').read()[-600:])
"
}
run product X=1
for sfx in "$@"; do run $sfx SNAPB200_LIB=$PWD/rust-snappy_b200/libsnapb200_$sfx.so; done
