# build an experimental variant of the library next to the product one: build_variant.sh <suffix> <extra nvcc flags...>
# select it at run time with SNAPB200_LIB=rust-snappy_b200/libsnapb200_<suffix>.so
set -e
cd "$(dirname "$0")/.."
sfx=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -Iinclude "$@" \
  -o rust-snappy_b200/libsnapb200_$sfx.so rust-snappy_b200/csrc/snapb200.cu
