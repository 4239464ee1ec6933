mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
NGS="${NGS:-4 5 6 7}" bash tools/ng_exp.sh
