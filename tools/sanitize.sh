mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
# K1 must not read past the end of a caller allocation: caching allocator off so that tensors end where they end
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck_edge.log python -m pytest tests/test_gpu_parity.py -x -q -k "reads_stay_inside or unit_limits" 2>&1 | tail -2
echo "memcheck(edge) rc=$?"; grep -E "ERROR SUMMARY|Invalid|out of bounds" gpurun_out/memcheck_edge.log | head -6
if [ "$1" = "full" ]; then
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log python -m pytest tests/test_gpu_parity.py -x -q -k "golden or kats or close_to_end or tiny or config1 or foreign_tag or crc32c or html or sweeps or device_frame" 2>&1 | tail -4
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|out of bounds" gpurun_out/memcheck.log | head -8
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/racecheck.log python -m pytest tests/test_gpu_parity.py -x -q -k "golden or config1 or sweeps" 2>&1 | tail -3
grep -E "RACECHECK SUMMARY|hazard" gpurun_out/racecheck.log | head -8
fi
