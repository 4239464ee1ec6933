# K1 hybrid-table experiment: extra chains per SM with L2-resident tables (SNAPB200_K1_NG)
mkdir -p gpurun_out
for ng in ${NGS:-0 4 7}; do
  SNAPB200_K1_NG=$ng timeout 200 python bench.py --blocks 131072 --wave 65536 --steps 2 --no-e2e --no-cpu-baseline > gpurun_out/ng$ng.json 2> gpurun_out/ng$ng.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/ng$ng.json')); print('NG=$ng compress', round(d['compress_gbs'],2), 'decompress', round(d['decompress_gbs'],2), 'value', round(d['value'],2), d['config']['parity'])
except Exception as e:
    print('NG=$ng FAILED', e); print(open('gpurun_out/ng$ng.err').read()[-800:])
"
done
for ng in ${PROF_NGS:-}; do
  SNAPB200_K1_NG=$ng timeout 200 ncu --metrics gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.avg.per_cycle_active,sm__warps_active.avg.per_cycle_active,lts__t_sectors.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k1_m7 -s 3 -c 1 --csv --log-file gpurun_out/ngprof$ng.csv python bench.py --blocks 32768 --wave 32768 --steps 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/ngprof$ng.log 2>&1
  echo "PROF NG=$ng"; grep -v "^==" gpurun_out/ngprof$ng.csv | cut -d, -f13- | tail -9
done
