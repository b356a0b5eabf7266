M=gpu__time_duration.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,sm__icc_request_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for b in 1480 2960 5920 14800 65536; do
  python scripts/phase_probe.py $b | cut -c1-100
  ncu --metrics $M --clock-control none -k regex:lse_kernel -s 3 -c 1 --csv --log-file gpurun_out/pp_$b.csv python scripts/phase_probe.py $b > /dev/null 2>&1
  python - gpurun_out/pp_$b.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
h=rows[0]
print("   ", " ".join("%s=%s"%(dict(zip(h,r))['Metric Name'].split('__')[1][:30],dict(zip(h,r))['Metric Value']) for r in rows[1:]))
PY
done
