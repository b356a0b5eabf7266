# instruction-cache counters of one kernel launch: icc_probe.sh <out-prefix> <kernel-regex> <skip> <quick_perf args...>
M=gpu__time_duration.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,gcc__average_cache_request_hit_rate.pct,sm__icc_requests.sum,sm__icc_request_hit_rate.pct,sm__icc_requests_lookup_miss_tag_miss.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,gcc__xbar2gcc_sectors.sum
out=$1; k=$2; skip=$3; shift 3
ncu --metrics $M --clock-control none -k regex:$k -s $skip -c 1 --csv --log-file gpurun_out/$out.csv python scripts/quick_perf.py "$@" > /dev/null 2>&1
python - gpurun_out/$out.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
h=rows[0]
print(sys.argv[1], " ".join("%s=%s"%(dict(zip(h,r))['Metric Name'].split('__')[1][:34],dict(zip(h,r))['Metric Value']) for r in rows[1:]))
PY
