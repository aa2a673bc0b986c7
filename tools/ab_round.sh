#!/bin/bash
# A/B of kernel build variants (build/libfiss_*.so): short bench + executed-instruction / issue counters of both lattice
# kernel variants per library.  usage: gpurun -- 'bash tools/ab_round.sh TAG [pytest]'  (pytest runs on the in-tree library)
TAG=${1:-ab}
mkdir -p gpurun_out
if [ "$2" == "pytest" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.log
fi
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
for lib in build/libfiss_*.so; do
  n=$(basename $lib .so)
  FISSGPU_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu-baseline --no-closed-loop --steps 100 > gpurun_out/${TAG}_$n.json 2>gpurun_out/${TAG}_$n.err
  FISSGPU_LIB=$PWD/$lib timeout 300 ncu --metrics $M --clock-control none -k regex:fiss_grid_kernel -s 6 -c 24 --csv \
     --log-file gpurun_out/${TAG}_$n.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-closed-loop > /dev/null 2>>gpurun_out/${TAG}_$n.err
  python - <<PY
import csv, json, collections
try:
    d = json.load(open("gpurun_out/${TAG}_$n.json"))
    print("$n kernel_ms=%.4f frac=%.3f value=%.1fM winner_only=%.1fM e2e=%.1fM p50=%.4f" % (d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["value"]/1e6, d["value_winner_only"]/1e6, d["e2e"]["value"]/1e6, d["plan_cycle_p50_ms"]))
except Exception as e:
    print("$n bench FAILED", e)
try:
    rows = [r for r in csv.reader(open("gpurun_out/${TAG}_$n.csv")) if len(r) > 10]
    h = rows[0]
    acc = collections.defaultdict(list)
    for r in rows[1:]:
        # only the B = 512 launches of the bench loop (grid = all resident CTAs)
        acc[(r[h.index("Kernel Name")][:34], r[h.index("Grid Size")], r[h.index("Metric Name")])].append(float(r[h.index("Metric Value")].replace(",", "")))
    for k in sorted(acc):
        v = acc[k]
        print("     %-34s grid %-12s %-55s %14.1f  (n=%d)" % (k[0], k[1], k[2], sum(v) / len(v), len(v)))
except Exception as e:
    print("$n counters FAILED", e)
PY
done
# optional: one full capture of the winner-only lattice kernel at B = 512 (the e2e path) with the in-tree library
if [ "$3" == "full0" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fiss_grid_kernel -s 14 -c 1 -f -o gpurun_out/${TAG}_grid0 \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-closed-loop > gpurun_out/${TAG}_ncu_grid0.log 2>&1
  ls -la gpurun_out/${TAG}_grid0.ncu-rep
fi
