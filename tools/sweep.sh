#!/bin/sh
# parameter sweep of the trace kernel (developer tool)
for occ in 4 6 8; do for refill in 4 8 16; do for tg in 8 16; do
  echo "occ=$occ refill=$refill grid=$tg: $(TRB_TRACE_OCC=$occ TRB_REFILL=$refill TRB_TRACE_GRID=$tg python bench.py --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("%.1f Mrays/s %.1f ms" % (d["value"], d["ms_per_step"]))')"
done; done; done
