#!/bin/sh
# parameter sweep of the trace kernel (developer tool)
for occ in 6 7 8; do for ss in 8 12 16; do
  echo "occ=$occ smem_stack=$ss: $(TRB_TRACE_OCC=$occ TRB_SMEM_STACK=$ss python bench.py --steps 3 --warmup 1 --spp-per-step 8 --no-cpu-baseline 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("%.1f Mrays/s %.1f ms frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"]))')"
done; done
