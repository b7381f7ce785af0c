#!/bin/bash
# tools/ncu_kproj.sh <tag>: `ncu --set full` captures of single k_project_* launches of the second registration of
# tools/probe_once.py (enqueue-all loop: ncu cannot profile kernel nodes of conditional graphs).  Launch order per
# projection: one k_project launch (30 per registration in the enqueue-all loop).  The reports (44 MB each) stay on the
# box; their raw and SASS-level pages come back as CSV.
tag=$1
cap() { # name skip
  PLO_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:k_project --launch-skip $2 -c 1 -f \
    -o /tmp/${tag}_$1 python tools/probe_once.py > gpurun_out/${tag}_$1.log 2>&1
  ncu -i /tmp/${tag}_$1.ncu-rep --page raw --csv > gpurun_out/${tag}_$1_raw.csv 2>/dev/null
  ncu -i /tmp/${tag}_$1.ncu-rep --page source --csv --print-source sass > gpurun_out/${tag}_$1_sass.csv 2>/dev/null
}
cap walk_it1 30       # first projection: tree walk, no temporal bound
cap store_it3 32      # third projection: tree walk that leaves the tiles behind
cap tiles_it5 34      # fifth projection: groups from tiles + the misses through the tree
cap tiles_it7 36      # seventh projection
ls -la gpurun_out/${tag}_*
