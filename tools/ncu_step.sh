#!/bin/bash
# tools/ncu_step.sh <tag>: `ncu --set full` of EVERY kernel of one scan-to-map step (index build, source upload,
# k_register_loop) -- the second step of tools/probe_once.py -- plus the PCA-normal pass and the front-end kernels.
# The reports stay on the box; the raw pages (one row per launch) and the loop kernel's SASS page come back as CSV.
tag=$1
ncu --set full --clock-control none --import-source on --launch-skip 25 -c 25 -f -o /tmp/${tag}_step python tools/probe_once.py > gpurun_out/${tag}_step.log 2>&1
ncu -i /tmp/${tag}_step.ncu-rep --page raw --csv > gpurun_out/${tag}_step_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_register_loop --launch-skip 1 -c 1 -f -o /tmp/${tag}_loop python tools/probe_once.py > gpurun_out/${tag}_loop.log 2>&1
ncu -i /tmp/${tag}_loop.ncu-rep --page raw --csv > gpurun_out/${tag}_loop_raw.csv 2>/dev/null
ncu -i /tmp/${tag}_loop.ncu-rep --page source --csv --print-source sass > gpurun_out/${tag}_loop_sass.csv 2>/dev/null
ncu --set full --clock-control none -k regex:k_pca_normals -c 1 -f -o /tmp/${tag}_pca python tools/probe_pca.py > gpurun_out/${tag}_pca.log 2>&1
ncu -i /tmp/${tag}_pca.ncu-rep --page raw --csv > gpurun_out/${tag}_pca_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:k_fe_ --launch-skip 22 -c 11 -f -o /tmp/${tag}_fe python tools/probe_frontend.py > gpurun_out/${tag}_fe.log 2>&1
ncu -i /tmp/${tag}_fe.ncu-rep --page raw --csv > gpurun_out/${tag}_fe_raw.csv 2>/dev/null
ls -la gpurun_out/${tag}_*
