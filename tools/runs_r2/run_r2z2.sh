set -x
cd $GRAFT_REPO_ROOT
python tools/e2e_trace.py 2> gpurun_out/r2z_e2e_trace.txt; grep -v "^+" gpurun_out/r2z_e2e_trace.txt | awk 'NR%9<3 || /---/ || /per call/' | cut -c1-1500
