set -x
cd $GRAFT_REPO_ROOT
python tools/pcie_probe.py > gpurun_out/r2z_pcie_probe.json 2> gpurun_out/r2z_pcie_probe.err; cat gpurun_out/r2z_pcie_probe.json; tail -3 gpurun_out/r2z_pcie_probe.err
