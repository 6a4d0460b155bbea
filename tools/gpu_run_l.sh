set -x
cd $GRAFT_REPO_ROOT
timeout 900 bash tools/ncu_summary.sh gpurun_out/ncu_r2n scan1d scan1d_bad minmax1d minmaximum > /dev/null 2>&1
ls gpurun_out/ncu_r2n
