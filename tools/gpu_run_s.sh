set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tools/sanitize_small.py
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 python tools/sanitize_small.py > gpurun_out/r2s_memcheck.txt 2>&1; echo "rc=$?" >> gpurun_out/r2s_memcheck.txt
tail -6 gpurun_out/r2s_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --launch-timeout 0 python tools/sanitize_small.py > gpurun_out/r2s_racecheck.txt 2>&1; echo "rc=$?" >> gpurun_out/r2s_racecheck.txt
grep -c "hazard" gpurun_out/r2s_racecheck.txt; grep "RACECHECK SUMMARY\|sanitize_small" gpurun_out/r2s_racecheck.txt; grep -A3 "hazard detected" gpurun_out/r2s_racecheck.txt | grep "at \|Function" | sed 's/^ *//' | sort | uniq -c | sort -rn | head -20
