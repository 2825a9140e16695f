#!/usr/bin/env bash
# Round-2 evidence run (on the GPU box): launch lists, ncu --set full captures of the kernels DESIGN.md quotes, sanitizers.
#   gpurun -- 'bash tools/profile_r2.sh'      -> gpurun_out/r2_*  (summarised into profiles/ by tools/ncu_summary.py, tools/launch_summary.py)
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extra > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_emd16k_b4_launches.csv python tools/emd_one.py 4 16384 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_emd2k_b32_launches.csv python tools/emd_one.py 32 2048 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r2_config4_launches.csv python tools/profile_targets.py config4 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:nn_search -s 3 -c 1 -f -o gpurun_out/prof_nn_r2 python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/ncu_nn_r2.log 2>&1
$NCU --set full --import-source on -k regex:chamfer_epilogue -s 4 -c 2 -f -o gpurun_out/prof_chamfer_epi_r2 python bench.py --steps 2 --warmup 3 --no-extra > /dev/null 2>&1
$NCU --set full --import-source on -k regex:emd_ -s 36 -c 36 -f -o gpurun_out/prof_emd_r2 python tools/emd_one.py 4 16384 > gpurun_out/ncu_emd_r2.log 2>&1
$NCU --set full --import-source on -k regex:emd_ -s 36 -c 36 -f -o gpurun_out/prof_emd_b32_r2 python tools/emd_one.py 32 16384 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:matchcost -c 3 -f -o gpurun_out/prof_matchcost_r2 python tools/profile_targets.py emd16k > /dev/null 2>&1
$NCU --set full --import-source on -k 'regex:group_point|three_interpolate' -s 6 -c 6 -f -o gpurun_out/prof_gathers_r2 python tools/profile_targets.py config4 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:fps_pruned -s 1 -c 1 -f -o gpurun_out/prof_fps_r2 python tools/profile_targets.py fps > /dev/null 2>&1
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 python tools/sanitize_targets.py > gpurun_out/r2_sanitizer_$tool.txt 2>&1
  echo "$tool exit code: $?" >> gpurun_out/r2_sanitizer_$tool.txt
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_memcheck_smoke.txt 2>&1
echo "memcheck(smoke) exit code: $?" >> gpurun_out/r2_sanitizer_memcheck_smoke.txt
ls -la gpurun_out | grep r2 | head -40
tail -3 gpurun_out/r2_sanitizer_*.txt
