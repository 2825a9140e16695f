#!/usr/bin/env bash
# Round-2 evidence run (on the GPU box): launch lists, ncu --set full captures of the kernels DESIGN.md quotes, sanitizers.
#   gpurun -- 'bash tools/profile_r2.sh'      -> gpurun_out/r2_*  (text only: the .ncu-rep files are summarised on the box and deleted,
#   gpurun_out/ may not exceed 64 MiB)
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
full() {   # full <summary name> <ncu selection args...> -- <command...>
  local name=$1; shift
  local sel=()
  while [ "$1" != "--" ]; do sel+=("$1"); shift; done
  shift
  $NCU --set full --import-source on "${sel[@]}" -f -o /tmp/prof_$name "$@" > /tmp/ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/prof_$name.ncu-rep gpurun_out/r2_${name}_full.txt > /dev/null 2>&1 || tail -5 /tmp/ncu_$name.log > gpurun_out/r2_${name}_full.txt
  rm -f /tmp/prof_$name.ncu-rep
}
$NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extra > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_emd16k_b4_launches.csv python tools/emd_one.py 4 16384 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_emd2k_b32_launches.csv python tools/emd_one.py 32 2048 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r2_config4_launches.csv python tools/profile_targets.py config4 > /dev/null 2>&1
full nn_filter -k regex:nn_filter -s 3 -c 1 -- python bench.py --steps 2 --warmup 3 --no-extra
full nn_search -k regex:nn_search -s 1 -c 1 -- python bench.py --steps 2 --warmup 3 --no-extra   # the direct kernel (bench times it beside the default)
full chamfer_epilogue -k regex:chamfer_epilogue -s 4 -c 2 -- python bench.py --steps 2 --warmup 3 --no-extra
full emd_row -k regex:emd_row_kernel -s 17 -c 3 -- python tools/emd_one.py 4 16384
full emd_row_b32 -k regex:emd_row_kernel -s 17 -c 3 -- python tools/emd_one.py 32 16384
full emd_pruned -k 'regex:emd_pruned|emd_mask|morton_sort' -s 9 -c 5 -- python tools/emd_one.py 4 16384
full emd_pair -k regex:emd_pair -s 2 -c 2 -- python tools/emd_one.py 4 16384
full matchcost -k regex:matchcost -c 2 -- python tools/profile_targets.py emd16k
full gathers -k 'regex:group_point|three_interpolate' -s 6 -c 6 -- python tools/profile_targets.py config4
full fps -k regex:fps_pruned -s 1 -c 1 -- python tools/profile_targets.py fps
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 python tools/sanitize_targets.py > gpurun_out/r2_sanitizer_$tool.txt 2>&1
  echo "$tool exit code: $?" >> gpurun_out/r2_sanitizer_$tool.txt
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_memcheck_smoke.txt 2>&1
echo "memcheck(smoke) exit code: $?" >> gpurun_out/r2_sanitizer_memcheck_smoke.txt
du -sh gpurun_out
for f in gpurun_out/r2_sanitizer_*.txt; do echo "== $f"; tail -n 4 $f; done
