set -x
timeout 600 python -m pytest tests/test_grouping_interp_gpu.py -x -q --timeout 200 2>&1 | tail -8
timeout 600 python tools/quick_bench.py group > gpurun_out/timings_group2.txt 2>&1; cat gpurun_out/timings_group2.txt
