#!/bin/bash
# round-2 GPU call 19 (last of the round, 17 GPU-minutes left): the periodic-path gpu tests first, then the full gpu suite on the
# final library, then memcheck of one periodic test and a short bench line if the budget allows
mkdir -p gpurun_out
(time timeout 240 python -m pytest tests/test_gpu_periodic.py -q -x --durations=8) > gpurun_out/c19_periodic.log 2>&1
tail -25 gpurun_out/c19_periodic.log
(time timeout 600 python -m pytest tests -m gpu -q -x --durations=5 --deselect tests/test_gpu_periodic.py) > gpurun_out/c19_pytest.log 2>&1
tail -12 gpurun_out/c19_pytest.log
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_periodic.py -q -x -k "loopy or roundtrip" > gpurun_out/c19_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/c19_memcheck.log | tail -3
timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/c19_bench.json 2> gpurun_out/c19_bench.err
grep "^{" gpurun_out/c19_bench.json | cut -c1-260; tail -2 gpurun_out/c19_bench.err
