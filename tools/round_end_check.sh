mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/ -x -q -m gpu) > gpurun_out/final_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_tests.log
tail -4 gpurun_out/final_tests.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
(time timeout 900 python bench.py) > gpurun_out/r2_bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final.err | cut -c1-300
head -c 400 gpurun_out/r2_bench_final.json; echo
