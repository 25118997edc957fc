mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_discriminator_gpu.py -q -m gpu -s > gpurun_out/t3.log 2>&1; echo "pytest(cat default) rc=$?" >> gpurun_out/t3.log
B200SAT_DISC_WGRAD=win timeout 600 python -m pytest tests/test_discriminator_gpu.py tests/test_ae_training_step_gpu.py -q -m gpu -s -k "not window and not batched" > gpurun_out/t3w.log 2>&1; echo "pytest(win) rc=$?" >> gpurun_out/t3w.log
grep -E "window wgrad|passed|failed|rc=|Error|assert " gpurun_out/t3.log | tail -20
grep -E "passed|failed|rc=|Error|assert " gpurun_out/t3w.log | tail -10
echo "--- cat"; timeout 200 python tools/disc_bench.py 32 2>&1 | tail -1
echo "--- win"; B200SAT_DISC_WGRAD=win timeout 200 python tools/disc_bench.py 32 2>&1 | tail -1
