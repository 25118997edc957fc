mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dit_gpu.py -q -m gpu -s -k "inpaint or sampler" > gpurun_out/t4.log 2>&1; echo "pytest(inpaint) rc=$?" >> gpurun_out/t4.log
grep -E "inpaint |passed|failed|rc=|Error|assert " gpurun_out/t4.log | tail -20
B200SAT_DISC_WGRAD=win timeout 600 python -m pytest tests/test_discriminator_gpu.py tests/test_ae_training_step_gpu.py -q -m gpu -k "not window and not batched" > gpurun_out/t4w.log 2>&1; echo "pytest(win) rc=$?" >> gpurun_out/t4w.log
grep -E "passed|failed|rc=|Error|assert " gpurun_out/t4w.log | tail -6
echo "--- cat"; timeout 200 python tools/disc_bench.py 32 2>&1 | tail -1
echo "--- win"; B200SAT_DISC_WGRAD=win timeout 200 python tools/disc_bench.py 32 2>&1 | tail -1
