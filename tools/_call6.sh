mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_discriminator_gpu.py tests/test_backward_kernels_gpu.py tests/test_ae_training_step_gpu.py tests/test_autoencoder_train_gpu.py tests/test_dit_gpu.py tests/test_dit_train_gpu.py -q -m gpu -x > gpurun_out/t6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t6.log
grep -E "passed|failed|rc=|Error|assert " gpurun_out/t6.log | tail -8
timeout 200 python tools/disc_bench.py 32 2>&1 | tail -1
