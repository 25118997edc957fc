mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_discriminator_gpu.py tests/test_ae_training_step_gpu.py tests/test_dit_gpu.py tests/test_dropin_gpu.py -x -q -m gpu -s > gpurun_out/t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t2.log
grep -E "rel err|cos|passed|failed|Error|error|assert" gpurun_out/t2.log | tail -40
echo "--- new"; timeout 200 python tools/disc_bench.py 32 2>&1 | tail -2
echo "--- old"; B200SAT_DISC_CONV0=simt B200SAT_DISC_WGRAD_CAT=0 B200SAT_DISC_POST_V1=1 timeout 200 python tools/disc_bench.py 32 2>&1 | tail -1
L="--profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 300 ncu $L --log-file gpurun_out/r2_launches_ae_adv_b32_after.csv python tools/profile_step.py ae_train 32 > gpurun_out/p4.log 2>&1
tail -n 2 gpurun_out/p4.log
