mkdir -p gpurun_out
L="--profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 400 ncu $L --log-file gpurun_out/r2_launches_ae_adv_b32.csv python tools/profile_step.py ae_train 32 > gpurun_out/p1.log 2>&1
timeout 200 ncu $L --log-file gpurun_out/r2_launches_sample_after.csv python tools/profile_step.py sample > gpurun_out/p2.log 2>&1
timeout 300 ncu $L --log-file gpurun_out/r2_launches_train_pre_after.csv python tools/profile_step.py train_pre > gpurun_out/p3.log 2>&1
timeout 200 python tools/gemm_sweep.py > gpurun_out/gemm_sweep.txt 2>&1
timeout 120 python tools/attn_bwd_bench.py > gpurun_out/attn_bench.txt 2>&1
tail -3 gpurun_out/p1.log gpurun_out/p2.log gpurun_out/p3.log; cat gpurun_out/gemm_sweep.txt; cat gpurun_out/attn_bench.txt
