#!/bin/bash
# ncu --set full captures of one launch each of the kernels that have no summary yet; reports land in gpurun_out/r2_ncu_<name>.ncu-rep
mkdir -p gpurun_out
cap() { name=$1; regex=$2; what=$3; skip=${4:-0}
  timeout 240 ncu --set full --import-source on --clock-control none --profile-from-start off -k "regex:$regex" -s $skip -c 1 -f -o gpurun_out/r2_ncu_$name python tools/profile_step.py $what > gpurun_out/r2_ncu_$name.log 2>&1
  tail -1 gpurun_out/r2_ncu_$name.log | cut -c1-160; }
cap layernorm_bwd "layernorm_bwd" train_pre 3
cap layernorm_fwd "layernorm_kernel" train_pre 3
cap adamw "adamw_ema" train_pre
cap gemm256 "gemm_bf16_tcgen05" train_pre 40
cap colsum "colsum" train_pre 3
cap conv_k7 "conv1d_tcgen05" ae 0
cap conv_k1 "conv1d_tcgen05" ae 1
cap conv_in "conv_in_fast" ae
cap snake_bwd "snake_bwd" ae_train 2
cap stft_loss "stft_loss_kernel|stft_loss_accum" ae_train 1
cap stft_bwd "stft_loss_bwd" ae_train 1
cap disc_conv0_wgrad "disc_conv0_wgrad" ae_train
cap disc_act_bwd "disc_act_bwd" ae_train 1
cap conv_wgrad "gemm_bf16_tcgen05" ae_train 5
