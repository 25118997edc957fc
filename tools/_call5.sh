mkdir -p gpurun_out
cap() { name=$1; regex=$2; skip=$3; shift 3
  timeout 300 ncu --set full --import-source on --clock-control none -k "regex:$regex" -s $skip -c 1 -f -o gpurun_out/r2_ncu_$name "$@" > gpurun_out/r2_ncu_$name.log 2>&1
  tail -1 gpurun_out/r2_ncu_$name.log | cut -c1-120; }
cap attention_fwd_v3 "attention_fwd" 0 python tools/attn_once.py
cap attention_bwd_dkv_v3 "attention_bwd_dkv" 0 python tools/attn_once.py
cap attention_bwd_dq_v3 "attention_bwd_dq" 0 python tools/attn_once.py
cap disc_wgrad_window "disc_wgrad_window" 6 python tools/disc_bench.py 32
cap disc_convpost_fwd "disc_convpost_fwd" 2 python tools/disc_bench.py 32
cap disc_spec_pack "disc_spec_pack" 2 python tools/disc_bench.py 32
timeout 400 python tools/seq_sweep.py --steps 4 > gpurun_out/r2_seq_sweep_1gpu_after.jsonl 2> gpurun_out/seq_sweep.err
cat gpurun_out/r2_seq_sweep_1gpu_after.jsonl | cut -c1-400
