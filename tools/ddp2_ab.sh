mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus 2 --steps 10 --warmup 3 --quick --no-sample --no-cpu-baseline --no-gpu-reference > gpurun_out/r2g_ddp2_$tag.json 2> gpurun_out/r2g_ddp2_$tag.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2g_ddp2_$tag.json')); print('$tag', round(d['value']), 'tok/s', round(d['ms_per_step'],2), 'ms; pre-encoded', round(d['train_pre_encoded']['ms_per_step'],2), 'ms', d['whole_step']['ddp'], d['clocks'])
except Exception as e: print('$tag failed', e); print(open('gpurun_out/r2g_ddp2_$tag.err').read()[-1500:])
"; }
run default X=1
run old B200SAT_DDP_NCCL_CTAS=0 B200SAT_DDP_SM_RESERVE=0
run c4r8 B200SAT_DDP_NCCL_CTAS=4 B200SAT_DDP_SM_RESERVE=8
run c16r32 B200SAT_DDP_NCCL_CTAS=16 B200SAT_DDP_SM_RESERVE=32
timeout 300 python -m pytest tests/test_ae_training_step_gpu.py -m gpu -q -s 2>&1 | grep -E "T2 step|passed|failed|^E |all watched" | head -20
