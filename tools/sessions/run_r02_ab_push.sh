#!/bin/bash
# Round-2 A/B (2 GPUs): end-of-tile CTA-wide halo push (shipped) vs per-group push (-DGGAD_PUSH_PER_GROUP=1).
# Build the variant HERE first (it travels with the snapshot):
#     python -m ggad_b200.build --out=$PWD/ab_push_group.so -DGGAD_PUSH_PER_GROUP=1
#     gpurun --gpus 2 --timeout 900 -- 'bash tools/sessions/run_r02_ab_push.sh'
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
# parity of the variant first (single-GPU push test + 2-GPU halo parity)
GGAD_B200_LIB=$PWD/ab_push_group.so timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_multi.py -m gpu -x -q -k "exchange or sharded or halo" > gpurun_out/r02_ab_pytest.log 2>&1
echo "variant pytest exit $?"; tail -3 gpurun_out/r02_ab_pytest.log
for lib in "" "$PWD/ab_push_group.so"; do
  tag=$([ -z "$lib" ] && echo shipped || echo pergroup)
  GGAD_B200_LIB=$lib CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/bench_variants.py --workload S64 2>&1 | grep -i "push\|plain unweighted" | sed "s/^/$tag /"
  GGAD_B200_LIB=$lib timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-cpu > gpurun_out/r02_ab_$tag.json 2> gpurun_out/r02_ab_$tag.err
  python - <<PY
import json
j=json.loads(open("gpurun_out/r02_ab_$tag.json").read().strip().splitlines()[-1])
print("$tag", round(j["ms_per_step"],3), round(j["value"]/1e9,2), j["segments_ms"]["per_rank"])
PY
done
