#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 300 $TR tools/p2p_probe.py > gpurun_out/n2_p2p_probe.json 2> gpurun_out/n2_p2p_probe.err; echo "probe exit $?"; cat gpurun_out/n2_p2p_probe.json
run() { # name, extra args..., env
  name=$1; shift
  timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu "$@" > gpurun_out/n2d_$name.json 2> gpurun_out/n2d_$name.err
  echo "$name exit $?"; python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/n2d_$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["ms_per_step"],3), j["segments_ms"]["per_rank"], j["config"].get("halo_rows_sent_frac"))
except Exception as e:
    print("$name", "parse failed", e)
PY
}
run halo_zero --exchange halo --peer-debug zero_mask
run halo_local --exchange halo --peer-debug local_peers
run fused_local --exchange fused --peer-debug local_peers
run halo --exchange halo
run fused --exchange fused
GGAD_B200_LIB=$PWD/ab_peer_st0.so run halo_st0 --exchange halo
GGAD_B200_LIB=$PWD/ab_peer_st0.so run fused_st0 --exchange fused
GGAD_B200_LIB=$PWD/ab_peer_st2.so run halo_st2 --exchange halo
run multicast --exchange multicast
