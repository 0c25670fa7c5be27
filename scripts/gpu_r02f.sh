# usage (under gpurun): bash scripts/gpu_r02f.sh <tag> — the GOP-based temporal filter wired into the encoder (config 14)
TAG=${1:-r02f}
O=gpurun_out; mkdir -p $O
timeout 900 bash integration/run_config.sh 14 gpu > $O/${TAG}_encoder_c14_gpu.log 2>&1; echo "c14 rc=$?"; grep -E "PARITY|wall|vtmcuda|DECODER" $O/${TAG}_encoder_c14_gpu.log; tail -3 $O/${TAG}_encoder_c14_gpu.log
