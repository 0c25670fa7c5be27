TAG=${1:-r02n}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_affine.py tests/test_gpu_encoder.py -m gpu -q 2>&1 | tail -6 > $O/${TAG}_gpu_tests.log; cat $O/${TAG}_gpu_tests.log
timeout 900 bash integration/run_config.sh 13 hooks > $O/${TAG}_encoder_c13_hooks.log 2>&1; echo "c13 hooks rc=$?"; grep -E "PARITY|wall|vtmcuda|DECODER" $O/${TAG}_encoder_c13_hooks.log
