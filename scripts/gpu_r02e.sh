# usage (under gpurun): bash scripts/gpu_r02e.sh <tag> — per-PU batching of the in-loop searches: config 1 alone (timing) with and
# without batching, then the other small configurations side by side (parity)
TAG=${1:-r02e}
O=gpurun_out; mkdir -p $O
timeout 600 bash integration/run_config.sh 13 gpu > $O/${TAG}_encoder_c13_gpu.log 2>&1; echo "c13 rc=$?"; grep -E "PARITY|vtmcuda" $O/${TAG}_encoder_c13_gpu.log
timeout 900 bash integration/run_config.sh 1 gpu > $O/${TAG}_encoder_c1_batch.log 2>&1; echo "c1 batch rc=$?"; grep -E "PARITY|wall|vtmcuda" $O/${TAG}_encoder_c1_batch.log
VTMME_BATCH=0 timeout 900 bash integration/run_config.sh 1 gpu > $O/${TAG}_encoder_c1_nobatch.log 2>&1; echo "c1 nobatch rc=$?"; grep -E "PARITY|wall|vtmcuda" $O/${TAG}_encoder_c1_nobatch.log
run() { ( timeout 1500 bash integration/run_config.sh $1 $2 $3 > $O/${TAG}_encoder_c$1_$2$3.log 2>&1; echo "config $1 $2 $3: rc=$?" >> $O/${TAG}_encoder_rc.txt ) & }
rm -f $O/${TAG}_encoder_rc.txt
run 4 gpu
run 5 gpu
run 6 gpu
run 7 gpu
run 8 gpu
run 2 gpu 3
wait
cat $O/${TAG}_encoder_rc.txt
for f in $O/${TAG}_encoder_c[2-8]*.log; do echo "== $f"; grep -E "PARITY|wall|vtmcuda|DECODER" $f; done
