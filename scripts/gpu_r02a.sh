# usage (under gpurun): bash scripts/gpu_r02a.sh <tag> — GPU test suite, smoke, bench, then the waiting encoder-level
# parity runs side by side (one encoder process per host core; they share the GPU)
TAG=${1:-r02a}
O=gpurun_out; mkdir -p $O
nproc > $O/${TAG}_host.txt; nvidia-smi -L >> $O/${TAG}_host.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/${TAG}_gpu_tests.log; cat $O/${TAG}_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
timeout 600 python bench.py --steps 8 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-1500 $O/${TAG}_bench.json; tail -5 $O/${TAG}_bench.err
run() { ( timeout 2100 bash integration/run_config.sh $1 $2 $3 > $O/${TAG}_encoder_c$1_$2$3.log 2>&1; echo "config $1 $2 $3: rc=$?" >> $O/${TAG}_encoder_rc.txt ) & }
rm -f $O/${TAG}_encoder_rc.txt
run 3 gpu 2
run 2 gpu 12
run 12 gpu
run 9 gpu
run 10 gpu
run 11 gpu
run 13 gpu
run 13 hooks
wait
cat $O/${TAG}_encoder_rc.txt
for f in $O/${TAG}_encoder_c*.log; do echo "== $f"; grep -E "PARITY|wall|vtmcuda|DECODER" $f; done
