#!/bin/bash
# usage (under gpurun): bash scripts/gpu_r02u.sh "<cfg> <frames>" ["<cfg> <frames>" ...] — encoder-level parity runs side by side
# (one encoder process per configuration, they share the GPU); exit status 0 only if every run reports PARITY OK
O=gpurun_out; mkdir -p $O
rm -f $O/r02u_encoder_rc.txt
for spec in "$@"; do
  set -- $spec
  ( timeout ${TMO:-3300} bash integration/run_config.sh $1 gpu $2 > $O/r02u_encoder_c$1_gpu$2.log 2>&1; echo "config $1 gpu $2: rc=$?" >> $O/r02u_encoder_rc.txt ) &
done
wait
cat $O/r02u_encoder_rc.txt
for f in $O/r02u_encoder_c*.log; do echo "== $f"; grep -E "PARITY|wall|vtmcuda|DECODER" $f; done
rm -f $O/enc_c*/in.yuv $O/enc_c*/rec_*.yuv $O/enc_c*/dec_*.yuv
! grep -v "rc=0" $O/r02u_encoder_rc.txt
