#!/bin/bash
# Encoder-level parity run (BASELINE configs 1-3): patched VTM (oracle/_ref/EncoderAppCUDA) with the GPU motion
# search against the golden md5 of the unmodified CPU encoder (tests/golden/encoder_md5.json).
#   bash integration/run_config.sh <1..12> [gpu|cpu] [frames]   (4 = random access, 5 = LD-B with SearchRange 128, both on the small clip;
#   6 / 7 = configs 1 / 4 with FastSearch=1, the TZ search: xTZSearch on the GPU; 8 = config 1 with FastSearch=2, the selective search;
#   9 = config 1 with FastSearch=3; 10 / 11 = config 4 with FastSearch=2 / 3; 12 = config 2 (first 3 pictures) with FastSearch=1;
#   13 = 64x64, 2 pictures, low delay P: small enough for MODE=hooks; 14 = small RA clip, FastSearch=1, with the GOP-based temporal
#   filter: EncTemporalFilter::filter on the GPU)
# MODE: gpu = VTMME_ENABLE=1; hooks = VTMME_ENABLE=1 VTMME_TABLE_HOOKS=1 (distortion / interpolation dispatch tables on the
# GPU too, one launch per table call); cpu = the unmodified encoder (regenerates a golden).
# Exit status: 0 only when the bitstream md5 equals the golden and the decoder's picture hashes are all OK.
# Work directory gpurun_out/enc_c<n>_<frames>f_<mode>: several configurations can run side by side.
set -e
CFGN=${1:-1}; MODE=${2:-gpu}; FS=0; ROOT=$(cd "$(dirname "$0")/.." && pwd)
case $CFGN in
  1) WD=416; HT=240; FR=${3:-8}; BITS=8; CFG=encoder_lowdelay_P_vtm.cfg; SR=64; EXTRA="-q 32";;
  2) WD=1920; HT=1080; FR=${3:-32}; BITS=10; CFG=encoder_randomaccess_vtm.cfg; SR=64; EXTRA="-q 32 --IntraPeriod=32";;
  4) WD=416; HT=240; FR=${3:-9}; BITS=8; CFG=encoder_randomaccess_vtm.cfg; SR=64; EXTRA="-q 32 --IntraPeriod=32";;   # small RA (bi-pred) case
  5) WD=416; HT=240; FR=${3:-5}; BITS=8; CFG=encoder_lowdelay_vtm.cfg; SR=128; EXTRA="-q 32";;   # config 3's tools (LD-B, SR=128) on the small clip
  6) WD=416; HT=240; FR=${3:-8}; BITS=8; CFG=encoder_lowdelay_P_vtm.cfg; SR=64; EXTRA="-q 32"; FS=1;;   # config 1 with the TZ search
  7) WD=416; HT=240; FR=${3:-9}; BITS=8; CFG=encoder_randomaccess_vtm.cfg; SR=64; EXTRA="-q 32 --IntraPeriod=32"; FS=1;;   # small RA with the TZ search
  8) WD=416; HT=240; FR=${3:-4}; BITS=8; CFG=encoder_lowdelay_P_vtm.cfg; SR=64; EXTRA="-q 32"; FS=2;;   # config 1 with the selective TZ search (xTZSearchSelective, staged SAD)
  9) WD=416; HT=240; FR=${3:-8}; BITS=8; CFG=encoder_lowdelay_P_vtm.cfg; SR=64; EXTRA="-q 32"; FS=3;;   # config 1 with the enhanced TZ search
  10) WD=416; HT=240; FR=${3:-9}; BITS=8; CFG=encoder_randomaccess_vtm.cfg; SR=64; EXTRA="-q 32 --IntraPeriod=32"; FS=2;;   # small RA with the selective search
  11) WD=416; HT=240; FR=${3:-9}; BITS=8; CFG=encoder_randomaccess_vtm.cfg; SR=64; EXTRA="-q 32 --IntraPeriod=32"; FS=3;;   # small RA with the enhanced TZ search
  12) WD=1920; HT=1080; FR=${3:-3}; BITS=10; CFG=encoder_randomaccess_vtm.cfg; SR=64; EXTRA="-q 32 --IntraPeriod=32"; FS=1;;   # config 2's picture size with the TZ search (CTC default)
  3) WD=3840; HT=2160; FR=${3:-16}; BITS=10; CFG=encoder_lowdelay_vtm.cfg; SR=128; EXTRA="-q 32";;
  14) WD=416; HT=240; FR=${3:-11}; BITS=8; CFG=encoder_randomaccess_vtm.cfg; SR=64; EXTRA="-q 32 --IntraPeriod=32 --TemporalFilter=1"; FS=1;;   # small RA clip with the GOP-based temporal filter (pictures 0 and 8 filtered)
  13) WD=64; HT=64; FR=${3:-2}; BITS=8; CFG=encoder_lowdelay_P_vtm.cfg; SR=64; EXTRA="-q 32";;   # tiny clip for the table hooks
  *) echo "unknown configuration $CFGN"; exit 2;;
esac
W=$ROOT/gpurun_out/enc_c${CFGN}_${FR}f_$MODE${WD_SUFFIX:-}; mkdir -p $W; cd $W
python $ROOT/integration/make_yuv.py in.yuv --width $WD --height $HT --frames $FR --bits $BITS
echo "input md5: $(md5sum in.yuv | cut -d' ' -f1)"
ARGS="-c $ROOT/oracle/_ref/cfg/$CFG -i in.yuv -wdt $WD -hgt $HT -fr 30 -f $FR $EXTRA --InputBitDepth=$BITS --FastSearch=$FS --SearchRange=$SR --SEIDecodedPictureHash=1"
if [ "$MODE" = gpu ]; then
  export VTMME_ENABLE=1; BIN=$ROOT/oracle/_ref/EncoderAppCUDA
elif [ "$MODE" = hooks ]; then
  export VTMME_ENABLE=1 VTMME_TABLE_HOOKS=1; BIN=$ROOT/oracle/_ref/EncoderAppCUDA
else
  BIN=$ROOT/oracle/_ref/EncoderApp
fi
T0=$SECONDS
$BIN $ARGS -b out_$MODE.bin -o rec_$MODE.yuv > enc_$MODE.log 2> enc_$MODE.err || { tail -20 enc_$MODE.err; tail -5 enc_$MODE.log; exit 1; }
echo "wall $((SECONDS-T0)) s ($MODE)"
grep -E "Total Time" enc_$MODE.log; tail -3 enc_$MODE.err
DECRC=0; $ROOT/oracle/_ref/DecoderApp -b out_$MODE.bin -o dec_$MODE.yuv > dec_$MODE.log 2>&1 || DECRC=$?
echo "bitstream md5: $(md5sum out_$MODE.bin | cut -d' ' -f1)"
echo "recon md5:     $(md5sum rec_$MODE.yuv | cut -d' ' -f1)"
echo "decoded md5:   $(md5sum dec_$MODE.yuv | cut -d' ' -f1)  (decoder hash check: $(grep -c '(OK)' dec_$MODE.log) OK, $(grep -c 'ERROR' dec_$MODE.log) ERROR)"
NERR=$(grep -c 'ERROR' dec_$MODE.log || true)
RC=0; python - <<PY || RC=$?
import hashlib, json, sys
gold = json.load(open("$ROOT/tests/golden/encoder_md5.json"))
bs = hashlib.md5(open("out_$MODE.bin","rb").read()).hexdigest()
# several goldens may exist per configuration (e.g. config2 = first 3 pictures, config2_full = all 32): match on -f
hit = [g for k, g in gold.items() if k.split("_")[0] == "config$CFGN" and g["args"].split("-f ")[1].split()[0] == "$FR"]
rc = 0
if "$MODE" == "cpu":
    print("golden run: nothing to compare against")
elif hit:
    ok = bs == hit[0]["bitstream_md5"]
    print("PARITY", "OK" if ok else "MISMATCH", "bitstream md5 vs golden", hit[0]["bitstream_md5"])
    rc = 0 if ok else 1
else:
    print("no golden for this configuration/frame count")
    rc = 1
if $DECRC != 0 or $NERR != 0:
    print("DECODER FAILED: exit status $DECRC, $NERR picture hash errors")
    rc = 1
sys.exit(rc)
PY
rm -f in.yuv rec_$MODE.yuv dec_$MODE.yuv
exit $RC
