"""GPU: the patched encoder (VTM 9.3 + libvtmme hooks, oracle/_ref/EncoderAppCUDA, built by __graft_entry__.build() where
/root/reference exists) on a tiny clip — bitstream md5 against the golden of the unmodified CPU encoder
(tests/golden/encoder_md5.json, `config13`), once with the motion search on the GPU (VTMME_ENABLE=1) and once with the
distortion / interpolation dispatch tables on the GPU as well (VTMME_TABLE_HOOKS=1: RdCost::initRdCostCUDA and
InterpolationFilter::initInterpolationFilterCUDA, the siblings of x86/InitX86.cpp:57-75,104-122)."""
import hashlib
import json
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENC = os.path.join(ROOT, "oracle", "_ref", "EncoderAppCUDA")
DEC = os.path.join(ROOT, "oracle", "_ref", "DecoderApp")


def _encode(tmp_path, env_extra, tag):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "encoder_md5.json")))["config13"]
    yuv = tmp_path / "in.yuv"
    if not yuv.exists():
        subprocess.check_call([sys.executable, os.path.join(ROOT, "integration", "make_yuv.py"), str(yuv), "--width", "64",
                               "--height", "64", "--frames", "2", "--bits", "8"])
    assert hashlib.md5(yuv.read_bytes()).hexdigest() == gold["input_md5"]
    args = gold["args"].replace("-c cfg/", "-c " + os.path.join(ROOT, "oracle", "_ref", "cfg") + "/").split()
    out = tmp_path / ("out_%s.bin" % tag)
    env = dict(os.environ, **env_extra)
    p = subprocess.run([ENC] + args + ["-i", str(yuv), "-b", str(out), "-o", str(tmp_path / "rec.yuv")], env=env,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return gold, out, p.stderr


@pytest.mark.skipif(not os.path.exists(ENC), reason="oracle/_ref/EncoderAppCUDA was not built (needs /root/reference at build time)")
def test_encoder_md5_with_gpu_search(tmp_path):
    gold, out, err = _encode(tmp_path, {"VTMME_ENABLE": "1"}, "gpu")
    m = re.search(r"GPU motion searches: (\d+)", err)
    assert m and int(m.group(1)) > 1000, err[-500:]
    assert hashlib.md5(out.read_bytes()).hexdigest() == gold["bitstream_md5"]
    dec = subprocess.run([DEC, "-b", str(out), "-o", str(tmp_path / "dec.yuv")], capture_output=True, text=True)
    assert dec.returncode == 0 and "ERROR" not in dec.stdout and dec.stdout.count("(OK)") == 2


@pytest.mark.skipif(not os.path.exists(ENC), reason="oracle/_ref/EncoderAppCUDA was not built (needs /root/reference at build time)")
def test_encoder_md5_with_table_hooks(tmp_path):
    gold, out, err = _encode(tmp_path, {"VTMME_ENABLE": "1", "VTMME_TABLE_HOOKS": "1"}, "hooks")
    m = re.search(r"table hooks: distortion (\d+) on the GPU / (\d+) delegated, filters (\d+) on the GPU / \d+ delegated, "
                  r"affine gradient entries (\d+) on the GPU", err)
    assert m, err[-500:]
    # the entries of all three dispatch tables (RdCost, InterpolationFilter, AffineGradientSearch) really ran on the GPU
    assert int(m.group(1)) > 10000 and int(m.group(3)) > 10000 and int(m.group(4)) > 1000
    assert hashlib.md5(out.read_bytes()).hexdigest() == gold["bitstream_md5"]
