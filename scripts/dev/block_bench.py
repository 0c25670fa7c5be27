import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from vtm_b200.lib import load_library
L = load_library()
L.vtmme_dev_sad_block_bench.restype = C.c_double
L.vtmme_dev_sad_block_bench.argtypes = [C.c_int] * 6
torch.zeros(1, device="cuda")
peak = 64 * 148 * 1.965e9
for (nfp, fpu, dy, thr, ctas) in [(0,2,1,256,3),(2,2,1,256,3),(3,2,1,256,3),(12,2,1,256,3),(13,2,1,256,3),(14,2,1,256,3),(13,2,1,224,3),(14,2,1,224,3),(0,0,1,256,3),(0,1,1,256,3),(2,1,1,256,3),(2,1,1,224,3),(3,1,1,256,3),(1,1,1,256,3),(0,0,2,256,2),(2,1,2,256,2),(2,1,2,224,2),(2,1,2,160,3),(2,0,2,256,2),(3,1,2,256,2),(1,1,2,256,2),(2,1,1,128,6),(2,1,1,192,4)]:
    r = L.vtmme_dev_sad_block_bench(nfp, fpu, dy, thr, ctas, 40)
    print("nfp=%d fpu=%d dy=%d thr=%d ctas=%d : %.3e px-cand/s = %.3f of VABSDIFF roofline (%.1f px/clk/SM)" % (nfp, fpu, dy, thr, ctas, r, r / peak, r / 148 / 1.965e9))
