import numpy as np, sys
sys.path.insert(0,'.')
import vtm_b200
ms=vtm_b200.MotionSearch(0)
for (w,h) in [(8,8),(24,16),(8,8),(40,56),(136,72),(1920,1080)]:
    a=np.ascontiguousarray(np.random.randint(0,1023,(h,w)).astype(np.int16))
    try:
        ms.upload_picture(30,a); print(w,h,"ok")
    except Exception as e: print(w,h,"ERR",str(e)[:200])
ms2=vtm_b200.MotionSearch(0)
for (w,h) in [(24,16),(8,8)]:
    a=np.ascontiguousarray(np.random.randint(0,1023,(h,w)).astype(np.int16))
    try:
        ms2.upload_picture(30,a); print("fresh ctx",w,h,"ok")
    except Exception as e: print("fresh",w,h,"ERR",str(e)[:200])
