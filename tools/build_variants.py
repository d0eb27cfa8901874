"""Developer builds next to the product library (loaded with DSVC_LIB=...):
  libdsvc_tl.so     -DDSVC_TIMELINE                   in-kernel cycle stamps (dsvc_diffnet_run_layer prints them)
  libdsvc_late.so   -DDSVC_NO_EPI_HOIST               epilogue row inputs loaded after the accumulator wait (the old order)
  libdsvc_wd.so     -DDSVC_WATCHDOG                   every mbarrier wait traps after ~2 s instead of hanging the GPU (first runs
                                                      of a new barrier protocol)"""
import os, sys
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffsvc_b200 import _lib

V = {"libdsvc_tl.so": ["-DDSVC_TIMELINE"], "libdsvc_late.so": ["-DDSVC_NO_EPI_HOIST"],
     "libdsvc_wd.so": ["-DDSVC_WATCHDOG"]}
names = sys.argv[1:] or list(V)
with ThreadPoolExecutor(len(names)) as ex:
    for r in ex.map(lambda n: _lib.build(force=True, extra_flags=V[n], out=os.path.join(_lib.LIB_DIR, n)), names):
        print("built", r)
