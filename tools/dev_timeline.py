"""In-kernel cycle stamps of the WaveNet layer kernels (needs a -DDSVC_TIMELINE build: tools/build_variants.py, then
DSVC_LIB=diffsvc_b200/lib/libdsvc_tl.so python tools/dev_timeline.py [T]).  The library prints the stamps itself."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PARTS = [int(x) for x in os.environ.get("DSVC_TL_PARTS", "0,1,2").split(",")]   # 2 = the fused layer kernel
if 2 in PARTS:
    os.environ["DSVC_FUSED_LAYER"] = "2"
import torch
import diffsvc_b200 as D
from diffsvc_b200 import _lib
from diffsvc_b200.hparams import hparams, DEFAULTS_44K
from oracle import diffsvc_oracle as O

T = int(sys.argv[1]) if len(sys.argv) > 1 else 862
hparams.clear(); hparams.update(DEFAULTS_44K); hparams["pndm_speedup"] = 1
lib = _lib.load()
dn = D.DiffNet(128, math_mode="tc3f16"); dn.load_state_dict(O.synth_diffnet_weights())
gd = D.GaussianDiffusion(None, 128, dn, timesteps=1000, K_step=1000, spec_min=[-5.0], spec_max=[0.0]).cuda().eval()
g = torch.Generator().manual_seed(1)
cond = (torch.randn(1, 256, T, generator=g) * 0.5).cuda(); x0 = torch.randn(1, 1, 128, T, generator=g).cuda()
gd.sample(x0, cond, 2, None, None, seed=1); torch.cuda.synchronize()
h = dn.handle()
print("== lib %s  T=%d" % (os.environ.get("DSVC_LIB", "product"), T), flush=True)
for part in PARTS:
    _lib.check(lib.dsvc_diffnet_run_layer(h, 3, part, 4, _lib.current_stream())); torch.cuda.synchronize()
if 2 in PARTS:
    os.environ["DSVC_FUSED_FENCE"] = "1"
    print("== fused, device-scope fence (DSVC_FUSED_FENCE=1)", flush=True)
    _lib.check(lib.dsvc_diffnet_run_layer(h, 3, 2, 4, _lib.current_stream())); torch.cuda.synchronize()
