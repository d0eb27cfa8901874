"""Phase stamps of the step kernel (needs libdsvc_tl.so: tools/build_variants.py libdsvc_tl.so)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffsvc_b200 as D
from diffsvc_b200 import _lib
from diffsvc_b200.hparams import hparams, DEFAULTS_44K
import synthetic as S
T = int(sys.argv[1]) if len(sys.argv) > 1 else 862
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hparams.clear(); hparams.update(DEFAULTS_44K); hparams["pndm_speedup"] = 1
lib = _lib.load()
dn = D.DiffNet(128, math_mode="tc3f16"); dn.load_state_dict(S.synth_diffnet_weights())
gd = D.GaussianDiffusion(None, 128, dn, timesteps=1000, K_step=1000, spec_min=[-5.0], spec_max=[0.0]).cuda().eval()
g = torch.Generator().manual_seed(1)
cond = (torch.randn(B, 256, T, generator=g) * 0.5).cuda(); x0 = torch.randn(B, 1, 128, T, generator=g).cuda()
gd.sample(x0, cond, 2, None, None, seed=1); torch.cuda.synchronize()
print("== lib %s  B=%d T=%d" % (os.environ.get("DSVC_LIB", "product"), B, T), flush=True)
_lib.check(lib.dsvc_diffnet_run_layer(dn.handle(), 0, 3, 6, _lib.current_stream())); torch.cuda.synchronize()
