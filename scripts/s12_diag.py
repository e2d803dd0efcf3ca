"""Where does the cfg2 S=12 gradient error against the fp64 oracle sit?  Per-image errors + clamp-tie census."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.golden import cases
from tests.helpers import rel_err, run_oracle_loss
from tests.gpu_helpers import run_cuda_loss
name = "_cfg2_s12"
S = int(sys.argv[1]) if len(sys.argv) > 1 else 12
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 122
cases.LOSS_CASES[name] = (32, 768, 90, dict(feature_samples=S, pos_intra_shift=0.2103, pos_inter_shift=0.1233,
                                            neg_inter_shift=0.9748, depth_feat_shift=0.0359), seed)
inputs = cases.make_loss_inputs(name)
cfg, t, r = run_cuda_loss(name, channels_last=True, inputs=inputs)
_, _, w32 = run_oracle_loss(name)
_, _, w64 = run_oracle_loss(name, dtype=torch.float64)
for key in ("d_code", "d_code_pos"):
    print(key, "ours-vs-64", rel_err(r[key], w64[key]), "ref32-vs-64", rel_err(w32[key], w64[key]))
    e = [(rel_err(r[key][b], w64[key][b]), np.abs(r[key][b] - w64[key][b]).max()) for b in range(32)]
    print("  per image rel err:", " ".join(f"{x[0]:.1e}" for x in e))
    bad = int(np.argmax([x[0] for x in e]))
    d = np.abs(r[key][bad] - w64[key][bad]).sum(0)      # [28,28] error map
    print("  worst image", bad, "error map top pixels:", np.argsort(d.ravel())[-6:], np.sort(d.ravel())[-6:])
out = w64["out"]
for nm, cd in (("intra", out[1]), ("inter", out[3]), ("neg", out[5])):
    a = cd.detach().abs().flatten(1)
    print(nm, "min |cd| per image-pair (smallest 5):", np.sort(a.min(1)[0].numpy())[:5], " #|cd|<1e-6:", int((a < 1e-6).sum()), " #<1e-7:", int((a < 1e-7).sum()))
