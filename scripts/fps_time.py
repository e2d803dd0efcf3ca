"""FPS kernel alone: CUDA-event time per launch for several S (setup vs rounds), 2 x 32 images of 224 x 224."""
import sys, torch
sys.path.insert(0, ".")
from depthg_b200 import modules as M

torch.manual_seed(0)
B = 32
sets = [(torch.rand(B, 1, 224, 224, device="cuda"), torch.rand(B, 1, 224, 224, device="cuda")) for _ in range(12)]
for S in (2, 6, 11, 12, 16):
    for _ in range(5):
        M._fps(sets[0][0], sets[0][1], 28, 28, S, True, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 120
    e0.record()
    for i in range(n):
        d = sets[i % len(sets)]
        M._fps(d[0], d[1], 28, 28, S, True, False)
    e1.record()
    torch.cuda.synchronize()
    print(f"S={S:2d} rounds={S*S-1:3d}  {1e3 * e0.elapsed_time(e1) / n:7.2f} us per launch (back to back, {2*B} images)")
