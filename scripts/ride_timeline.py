"""Timeline of the correlation kernel with the next batch's sampling riding along: %globaltimer stamps per CTA."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, numpy as np
from depthg_b200 import modules as M, _lib
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
sets = [bench.synth_inputs(32, gen, dev) for _ in range(3)]
for s in sets:
    s["code"].requires_grad_(True); s["code_pos"].requires_grad_(True)
fn = M.ContrastiveCorrelationLoss(bench.make_cfg())
def step(i, ride):
    s, n = sets[i % 3], sets[(i + 1) % 3]
    s["code"].grad = None; s["code_pos"].grad = None
    if ride: fn.queue_next_sampling(n["depth"], n["depth_pos"])
    out = fn(s["feats"], s["feats_pos"], None, None, s["code"], s["code_pos"], s["depth"], s["depth_pos"])
    bench.backprop(out)
for ride in (False, True):
    for i in range(6): step(i, ride)
    torch.cuda.synchronize()
    clk = torch.zeros((256, 16), dtype=torch.int64, device=dev)
    _lib.lib().dg_debug_set_clock_buffer(clk.data_ptr())
    step(6, ride); torch.cuda.synchronize()
    _lib.lib().dg_debug_set_clock_buffer(None)
    c = clk.cpu().numpy().astype(np.float64)
    main = c[:148]; t0 = main[main[:, 1] > 0][:, 1].min()
    e = (main[:, 9] - t0) / 1e3
    print(f"ride={ride}: main CTAs end: one-item (76..147) median {np.median(e[76:]):.1f} us max {e[76:].max():.1f}; "
          f"two-item (0..75) median {np.median(e[:76]):.1f} max {e[:76].max():.1f}")
    r = c[148:214]; live = r[:, 9] > 0
    if live.any():
        st, en = (r[live][:, 1] - t0) / 1e3, (r[live][:, 9] - t0) / 1e3
        print(f"   ride CTAs: {int(live.sum())} ran; start min {st.min():.1f} median {np.median(st):.1f} max {st.max():.1f}; "
              f"duration median {np.median(en - st):.1f} max {(en - st).max():.1f}; end max {en.max():.1f}")
        img = r[live][:65]
        ph = [np.median((img[:, k] - img[:, 1]) / 1e3) for k in (3, 4)]
        print(f"   ride FPS phases (median, us after CTA start): pooled+lifted {ph[0]:.1f}  rounds done {ph[1]:.1f}")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(10): step(i, ride)
    torch.cuda.synchronize(); e0.record()
    for i in range(100): step(i, ride)
    e1.record(); torch.cuda.synchronize()
    print(f"   step {e0.elapsed_time(e1) * 10:.1f} us")
