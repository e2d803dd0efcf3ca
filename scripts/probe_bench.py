"""Time the fused probe losses against the reference's stock-torch op sequence on the same GPU
(cfg2 shapes: B=32, D=90, 28x28 code, 27 classes, 224x224 labels).  Prints one JSON line."""
import json
import sys
import os

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from depthg_b200 import _lib  # noqa: E402
from depthg_b200.probes import ClusterLookup, linear_probe_loss  # noqa: E402


def timeit(fn, n=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    dev = torch.device("cuda:0")
    B, D, K, h, Hl = 32, 90, 27, 28, 224
    g = torch.Generator(device=dev).manual_seed(0)
    code = torch.randn(B, D, h, h, device=dev, generator=g).contiguous(memory_format=torch.channels_last)
    label = torch.randint(-1, K, (B, Hl, Hl), device=dev, generator=g)
    weight = (torch.randn(K, D, 1, 1, device=dev, generator=g) / D ** 0.5).requires_grad_(True)
    bias = torch.zeros(K, device=dev, requires_grad=True)
    probe = ClusterLookup(D, K).to(dev)

    def fused_linear():
        weight.grad = bias.grad = None
        linear_probe_loss(code, weight, bias, label).backward()

    def torch_linear():   # src/train_segmentation.py:419-437 as written
        weight.grad = bias.grad = None
        flat = label.reshape(-1)
        mask = (flat >= 0) & (flat < K)
        lg = F.conv2d(torch.clone(code.detach()), weight, bias)
        lg = F.interpolate(lg, label.shape[-2:], mode="bilinear", align_corners=False)
        lg = lg.permute(0, 2, 3, 1).reshape(-1, K)
        F.cross_entropy(lg[mask], flat[mask]).mean().backward()

    def fused_cluster():
        probe.clusters.grad = None
        probe(code, None)[0].backward()

    def torch_cluster():  # src/modules.py:659-675 as written
        probe.clusters.grad = None
        nc = F.normalize(probe.clusters, dim=1)
        nf = F.normalize(code, dim=1)
        ip = torch.einsum("bchw,nc->bnhw", nf, nc)
        cp = F.one_hot(torch.argmax(ip, dim=1), K).permute(0, 3, 1, 2).to(torch.float32)
        (-(cp * ip).sum(1).mean()).backward()

    lib = _lib.lib()
    res = {}
    for name, fn in (("linear_fused", fused_linear), ("linear_torch", torch_linear), ("cluster_fused", fused_cluster),
                     ("cluster_torch", torch_cluster)):
        res[name + "_us"] = round(timeit(fn), 1)
    lib.dg_profile_enable(1)
    fused_linear(); fused_cluster()
    torch.cuda.synchronize()
    buf = bytes(4096)
    import ctypes
    cbuf = ctypes.create_string_buffer(4096)
    lib.dg_profile_collect(cbuf, 4096)
    res["kernels"] = cbuf.value.decode()
    lib.dg_profile_enable(0)
    res["label_bytes"] = label.numel() * 8
    print(json.dumps(res))


if __name__ == "__main__":
    main()
