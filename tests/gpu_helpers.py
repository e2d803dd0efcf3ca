"""Run the CUDA path on a golden case (mirror of helpers.run_oracle_loss)."""
import numpy as np
import torch

from tests.golden import cases
from tests.helpers import weighted_total


def run_cuda_loss(name, channels_last=False, materialize=False, inputs=None):
    from depthg_b200.modules import ContrastiveCorrelationLoss
    cfg, t = inputs if inputs is not None else cases.make_loss_inputs(name)
    dev = torch.device("cuda:0")

    def put(x):
        x = x.to(dev)
        if channels_last:  # the layout the live trainer produces (hard part 6)
            x = x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        return x

    feats, feats_pos = put(t["feats"]), put(t["feats_pos"])
    code = put(t["code"]).detach().requires_grad_(True)
    code_pos = put(t["code_pos"]).detach().requires_grad_(True)
    depth, depth_pos = t["depth"].to(dev), t["depth_pos"].to(dev)
    fn = ContrastiveCorrelationLoss(cfg, materialize_cd=materialize)
    perm_it = iter(t["perms"].to(dev))
    rand_it = iter([t["rand1"].to(dev), t["rand2"].to(dev)])
    fn.perm_fn = lambda B, device: next(perm_it).clone()
    fn.rand_fn = lambda shape, device: next(rand_it).clone()
    out = fn(feats, feats_pos, None, None, code, code_pos, depth, depth_pos)
    depth_term = cfg.depth_feat_correlation_loss
    assert len(out) == (8 if depth_term else 6)
    L = weighted_total(out, depth_term)
    L.backward()
    torch.cuda.synchronize()
    res = dict(coords1=fn.last_coords[0].cpu().numpy(), coords2=fn.last_coords[1].cpu().numpy(),
               scalars=np.array([out[0].item(), out[2].item(), out[4].mean().item(),
                                 out[6].item() if depth_term else np.nan]),
               cd_means=np.array([out[1].mean().item(), out[3].mean().item(), out[5].mean().item(),
                                  out[7].mean().item() if depth_term else np.nan]),
               total=L.item(), d_code=code.grad.cpu().numpy(), d_code_pos=code_pos.grad.cpu().numpy(), out=out,
               grad_strides=(code.grad.stride(), code.stride()))
    return cfg, t, res
