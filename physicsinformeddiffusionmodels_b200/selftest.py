"""smoke(): one tiny PIDM training step on cuda:0, checked against the CPU oracle.  Called by
__graft_entry__.smoke(); the oracle is imported here as the CHECKER only."""
import os
import sys

import torch


def smoke():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import pidm_oracle as O
    from . import ops
    from .denoising_utils import DenoisingDiffusion
    from .residuals_darcy import ResidualsDarcy
    from .unet_model import Unet3D

    assert torch.cuda.is_available(), 'smoke() needs a CUDA device'
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    cfg = O.unet_config(dim=32, channels=2)
    sd = O.make_test_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(3)
    B = 2
    x0 = torch.randn(B, 2, 64, 64, generator=g)
    t = torch.tensor([5, 60])
    e = torch.randn(B, 2, 64, 64, generator=g)
    tables = O.diffusion_tables(100)
    sdr = {k: v.clone().requires_grad_('freqs' not in k) for k, v in sd.items()}
    loss_ref, aux = O.darcy_training_loss(sdr, cfg, x0, t, e, tables)
    loss_ref.backward()

    results = {}
    for mode, tol_out, tol_loss in (('fp32', 2e-4, 2e-4), ('bf16', 5e-2, 5e-2)):
        ops.set_precision(mode)
        model = Unet3D(dim=32, channels=2).to(dev)
        model.load_state_dict(sd)
        diff = DenoisingDiffusion(100, dev)
        res = ResidualsDarcy(model=model, fd_acc=2, pixels_per_dim=64, pixels_at_boundary=True, reverse_d1=True,
                             device=dev, bcs='none', domain_length=1.)
        loss, data_l, rabs, _, _ = diff.darcy_loss_from_draws(x0.to(dev), t.to(dev), e.to(dev), res, 1.0, 1e-3)
        loss.backward()
        torch.cuda.synchronize()
        gw = model.final_conv[1].weight.grad.float().cpu()
        gref = sdr['final_conv.1.weight'].grad
        err_l = abs(loss.item() / loss_ref.item() - 1)
        err_g = ((gw - gref).norm() / gref.norm()).item()
        results[mode] = (err_l, err_g)
        print(f'[smoke] {mode}: loss {loss.item():.6e} (oracle {loss_ref.item():.6e}, rel {err_l:.2e}); '
              f'grad(final_conv.1.weight) rel {err_g:.2e}; data {data_l:.4e} mean|r| {rabs:.4e}')
        assert err_l < tol_loss, f'{mode}: loss mismatch {err_l}'
        assert err_g < 10 * tol_out, f'{mode}: gradient mismatch {err_g}'
    ops.set_precision('bf16')
    print('[smoke] OK')
    return results
