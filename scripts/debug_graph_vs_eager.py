"""Debug aid: which parameters' gradients differ between the eager and the CUDA-graph training step."""
import os, sys, torch
sys.path.insert(0, '.')
from oracle import pidm_oracle as O
from physicsinformeddiffusionmodels_b200 import ops
from physicsinformeddiffusionmodels_b200.denoising_utils import DenoisingDiffusion
from physicsinformeddiffusionmodels_b200.engine import TrainEngine
from physicsinformeddiffusionmodels_b200.residuals_darcy import ResidualsDarcy
from physicsinformeddiffusionmodels_b200.unet_model import Unet3D
DEV = 'cuda'
mode = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
ops.set_precision(mode)
cfg = O.unet_config(dim=32, channels=2)
sd = O.make_test_state_dict(cfg, 0)
g = torch.Generator().manual_seed(500 + B)
x0 = (0.7 * torch.randn(B, 2, 64, 64, generator=g)).to(DEV)
t = torch.randint(0, 100, (B,), generator=g).to(DEV)
e = torch.randn(B, 2, 64, 64, generator=g).to(DEV)
out = {}
for tag, use_graph in (('eager', False), ('eager2', False), ('graph', True)):
    model = Unet3D(dim=32, channels=2).to(DEV)
    model.load_state_dict(sd)
    diff = DenoisingDiffusion(100, DEV)
    res = ResidualsDarcy(model=model, fd_acc=2, pixels_per_dim=64, pixels_at_boundary=True, reverse_d1=True, device=DEV)
    eng = TrainEngine(model, diff, res, use_graph=use_graph, snapshot_grad=True)
    o1, o2 = torch.randint, torch.randn_like
    torch.randint = lambda *a, **k: t
    torch.randn_like = lambda *a, **k: e
    try:
        loss, _, _ = eng.step(x0)
    finally:
        torch.randint, torch.randn_like = o1, o2
    torch.cuda.synchronize()
    names = {id(p): n for n, p in model.named_parameters()}
    grads = {names[id(p)]: eng.grad_snapshot[o:o + p.numel()].clone() for p, o in zip(eng.fp.params, eng.fp.offsets)}
    out[tag] = (loss.item(), grads)
for a, b in (('eager', 'eager2'), ('eager', 'graph')):
    print(f'== {a} vs {b}: loss {out[a][0]:.8e} {out[b][0]:.8e}')
    rows = []
    for n in out[a][1]:
        ga, gb = out[a][1][n].double(), out[b][1][n].double()
        r = ((ga - gb).norm() / ga.norm().clamp_min(1e-30)).item()
        rows.append((r, n, ga.norm().item()))
    rows.sort(reverse=True)
    print('   mismatching (>1e-3):', [n for r, n, _ in rows if r > 1e-3])
    for r, n, nn_ in rows[:3]:
        print(f'   {r:.3e}  {n}  |ga|={nn_:.3e} |gb|={out[b][1][n].double().norm().item():.3e}')
