"""torchrun --nproc-per-node 2 scripts/check_bucket_ar.py : the bucketed (overlapped) gradient all-reduce must give the
same parameters as the single all-reduce after backward (same init, data and noise), eager and CUDA-graph."""
import os, sys, torch
import torch.distributed as dist
sys.path.insert(0, '.')
from physicsinformeddiffusionmodels_b200 import ops
from physicsinformeddiffusionmodels_b200.denoising_utils import DenoisingDiffusion
from physicsinformeddiffusionmodels_b200.engine import TrainEngine
from physicsinformeddiffusionmodels_b200.residuals_darcy import ResidualsDarcy
from physicsinformeddiffusionmodels_b200.unet_model import Unet3D

rank, world, lr_ = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr_)
dev = torch.device('cuda', lr_)
dist.init_process_group('nccl', device_id=dev)
ops.set_precision('bf16')


def run(bucketed, use_graph):
    torch.manual_seed(0)
    model = Unet3D(dim=32, channels=2).to(dev)
    diff = DenoisingDiffusion(100, dev)
    res = ResidualsDarcy(model=model, fd_acc=2, pixels_per_dim=64, pixels_at_boundary=True, reverse_d1=True, device=dev,
                         bcs='none', domain_length=1.)
    eng = TrainEngine(model, diff, res, use_graph=use_graph, world=world, bucketed_allreduce=bucketed)
    torch.manual_seed(100 + rank)
    x0 = torch.randn(8, 2, 64, 64, device=dev)
    torch.manual_seed(200 + rank)          # noise / t draws
    for _ in range(4):
        eng.step(x0)
    torch.cuda.synchronize()
    named = {n: p.detach().float().clone() for n, p in model.named_parameters()}
    return named, eng


ok = True
for use_graph in (False, True):
    a, _ = run(False, use_graph)
    b, eng = run(True, use_graph)
    worst = 0.0
    for n in a:
        d = (a[n] - b[n]).norm().item() / max(a[n].norm().item(), 1e-12)
        worst = max(worst, d)
    # ranks must also agree with each other
    flat = torch.cat([b[n].reshape(-1) for n in sorted(b)])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    across = (flat - ref).abs().max().item()
    if rank == 0:
        print(f'graph={use_graph}: bucketed vs single all-reduce worst rel diff {worst:.3e}; groups {eng.fp.group_bounds}; '
              f'reduced early {sorted(eng._reduced)}; max |rank diff| {across:.3e}', flush=True)
    ok = ok and worst < 2e-2 and across == 0.0
if rank == 0:
    print('BUCKET_AR_OK' if ok else 'BUCKET_AR_MISMATCH', flush=True)
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
