"""Per-call device time of every distinct libpidm call of one training step, measured by replaying each call from a
CUDA graph (20 launches per replay), i.e. without host launch overhead.  Debugging / optimisation aid.

    python scripts/layer_times.py [name-filter]
"""
import sys, collections, torch
sys.path.insert(0, '.')
from physicsinformeddiffusionmodels_b200 import _lib, ops, packing, denoising_utils, engine as eng_mod
from physicsinformeddiffusionmodels_b200.denoising_utils import DenoisingDiffusion
from physicsinformeddiffusionmodels_b200.engine import TrainEngine
from physicsinformeddiffusionmodels_b200.residuals_darcy import ResidualsDarcy
from physicsinformeddiffusionmodels_b200.unet_model import Unet3D

flt = sys.argv[1] if len(sys.argv) > 1 else ''
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device('cuda', 0)
ops.set_precision('bf16')
torch.manual_seed(0)
model = Unet3D(dim=32, channels=2).to(dev)
diff = DenoisingDiffusion(100, dev)
res = ResidualsDarcy(model=model, fd_acc=2, pixels_per_dim=64, pixels_at_boundary=True, reverse_d1=True, device=dev,
                     bcs='none', domain_length=1.)
eng = TrainEngine(model, diff, res, lr=1e-4, max_norm=1.0, ema_mu=0.99, c_data=1.0, c_residual=1e-3, use_graph=False,
                  world=1)
x0 = torch.randn(B, 2, 64, 64, device=dev)
for _ in range(2):
    eng.step(x0)
torch.cuda.synchronize()

records = []
orig = _lib.call


def rec(name, *a):
    if name not in _lib._VALUE_RETURN:
        records.append((name, a))
    return orig(name, *a)


mods = [ops, packing, denoising_utils, eng_mod]
for mo in mods:
    mo.call = rec
eng._step_body(x0)
torch.cuda.synchronize()
for mo in mods:
    mo.call = orig

SKIP = ('pidm_adam_ema_step',)
groups = collections.OrderedDict()
for name, a in records:
    if name in SKIP or flt not in name:
        continue
    key = (name,) + tuple(x for x in a if isinstance(x, int) and not isinstance(x, bool) and x < (1 << 24))
    groups.setdefault(key, []).append(a)

side = torch.cuda.Stream()
rows = []
for key, lst in groups.items():
    name, a = key[0], lst[0]
    # the stream argument is the last int >= 2^24 or 0: rebuild the arg list with the side stream's handle
    a = list(a)
    a[-1] = side.cuda_stream
    try:
        with torch.cuda.stream(side):
            for _ in range(2):
                orig(name, *a)
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(20):
                    orig(name, *a)
            g.replay(); side.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(side)
            for _ in range(5):
                g.replay()
            e1.record(side)
            side.synchronize()
            us = e0.elapsed_time(e1) * 1000 / 100
    except Exception as ex:   # noqa
        print('skip', key, ex)
        continue
    rows.append((us * len(lst), us, len(lst), key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f'total {tot / 1000:.3f} ms over {sum(r[2] for r in rows)} calls')
agg = collections.defaultdict(float)
for t, us, n, key in rows:
    agg[key[0]] += t
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f'  {k:34s} {v / 1000:8.3f} ms')
for t, us, n, key in rows[:70]:
    print(f'{t:9.1f} us = {n:3d} x {us:8.2f}  {key[0][5:]:24s} {list(key[1:])}')
