"""clock64 timeline of CTA 0 of one tensor-core conv launch (debugging aid)."""
import sys, math, torch
sys.path.insert(0, '.')
from physicsinformeddiffusionmodels_b200 import ops, packing
from physicsinformeddiffusionmodels_b200._lib import call
B, H, Cin, Cout, k = [int(a) for a in sys.argv[1:6]] if len(sys.argv) > 5 else (32, 64, 32, 32, 3)
dev = 'cuda'
x = torch.randn(B, H, H, Cin, device=dev).bfloat16()
w = torch.nn.Parameter(torch.randn(Cout, Cin, 1, k, k, device=dev) / math.sqrt(Cin * k * k))
b = torch.nn.Parameter(torch.zeros(Cout, device=dev))
spec = packing.ConvSpec(w, 'conv', k, k, 1, k // 2)
pk = packing.WeightPacker(); pk.add(spec); pk.refresh(torch.bfloat16)
trace = torch.zeros(8192, dtype=torch.int64, device=dev)
def graph_us(fn):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(20):
                fn()
        g.replay(); side.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(5):
            g.replay()
        e1.record(side); side.synchronize()
    return e0.elapsed_time(e1) * 1000 / 100
with torch.no_grad():
    print(f'conv B={B} {H}x{H} {Cin}->{Cout} k={k}: us per launch (graph): no stats',
          round(graph_us(lambda: ops.conv2d(x, w, b, spec)), 2), ' gn stats',
          round(graph_us(lambda: ops.conv2d(x, w, b, spec, gn_link={'groups': 8})), 2))
    if len(sys.argv) > 6:
        sys.exit(0)
    call('pidm_debug_set_trace', trace)
    ops.conv2d(x, w, b, spec, gn_link={'groups': 8})
    torch.cuda.synchronize()
    call('pidm_debug_set_trace', None)
t = trace.cpu().tolist()
t0 = min(v for v in t if v > 0)
print('MMA warp per tile: [start, acc_empty ok, first k-step done, all issued]')
for lt in range(9):
    row = t[1024 + lt * 4: 1024 + lt * 4 + 4]
    if row[0]: print(lt, [v - t0 for v in row])
print('epilogue per tile: [start wait, acc_full ok, done]')
for lt in range(9):
    row = t[2048 + lt * 4: 2048 + lt * 4 + 3]
    if row[0]: print(lt, [v - t0 for v in row])
print('epilogue chunk 0: [acc_full ok -> ld done, -> stored, -> stats done]')
for lt in range(9):
    a = t[2048 + lt * 4 + 1]; row = t[3072 + lt * 4: 3072 + lt * 4 + 3]
    if row[0]: print(lt, [row[0] - a, row[1] - row[0], row[2] - row[1]])
print('split-K: [phase A done, exchange barrier passed, phase B done]', [v - t0 for v in t[3072:3075] if v])
print('producer tile starts:', [t[i * 2] - t0 for i in range(9) if t[i * 2]])
print('MMA K-steps: [full-wait done, issued] relative; producer issue time')
for git in range(40):
    a, b2, c = t[4096 + git * 2], t[4096 + git * 2 + 1], t[6144 + git]
    if a: print(git, a - t0, b2 - t0, 'producer', (c - t0) if c else None)
