"""Isolated timing of the step's conv layers at the bench shape (batch 32, 512x512), for A/B experiments with the
AST_CONV_* environment switches.  usage: python tools/bench_conv.py [layer ...]   (names: enc2 .. enc9, dec1 .. dec8)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arbitrarystyletransfer_b200 import _lib as L, engine as E

N, S = int(os.environ.get("BENCH_N", 32)), 512
dev = torch.device("cuda", 0)
torch.manual_seed(0)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


layers = {}
h = S
for i, (cin, cout, pool) in enumerate(E.vgg_layer_plan(9)):
    if i > 0:
        layers[f"enc{i + 1}"] = (cin, cout, h, L.EPI_POOL2 if pool else L.EPI_PLAIN, L.HALO_KEEP, False)
    if pool:
        h //= 2
for i, (cin, cout, relu, up) in enumerate(E.DECODER_SPEC[:8]):
    folded_in = i in E.FOLD_LAYERS
    folded_out = up and (i + 1) in E.FOLD_LAYERS
    if folded_in:
        layers[f"dec{i + 1}"] = (cin, cout, h, L.EPI_UPFOLD, L.HALO_CLAMP if folded_out else L.HALO_REFLECT, True)
        h *= 2
    else:
        layers[f"dec{i + 1}"] = (cin, cout, h, L.EPI_PLAIN, L.HALO_CLAMP if folded_out else L.HALO_REFLECT, False)
        if up and not folded_out:
            h *= 2
want = sys.argv[1:] or list(layers)
for name in want:
    cin, cout, h, epi, halo, fold = layers[name]
    x = torch.randn(N, h + 2, h + 2, cin, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device=dev) * (2.0 / (9 * cin)) ** 0.5
    wpk = E.pack_conv_weight_fold(w) if fold else E.pack_conv_weight(w)
    b = torch.zeros(cout, device=dev)
    ho = h // 2 if epi == L.EPI_POOL2 else (2 * h if epi == L.EPI_UPFOLD else h)
    y = torch.empty(N, ho + 2, ho + 2, cout, device=dev, dtype=torch.bfloat16)
    ms = timed(lambda: E.conv3x3(x, wpk, b, y, N=N, H=h, W=h, cin=cin, cout=cout, relu=True, epilogue=epi, halo=halo))
    hout = 2 * h if fold else h
    fl = 2.0 * cout * cin * 9 * hout * hout * N
    fx = fl * (16 / 36 if fold else 1)
    print(f"{name}: {cin}->{cout} @{h} epi={epi} halo={halo}: {ms * 1e3:.1f} us  algorithmic {fl / ms / 1e9:.0f} TFLOP/s  executed {fx / ms / 1e9:.0f} TFLOP/s", flush=True)
