"""Weight gradient of every classic-decoder conv at the config-2 shapes (batch 8, 256x256 image), timed alone:
K2wn (csrc/wgrad_mn.cu: native NHWC operands, three kw taps per CTA) against the round-1 path (four channel-planar
copies + csrc/wgrad_tc.cu, one tap per work item).  Also the config-2 step from one CUDA graph (bench.time_train_step).
  python tools/bench_wgrad.py [--step]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arbitrarystyletransfer_b200 import _lib as L, engine as E, train_ops as T

dev = torch.device("cuda")
N, size = 8, 256


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


h = size // 8
tot_old = tot_new = 0.0
for i in range(9):
    cin, cout, relu, up = E.DECODER_SPEC[i] if i < 8 else (64, 3, False, False)
    H = W = h
    cz = 64 if cout == 3 else cout
    x = (torch.randn(N, H + 2, W + 2, cin, device=dev) * 0.5).to(torch.bfloat16)
    dz = torch.zeros(N, H + 4, W + 4, cz, device=dev, dtype=torch.bfloat16)
    dz[:, 2:-2, 2:-2, :cout] = (torch.randn(N, H, W, cout, device=dev) * 0.5).to(torch.bfloat16)
    w = torch.empty(cout, cin, 3, 3, device=dev)
    b = torch.empty(cout, device=dev)

    def old():
        dzT = T.to_planar(dz, N, cz, H, W, 2, False)
        xT = T.to_planar(x, N, cin, H, W, 1, True, nshift=3)
        return T.conv_wgrad(dzT, xT, N, H, W, cin, cout, w, b)

    def new():
        return T.conv_wgrad_native(dz, 2, x, N, H, W, cin, cout, w, b)

    go, gn = old(), new()
    rw = ((go[0] - gn[0]).norm() / go[0].norm()).item()
    rb = ((go[1] - gn[1]).norm() / go[1].norm()).item()
    to, tn = timeit(old), timeit(new)
    tot_old += to; tot_new += tn
    gf = 2.0 * N * H * W * cin * cout * 9 / 1e9
    print(f"dec_conv{i + 1}: {cin:3d} -> {cout:3d} @ {H:3d}x{W:<3d}  planar + K2w {to:7.1f} us   K2wn {tn:7.1f} us "
          f"({gf / tn * 1e3:6.1f} TFLOP/s)   new vs old: dW rel {rw:.1e}, db rel {rb:.1e}", flush=True)
    if up:
        h *= 2
print(f"sum over the nine layers: planar + K2w {tot_old:.0f} us, K2wn {tot_new:.0f} us")

if "--step" in sys.argv:
    import bench
    r = bench.time_train_step(dev, steps=20, warmup=3)
    print(json.dumps({k: r[k] for k in ("value", "ms_per_step", "eager_steps_per_s", "mode") if k in r}))
