"""Pointwise-conv (ast_pw_conv) calls of one AutoEncoder training step at batch N, 256x256: shapes as logged from the
model, each timed alone, with algorithmic GB/s against the HBM copy peak."""
import os, sys, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arbitrarystyletransfer_b200 import mobilenet as MB

N = int(os.environ.get("N", 32))
dev = torch.device("cuda")
calls = collections.OrderedDict()
orig = MB.pw_conv


def logged(x, w, bias, act, out_channels, residual=None, per_sample=False, want_raw=False, res_up2=False, f16=False):
    key = (tuple(x.shape), out_channels, int(act), residual is not None, bool(per_sample), bool(want_raw), bool(res_up2), bool(f16),
           bias is not None)
    calls[key] = calls.get(key, 0) + 1
    return orig(x, w, bias, act, out_channels, residual=residual, per_sample=per_sample, want_raw=want_raw, res_up2=res_up2, f16=f16)


MB.pw_conv = logged
torch.manual_seed(2)
ae = MB.AutoEncoder().to(dev).train()
x = torch.rand(N, 3, 256, 256, device=dev)
rec = ae(x)
rec.float().mean().backward()
torch.cuda.synchronize()
MB.pw_conv = orig
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6452.5) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6452.5
tot = 0.0
for key, cnt in calls.items():
    shape, cout, act, has_res, per_sample, want_raw, res_up2, f16, has_bias = key
    n, h, w_, cin = shape
    dt = torch.float16 if f16 else torch.bfloat16
    xx = torch.randn(n, h, w_, cin, device=dev).to(dt)
    ww = (torch.randn((n, cout, cin) if per_sample else (cout, cin), device=dev) * 0.1).to(dt)
    bb = torch.randn(cout, device=dev) if has_bias else None
    rs = None
    if has_res:
        rs = torch.randn(n, h // 2 if res_up2 else h, w_ // 2 if res_up2 else w_, cout, device=dev).to(dt)
    fn = lambda: orig(xx, ww, bb, act, cout, residual=rs, per_sample=per_sample, want_raw=want_raw, res_up2=res_up2, f16=f16)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    ev[0].record()
    for i in range(10):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))[5]
    px = n * h * w_
    by = px * 2 * (cin + cout * (2 if want_raw else 1) + (cout // (4 if res_up2 else 1) if has_res else 0))
    tot += ms * cnt
    print(f"{cnt:2d} x  px={px:8d} {cin:4d} -> {cout:4d} act={act} res={int(has_res)}{'u' if res_up2 else ' '} ps={int(per_sample)} raw={int(want_raw)} "
          f"{'f16' if f16 else 'bf16'}: {ms * 1e3:7.1f} us  {by / ms / 1e6:6.0f} GB/s = {by / ms / 1e6 / peak:4.2f} of HBM")
print(f"sum over the step's pw_conv calls: {tot:.2f} ms")
