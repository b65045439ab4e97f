"""Diagnostic for the tcgen05 conv kernel: one case per process, prints error structure so that a
descriptor / swizzle / pipeline bug can be located from a single GPU run."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from tests.gpu_util import bf16r, oracle_conv_native, native_to_padded_nchw


def main():
    N, H, W, cin, cout, epi, reflect, impl = [int(a) for a in sys.argv[1:9]]
    from arbitrarystyletransfer_b200 import _lib as L, engine as E
    g = torch.Generator().manual_seed(0)
    x = bf16r(torch.randn(N, cin, H, W, generator=g))
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    exp, _, _ = oracle_conv_native(x, w, b, True, epi, "reflect" if reflect else "zeros")
    xin = E.nchw_to_native(x.cuda(), reflect=bool(reflect))
    wpk = E.pack_conv_weight(w.cuda())
    Ho, Wo = (H // 2, W // 2) if epi == 1 else ((2 * H, 2 * W) if epi == 2 else (H, W))
    res = {}
    for name, im in (("tc", impl), ("direct", L.CONV_DIRECT)):
        out = torch.zeros(N, Ho + 2, Wo + 2, cout, device="cuda", dtype=torch.bfloat16)
        E.conv3x3(xin, wpk, b.cuda(), out, N=N, H=H, W=W, cin=cin, cout=cout, relu=True, epilogue=epi,
                  halo=L.HALO_REFLECT if reflect else L.HALO_KEEP, impl=im)
        torch.cuda.synchronize()
        res[name] = native_to_padded_nchw(out).cpu()[:, :, 1:-1, 1:-1]
    for name, got in res.items():
        d = (got - exp)
        rel = d.norm() / exp.norm()
        print(f"[{name}] rel_l2={rel:.3e} max_abs={d.abs().max():.3e} exp_absmax={exp.abs().max():.3f} "
              f"nonfinite={(~torch.isfinite(got)).sum().item()} zeros_frac={(got == 0).float().mean():.3f} "
              f"(exp zeros {(exp == 0).float().mean():.3f})")
    d = (res["tc"] - exp).abs()
    if d.max() > 0.05:
        print("per 64-channel block max err:", [round(d[:, c:c + 64].max().item(), 3) for c in range(0, cout, 64)])
        print("per 16-channel block (first 128) :", [round(d[:, c:c + 16].max().item(), 3) for c in range(0, min(cout, 128), 16)])
        if epi == 0:
            print("per row (h) max err:", [round(d[:, :, h].max().item(), 3) for h in range(min(H, 16))])
            print("per col (w) max err:", [round(d[:, :, :, ww].max().item(), 3) for ww in range(min(W, 32))])
        print("per image max err:", [round(d[n].max().item(), 3) for n in range(N)])
        # is the tc output a channel / pixel permutation of the expectation?
        a, e = res["tc"][0].flatten(1), exp[0].flatten(1)
        if a.shape == e.shape:
            cc = torch.corrcoef(torch.cat([a[:8], e[:16]], 0))[:8, 8:]
            print("corr(tc ch0..7, exp ch0..15) argmax:", cc.argmax(1).tolist(), [round(v, 2) for v in cc.max(1).values.tolist()])
    print("DONE")


if __name__ == "__main__":
    main()
