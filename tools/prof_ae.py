"""BASELINE config 3 (train_autoencoder.py:111-148 step on the MobileNet-style AutoEncoder) on one GPU:
eager and CUDA-graph timings, eval-mode inference timing, or (--profile) one step inside a profiler range
for the ncu launch list."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arbitrarystyletransfer_b200 import models as M, mobilenet as MB, losses as Ls

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--profile", action="store_true")
ap.add_argument("--no-graph", action="store_true")
ap.add_argument("--profile-eval", action="store_true", help="one eval-mode forward inside the profiler range")
args = ap.parse_args()
dev = torch.device("cuda")
x = torch.rand(args.batch, 3, args.size, args.size, generator=torch.Generator().manual_seed(301)).to(dev)


def build():
    torch.manual_seed(0)
    enc = M.PretrainedEncoder().to(dev).eval()
    M.calibrate_encoder_bias(enc, n_convs=16)
    for p in enc.parameters():
        p.requires_grad_(False)
    torch.manual_seed(2)
    ae = MB.AutoEncoder().to(dev).train()
    opt = torch.optim.Adam(ae.parameters(), lr=2e-4, betas=(0.9, 0.99), eps=1e-7, capturable=True, fused=True)

    def step(x):
        recon = ae(x)
        opt.zero_grad(set_to_none=True)
        recon_loss = Ls.compute_content_loss(recon, x)
        with torch.no_grad():
            cm = enc(x)
        rm = enc(recon)
        perp = None
        for a, b in zip(rm, cm):
            l = Ls.compute_content_loss(a, b)
            perp = l if perp is None else perp + l
        loss = 100.0 * recon_loss + 0.01 * perp
        loss.backward()
        torch.nn.utils.clip_grad_norm_(ae.parameters(), 10.0)
        opt.step()
        return loss
    return step, ae


def timeit(fn, steps):
    for _ in range(2):
        out = fn(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        out = fn(x)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps, out


if args.profile_eval:
    step, ae = build()
    ae.eval()
    with torch.no_grad():
        for _ in range(2):
            ae(x)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        ae(x)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print("ok")
    sys.exit(0)

if args.profile:
    step, ae = build()
    for _ in range(2):
        step(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("ok")
    sys.exit(0)

res = {"batch": args.batch, "size": args.size}
if not args.no_graph:
    try:
        from arbitrarystyletransfer_b200.graphs import GraphedStep
        step, ae = build()
        g = GraphedStep(step, [x.clone()])
        ms, loss = timeit(g, args.steps)
        res["graph_ms_per_step"] = ms
        res["graph_loss"] = loss.item()
        del g, step, ae
    except Exception as e:
        res["graph_error"] = repr(e)[:300]
    torch.cuda.empty_cache()
step, ae = build()
ms, loss = timeit(step, args.steps)
res["eager_ms_per_step"] = ms
res["eager_loss"] = loss.item()
res["max_mem_gb"] = torch.cuda.max_memory_allocated() / 2**30
ae.eval()
with torch.no_grad():
    ms, out = timeit(ae, args.steps)
res["eval_forward_ms"] = ms
res["eval_img_per_s"] = args.batch * 1e3 / ms
print(json.dumps(res))
