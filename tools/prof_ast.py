"""SURVEY section 8 row f1 on one GPU: the train.py:189-300 step on the AdaAttN network (bench.time_train_ast), or
(--profile) one eager step -- or (--layer) one AdaAttN layer forward + backward at the network's shape -- inside a
profiler range for the ncu launch list."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--profile", action="store_true")
ap.add_argument("--layer", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda")

if args.layer:
    from arbitrarystyletransfer_b200 import attention as AT
    torch.manual_seed(1)
    layer = AT.AdaAttN(128).to(dev)
    h = args.size // 8
    fc = torch.randn(args.batch, 128, h, h, device=dev).requires_grad_(True)
    fs = (torch.randn(args.batch, 128, h, h, device=dev) * 2 + 1).requires_grad_(True)
    for i in range(3):
        if i == 2:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        y = layer(fc, fs)
        y.backward(fc.detach())
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("layer ok", float(y.abs().mean()))
elif args.profile:
    # two warm eager steps, then one inside the profiler range
    g = torch.Generator().manual_seed(801)
    c = torch.rand(args.batch, 3, args.size, args.size, generator=g).to(dev)
    s = torch.rand(args.batch, 3, args.size, args.size, generator=g).to(dev)
    step = bench.build_ast_step(dev, c, s)[0]
    for _ in range(2):
        step(c, s)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    loss = step(c, s)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("loss", float(loss))
else:
    print(json.dumps(bench.time_train_ast(dev, batch=args.batch, size=args.size)))
