"""Native-layout AdaIN (K1n: statistics, coefficients, apply) at the bench shape (32, 64, 64, 512), whole batch in one
call and in image groups: with a group's content + style features (2 x 4.2 MB per image) inside the 126 MB L2, the
apply pass re-reads the content from L2 instead of HBM."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda")
eng = bench.build_engine(dev)
N = 32
fc = eng.buf.get("enc8c", N, 64, 64, 512, dev, True); fc.normal_()
fs = eng.buf.get("enc8s0", N, 64, 64, 512, dev, True); fs.normal_()


def run(group):
    for g0 in range(0, N, group):
        eng.adain(fc[g0:g0 + group], [fs[g0:g0 + group]], [1.0])


for group in (32, 16, 8, 4, 2):
    for _ in range(3): run(group)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): run(group)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 20 * 1e3
    print("native adain (32,64,64,512), %2d images per call: %.1f us per batch = %.0f GB/s of the 402 MB algorithmic" % (group, us, 402.65e6 / us / 1e3))
