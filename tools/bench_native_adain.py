import os, sys, torch
sys.path.insert(0, "/root/repo")
import bench
dev = torch.device("cuda")
eng = bench.build_engine(dev)
N = 32
fc = eng.buf.get("enc8c", N, 64, 64, 512, dev, True); fc.normal_()
fs = eng.buf.get("enc8s0", N, 64, 64, 512, dev, True); fs.normal_()
for _ in range(3): eng.adain(fc, [fs], [1.0])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): eng.adain(fc, [fs], [1.0])
b.record(); torch.cuda.synchronize()
print("native adain (32,64,64,512) per call: %.1f us" % (a.elapsed_time(b) / 20 * 1e3))
