"""Extract per-launch DRAM traffic (and a few headline metrics) of one kernel from an .ncu-rep into profiles/ncu_traffic.json.
usage: python tools/ncu_traffic.py <report.ncu-rep> <kernel substring> <key> [launch index]"""
import csv, io, json, os, subprocess, sys
rep, sub, key = sys.argv[1], sys.argv[2], sys.argv[3]
idx = int(sys.argv[4]) if len(sys.argv) > 4 else -1
out = (open(rep).read() if rep.endswith(".csv") else
       subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
sel = [r for r in rows[2:] if sub in r[hdr.index("Kernel Name")]]
r = sel[idx]
def val(name):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i]
    mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(u, 1.0)
    return v * mult
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
d = json.load(open(path)) if os.path.isfile(path) else {}
d[key] = {"kernel": r[hdr.index("Kernel Name")], "grid": r[hdr.index("Grid Size")],
          "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
          "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
          "duration_us_under_ncu": val("gpu__time_duration.sum"),
          "source": os.path.basename(rep) + " (ncu --set full --clock-control none)"}
json.dump(d, open(path, "w"), indent=1)
print(key, d[key])
