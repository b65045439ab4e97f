"""Average DRAM traffic per launch over a kernel FAMILY (several substrings) of an `ncu --page raw --csv` file, into
profiles/ncu_traffic.json.  usage: python tools/ncu_family_traffic.py <raw.csv> <key> <source note> <substr> [<substr> ...]"""
import csv, json, os, sys
raw, key, note, subs = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4:]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
def val(r, name):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(units[i], 1.0)
sel = [r for r in rows[2:] if any(s in r[hdr.index("Kernel Name")] for s in subs)]
rd = sum(val(r, "dram__bytes_read.sum") for r in sel)
wr = sum(val(r, "dram__bytes_write.sum") for r in sel)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
d = json.load(open(path)) if os.path.isfile(path) else {}
d[key] = {"kernel": " + ".join(subs) + f" ({len(sel)} launches of one stylise pass)", "launches": len(sel),
          "dram_bytes_per_launch": (rd + wr) / len(sel), "dram_read_bytes": rd / len(sel), "dram_write_bytes": wr / len(sel),
          "source": note}
json.dump(d, open(path, "w"), indent=1)
print(key, d[key])
