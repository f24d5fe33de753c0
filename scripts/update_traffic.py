"""profiles/sweep_traffic.json (and mg_traffic.json) from ncu reports of the CURRENT sources: bench.py reports roofline.traffic
only when the recorded source hash equals the hash of the library it runs.
usage: python scripts/update_traffic.py <sweep.ncu-rep> [<mg.ncu-rep>]"""
import csv, io, json, subprocess, sys, time
sys.path.insert(0, ".")
from pyro2_b200 import _lib


def dram(rep, want):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        if want in r[hdr.index("Kernel Name")]:
            tot = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v, u = float(r[hdr.index(k)]), units[hdr.index(k)]
                tot += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            return tot, float(r[hdr.index("gpu__time_duration.sum")])
    raise SystemExit(f"no {want} in {rep}")


b, ms = dram(sys.argv[1], "sweep_kernel")
rec = {"kernel": "pyro::sweep_kernel<0,0,0>", "dram_bytes_per_launch": b, "grid": "4096^2 Sedov", "ncu_duration_ms": ms,
       "source_hash": _lib.source_hash(), "captured": time.strftime("%Y-%m-%d"), "report": sys.argv[1]}
json.dump(rec, open("profiles/sweep_traffic.json", "w"), indent=1)
print(rec)
