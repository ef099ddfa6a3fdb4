#!/usr/bin/env python
"""profiles/traffic.json from ncu captures: DRAM bytes per launch of the dominant kernels, keyed by
kernel and by the sha256 of the very source file the capture was taken from (bench.py only quotes
`roofline.traffic` when the hash matches the source it is running).
usage: tools/make_traffic.py <rep> <kernel-key> <source.cu> key=value ..."""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, key, src = sys.argv[1:4]
    extra = dict(kv.split("=") for kv in sys.argv[4:])
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    vals = rows[2]
    units = rows[1]

    def get(name):
        i = hdr.index(name)
        v = float(vals[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    dram = get("dram__bytes_read.sum") + get("dram__bytes_write.sum")
    with open(os.path.join(ROOT, "friture_b200", "csrc", src), "rb") as f:
        sha = hashlib.sha256(f.read()).hexdigest()[:16]
    path = os.path.join(ROOT, "profiles", "traffic.json")
    data = json.load(open(path)) if os.path.isfile(path) else {}
    entry = {"kernel": vals[hdr.index("Kernel Name")], "source": src, "source_sha16": sha,
             "dram_bytes_per_launch": dram, "capture": os.path.basename(rep)}
    for k, v in extra.items():
        entry[k] = int(v)
    data[key] = entry
    json.dump(data, open(path, "w"), indent=1)
    print(json.dumps(entry))


if __name__ == "__main__":
    main()
