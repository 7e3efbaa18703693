#!/usr/bin/env python
"""Per-kernel stall summary from `ncu --page source --csv` output (handles several kernels per file)."""
import csv
import subprocess
import sys


def blocks(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = dict(name=r[1], header=None, data=[])
            yield cur
        elif cur is not None and cur["header"] is None:
            cur["header"] = r
        elif cur is not None and len(r) == len(cur["header"]):
            cur["data"].append(r)


def main(path, top=14):
    for b in list(blocks(path)):
        h, data = b["header"], b["data"]
        isrc, isamp, iex, ia = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Address")
        cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        tot = sum(int(r[isamp] or 0) for r in data) or 1
        print(f"== {b['name'][:90]}  samples={tot} instr={sum(int(r[iex] or 0) for r in data)}")
        agg = {}
        for r in data:
            for i in cols:
                agg[h[i]] = agg.get(h[i], 0) + int(r[i] or 0)
        print("   " + "  ".join(f"{k[6:]}={v / tot:.1%}" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
        for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:top]:
            st = sorted(((h[i][6:], int(r[i] or 0)) for i in cols if int(r[i] or 0) > 0), key=lambda x: -x[1])[:2]
            print(f"   {data.index(r):5d} {int(r[isamp]):7d} {int(r[isamp]) / tot:6.1%} ex={r[iex]:>10s} {r[isrc][:58]:58s} {st}")
        marks = {k: [i for i, r in enumerate(data) if k in r[isrc]] for k in ["LDTM", "STTM", "I2F.S64", "UTCIMMA", "UBLKCP", "BAR.SYNC", "SYNCS.PHASECHK"]}
        print("   marks", {k: (v[0], v[-1], len(v)) for k, v in marks.items() if v}, "n_sass", len(data))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 14)
