#!/usr/bin/env python
"""Aggregate the warp-stall samples of an `ncu --page source --csv` export: totals per stall reason, and the
instructions with the most samples.  usage: ncu_stalls.py X_src.csv [top_n]"""
import collections, csv, sys

def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    hdr = rows[1]
    ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    data = []
    for r in rows[2:]:
        try:
            data.append((int(r[isamp]), int(r[ie]), r[ia].strip(), r))
        except (ValueError, IndexError):
            continue
        for i in stall_cols:
            try:
                tot[hdr[i]] += int(r[i])
            except ValueError:
                pass
    S = sum(tot.values())
    print("SASS instructions:", len(data), " executed (warp):", sum(d[1] for d in data), " samples:", S)
    for k, v in tot.most_common(9):
        print(f"  {k:26s} {v:7d} {100 * v / S:5.1f}%")
    for s, e, src, r in sorted(data, key=lambda d: -d[0])[:top_n]:
        st = {hdr[i]: int(r[i]) for i in stall_cols if r[i] not in ("", "0")}
        top = ", ".join(f"{k[6:]} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2])
        print(f"{s:6d} {100 * s / S:5.1f}% x{e:8d}  {src[:64]:64s} {top}")

if __name__ == "__main__":
    main()
