#!/usr/bin/env python
"""Join an `ncu --page source --csv` export (per-SASS-instruction executed counts and stall samples) with the line table
of the object file (nvdisasm -g): executed warp instructions and samples per source line bucket.
usage: ncu_by_line.py X_src.csv object.o kernel_substring [bucket]"""
import collections, csv, re, subprocess, sys, tempfile, os, glob

def main():
    src_csv, obj, kname = sys.argv[1:4]
    bucket = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    lines, cur, fn = [], None, None
    for line in dis.split("\n"):
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\.text\.(\S+):", line)
        if m:
            fn, cur = m.group(1), None
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and fn and kname in fn:
            lines.append(cur)
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = []
    for r in rows[2:]:
        try:
            data.append((int(r[ie]), int(r[isamp])))
        except (ValueError, IndexError):
            pass
    if len(data) != len(lines):
        print(f"warning: {len(data)} profiled instructions vs {len(lines)} disassembled", file=sys.stderr)
    agg = collections.OrderedDict()
    for (e, s), ln in zip(data, lines):
        key = (ln[0], ln[1] // bucket * bucket) if ln else ("?", 0)
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += e; a[1] += s; a[2] += 1
    te, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    print(f"{'file':24s} {'line':>5s} {'sass':>5s} {'exec(warp)':>12s} {'%':>6s} {'samples':>8s} {'%':>6s}")
    for k, a in sorted(agg.items()):
        if a[0] * 200 > te or a[1] * 200 > ts:
            print(f"{k[0]:24s} {k[1]:5d} {a[2]:5d} {a[0]:12d} {100 * a[0] / te:6.1f} {a[1]:8d} {100 * a[1] / ts:6.1f}")

if __name__ == "__main__":
    main()
