#!/usr/bin/env python
"""Split an `ncu --page source --csv` export of k_pixels_to_tokens<true> into the kernel's stages (by the source line each
SASS instruction maps to): executed warp-instructions and warp-state samples (= warp time) per stage.
usage: k1_sections.py k1_src.csv k_tokens.o [tiles]      (line ranges below follow the current sources)"""
import collections, csv, glob, os, re, subprocess, sys, tempfile

SECTIONS = [  # (label, file, first line, last line)
    ("colour: numerators (IDP.2A)", "dct_core.cuh", 169, 190), ("colour: divisions, tie screens (ycc_row8n)", "dct_core.cuh", 191, 273),
    ("colour: loads, packing, chroma sums, tie list", "k_tokens.cu", 188, 236), ("colour: tie replay", "k_tokens.cu", 97, 187),
    ("colour: tie replay", "k_tokens.cu", 237, 267), ("colour: tie replay", "dct_core.cuh", 15, 54),
    ("DCT: AAN butterflies", "dct_core.cuh", 274, 294), ("DCT: quantisation brackets", "dct_core.cuh", 295, 321),
    ("DCT: zig-zag packing, mask", "dct_core.cuh", 322, 341), ("DCT: unpack, DC chain, block glue", "dct_core.cuh", 342, 400),
    ("fetch (bulk copies, mbarrier)", "dct_core.cuh", 143, 168), ("fetch (bulk copies, mbarrier)", "k_tokens.cu", 299, 323),
    ("tile bookkeeping, barrier, histogram flush", "k_tokens.cu", 268, 298), ("tile bookkeeping, barrier, histogram flush", "k_tokens.cu", 324, 417),
    ("token stage: run/offset prefix, DC + EOB tokens", "k_tokens.cu", 418, 503), ("token stage: walk start (search, descent)", "k_tokens.cu", 504, 563),
    ("token stage: AC walk loop", "k_tokens.cu", 564, 596), ("token stage: flush, run records", "k_tokens.cu", 597, 640)]


def label(f, ln):
    for lab, ff, a, b in SECTIONS:
        if f == ff and a <= ln <= b:
            return lab
    if f in ("jpegb200_internal.cuh",):
        return "token stage: AC walk loop" if ln < 90 else "tile bookkeeping, barrier, histogram flush"
    if f in ("device_atomic_functions.hpp",):
        return "token stage: AC walk loop"
    if f in ("math_functions.hpp", "sm_30_intrinsics.hpp", "sm_32_intrinsics.hpp", "device_functions.hpp"):
        return "intrinsic wrappers (unattributed)"
    return "other"


def main():
    src_csv, obj = sys.argv[1:3]
    tiles = float(sys.argv[3]) if len(sys.argv) > 3 else 38400.0
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    dis = subprocess.run(["nvdisasm", "-g", "-c", glob.glob(os.path.join(tmp, "*.cubin"))[0]], capture_output=True, text=True).stdout
    lines, cur, fn = [], None, None
    for line in dis.split("\n"):
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\.text\.(\S+):", line)
        if m:
            fn, cur = m.group(1), None
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and fn and "k_pixels_to_tokensILb1" in fn:
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
    agg = collections.OrderedDict((lab, [0, 0]) for lab, *_ in SECTIONS)
    for (e, s), ln in zip(data, lines):
        a = agg.setdefault(label(*ln) if ln else "other", [0, 0])
        a[0] += e
        a[1] += s
    te, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    print(f"| stage | warp-instr per tile | thread-instr per pixel | % instructions | % warp time (samples) |\n|---|---:|---:|---:|---:|")
    for k, a in agg.items():
        if a[0]:
            print(f"| {k} | {a[0] / tiles:.0f} | {a[0] / tiles * 32 / 4096:.1f} | {100 * a[0] / te:.1f} | {100 * a[1] / ts:.1f} |")
    print(f"| **total** | {te / tiles:.0f} | {te / tiles * 32 / 4096:.1f} | 100 | 100 ({ts} samples) |")


if __name__ == "__main__":
    main()
