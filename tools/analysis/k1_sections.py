#!/usr/bin/env python
"""Split an `ncu --page source --csv` export of k_pixels_to_tokens<true> into the kernel's stages (by the source line each
SASS instruction maps to): executed warp-instructions and warp-state samples (= warp time) per stage.
usage: k1_sections.py k1_src.csv k_tokens.o [tiles]      (line ranges below follow the current sources)"""
import collections, csv, glob, os, re, subprocess, sys, tempfile

SECTIONS = [  # (label, file, first line, last line)
    ("colour: numerators (IDP.2A)", "dct_core.cuh", 172, 193), ("colour: divisions, tie screens (ycc_row8n)", "dct_core.cuh", 194, 293),
    ("colour: loads, packing, chroma sums, tie list", "k_tokens.cu", 188, 238), ("colour: tie replay", "k_tokens.cu", 97, 187),
    ("colour: tie replay", "k_tokens.cu", 239, 269), ("colour: tie replay", "dct_core.cuh", 15, 56),
    ("DCT: AAN butterflies", "dct_core.cuh", 294, 366), ("DCT: quantisation brackets", "dct_core.cuh", 376, 392),
    ("DCT: zig-zag packing, mask", "dct_core.cuh", 367, 375), ("DCT: zig-zag packing, mask", "dct_core.cuh", 393, 409),
    ("DCT: unpack, DC chain, block glue", "dct_core.cuh", 410, 500),
    ("fetch (bulk copies, mbarrier)", "dct_core.cuh", 146, 171), ("fetch (bulk copies, mbarrier)", "k_tokens.cu", 301, 328),
    ("tile bookkeeping, barrier, histogram flush", "k_tokens.cu", 270, 300), ("tile bookkeeping, barrier, histogram flush", "k_tokens.cu", 329, 422),
    ("token stage: run/offset prefix, DC + EOB tokens", "k_tokens.cu", 423, 522), ("token stage: walk start (search, descent)", "k_tokens.cu", 523, 582),
    ("token stage: AC walk loop", "k_tokens.cu", 583, 615), ("token stage: flush, run records", "k_tokens.cu", 616, 640)]


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
    # -gi: every instruction is preceded by its inline chain, innermost first; an instruction is attributed to the innermost frame
    # that lies in this repo's sources (the packed-FP32 and atomic intrinsics live in CUDA headers)
    dis = subprocess.run(["nvdisasm", "-gi", "-c", glob.glob(os.path.join(tmp, "*.cubin"))[0]], capture_output=True, text=True).stdout
    OURS = ("dct_core.cuh", "k_tokens.cu", "walk.cuh", "jpegb200_internal.cuh")
    lines, cur, fn, in_chain, chain_done = [], None, None, False, False
    for line in dis.split("\n"):
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            ent = (os.path.basename(m.group(1)), int(m.group(2)))
            if not in_chain:
                cur, chain_done = ent, ent[0] in OURS
            elif not chain_done and ent[0] in OURS:
                cur, chain_done = ent, True
            in_chain = True
            continue
        in_chain = False
        m = re.match(r"\.text\.(\S+):", line)
        if m:
            fn, cur = m.group(1), None
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and fn and "k_pixels_to_tokensILb1ELb0" in fn:
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
