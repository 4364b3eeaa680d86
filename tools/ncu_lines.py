#!/usr/bin/env python
"""tools/ncu_lines.py -- per-source-line view of an ncu report (needs -lineinfo and --import-source on).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_describe [top_n]

Joins `ncu --page source --csv` (per-SASS-instruction counters) with `nvdisasm --print-line-info` of the
in-tree library, and prints for each source line: share of executed warp instructions, stall samples,
average active threads, global L1 tag requests and shared-memory wavefronts.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "extractorb_b200", "libextractorb_cuda.so")
SRC = os.path.join(ROOT, "extractorb_b200", "csrc", "orbx_kernels.cuh")


def sass_line_map(kernel):
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=td, check=True, capture_output=True)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout
    lines = txt.split("\n")
    starts = [i for i, l in enumerate(lines) if l.startswith(".text.") and kernel in l]
    if not starts:
        raise SystemExit("kernel %s not found in %s" % (kernel, LIB))
    m, cur = {}, None
    for l in lines[starts[0] + 1:]:
        if l.startswith(".text.") or l.startswith("//--------------------- .text"):
            if m:
                break
        g = re.search(r'//## File "(.*?)", line (\d+)', l)
        if g:
            cur = (os.path.basename(g.group(1)), int(g.group(2)))
            continue
        g = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if g:
            m[int(g.group(1), 16)] = cur
    return m


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    mangled = kernel
    if ":" in kernel:
        kernel, mangled = kernel.split(":", 1)
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if "Address" in r and "Instructions Executed" in r][0]
    hdr = rows[hi]
    col = {n: hdr.index(n) for n in ("Address", "Instructions Executed", "Thread Instructions Executed", "# Samples",
                                     "L1 Tag Requests Global", "L1 Wavefronts Shared")}
    # template instances: "k_fast_tiles:k_fast_tilesILb1" = ncu kernel-name regex : fragment of the mangled name in the cubin
    amap = sass_line_map(mangled)
    agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
    base = None
    for r in rows[hi + 1:]:
        try:
            a = int(r[col["Address"]], 16)
            n = int(r[col["Instructions Executed"]])
        except (ValueError, IndexError):
            continue
        if base is None:
            base = a
        ln = amap.get(a - base)
        v = agg[ln]
        v[0] += n
        v[1] += int(r[col["Thread Instructions Executed"]] or 0)
        v[2] += int(r[col["# Samples"]] or 0)
        v[3] += int(r[col["L1 Tag Requests Global"]] or 0)
        v[4] += int(r[col["L1 Wavefronts Shared"]] or 0)
    tot = sum(v[0] for v in agg.values()) or 1
    smp = sum(v[2] for v in agg.values()) or 1
    srcs = {}
    for f in os.listdir(os.path.dirname(SRC)):
        try:
            srcs[f] = open(os.path.join(os.path.dirname(SRC), f)).read().split("\n")
        except (OSError, UnicodeDecodeError):
            pass
    print("total warp instructions %d, stall samples %d" % (tot, smp))
    print(" inst%  stall%  thr/inst  L1tagReq(M)  smemWave(M)  line  source")
    for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
        fname, lno = ln if ln else ("", 0)
        src = srcs.get(fname, [])
        text = src[lno - 1].strip()[:86] if lno and lno <= len(src) else ""
        tag = ("%s:%d" % (fname.replace("orbx_", "").replace(".cuh", ""), lno)) if lno else "?"
        print("%5.1f  %5.1f   %5.1f   %9.2f   %9.2f   %-12s %s" % (100.0 * v[0] / tot, 100.0 * v[2] / smp, v[1] / max(v[0], 1),
                                                                 v[3] / 1e6, v[4] / 1e6, tag, text))


if __name__ == "__main__":
    main()
