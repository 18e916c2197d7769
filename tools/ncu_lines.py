#!/usr/bin/env python
"""Attribute ncu's per-SASS-instruction warp-stall samples to CUDA source lines.

ncu's source page in CSV form carries SASS only; nvdisasm -gi knows the source line (with inlining) of every SASS
instruction. Both list the kernel's instructions in the same order, so they are joined by position.

  python tools/ncu_lines.py gpurun_out/x.ncu-rep object-pose-estimation_b200/libope_cuda.so icp_kernel [top]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so, kernel):
    """[(sass text, 'file:line' of the innermost frame, 'file:line' chain)] for the kernel's instructions, in order"""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, check=True)
    out = []
    for f in sorted(os.listdir(tmp)):
        txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, f)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                             text=True).stdout
        if kernel not in txt:
            continue
        in_k = False
        cur = "?"
        chain = "?"
        prev_was_loc = False
        for line in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+)", line)
            if m:
                in_k = kernel in m.group(1) and not out   # the first matching function only
                continue
            if not in_k:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
            if m:
                loc = "%s:%s" % (os.path.basename(m.group(1)), m.group(2))
                if prev_was_loc:          # outer frame of the same inline chain
                    chain += " <- " + loc
                else:                     # innermost frame of a new group
                    cur = loc
                    chain = loc
                prev_was_loc = True
                continue
            prev_was_loc = False
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                out.append((m.group(2).strip(), cur, chain))
        if out:
            break
    return out


def main():
    rep, so, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    # a report may hold several kernels: keep the launches of the one asked for (first of them)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + re.sub(r"ILi\d+.*", "", kernel), "--launch-count", "1"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr = rows[hi]
    nxt = next((i for i, r in enumerate(rows) if i > hi and "Source" in r and "# Samples" in r), len(rows))
    rows = rows[:nxt]
    si, ci, ii = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, c) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    sl = sass_lines(so, kernel)
    if len(sl) != len(body):
        print("# warning: %d SASS instructions in the report, %d in the cubin (stale build?)" % (len(body), len(sl)))
    agg = collections.OrderedDict()
    tot = 0.0
    for k, r in enumerate(body):
        n = float(r[ci] or 0)
        tot += n
        loc = sl[k][2] if k < len(sl) else "?"
        a = agg.setdefault(loc, [0.0, 0.0, collections.Counter()])
        a[0] += n
        a[1] += float(r[ii] or 0)
        for i, c in stall_cols:
            v = float(r[i] or 0)
            if v:
                a[2][c[6:]] += v
    print("# warp-stall samples by source line (innermost inlined frame <- call site), kernel %s, total %d samples" % (kernel, tot))
    print("%8s %6s %12s  %-60s %s" % ("samples", "share", "warp_inst", "where", "top stall reasons"))
    for loc, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        why = ", ".join("%s %.0f%%" % (k, 100 * v / max(sum(a[2].values()), 1)) for k, v in a[2].most_common(3))
        print("%8d %5.1f%% %12d  %-60s %s" % (a[0], 100 * a[0] / max(tot, 1), a[1], loc, why))


if __name__ == "__main__":
    main()
