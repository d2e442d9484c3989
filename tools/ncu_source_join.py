"""Per-source-line instruction counts of one kernel: joins `ncu --page source --csv` (SASS view: executed instructions
per SASS instruction) with `nvdisasm -g` line info of the same cubin (extracted from the built .so with
`cuobjdump -xelf all`).  Usage:
    python tools/ncu_source_join.py report.ncu-rep file.cubin kernel_substring [source_file_substring]
Prints the hottest source lines and the total; used for the instruction-share tables in DESIGN.md / profiles/."""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def sass_counts(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    cols = rows[hdr]
    ci, cs, cw = cols.index("Instructions Executed"), cols.index("Source"), cols.index("Warp Stall Sampling (All Samples)")
    return [(r[cs].strip(), int(r[ci] or 0), int(r[cw] or 0)) for r in rows[hdr + 1:] if len(r) > ci]


def line_info(cubin, kernel):
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    res, cur, active = [], None, False
    for ln in out.splitlines():
        if ln.startswith(".text."):
            active = kernel in ln
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        if active and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            res.append(cur)
    return res


def main():
    rep, cubin, kernel = sys.argv[1:4]
    want = sys.argv[4] if len(sys.argv) > 4 else None
    sc, li = sass_counts(rep), line_info(cubin, kernel)
    assert len(sc) == len(li), (len(sc), len(li))
    per, stall = defaultdict(int), defaultdict(int)
    for (src, n, w), loc in zip(sc, li):
        per[loc] += n
        stall[loc] += w
    tot, tots = sum(per.values()), sum(stall.values())
    print(f"total warp instructions {tot}, stall samples {tots}")
    files = {}
    for (loc, n) in sorted(per.items(), key=lambda kv: -kv[1])[:60]:
        if loc is None or (want and want not in loc[0]):
            print(f"{n:10d} {100 * n / tot:5.1f}%  {loc}")
            continue
        if loc[0] not in files:
            try:
                files[loc[0]] = open(loc[0]).read().splitlines()
            except OSError:
                files[loc[0]] = []
        src = files[loc[0]][loc[1] - 1].strip() if loc[1] - 1 < len(files[loc[0]]) else ""
        print(f"{n:10d} {100 * n / tot:5.1f}%  stall {100 * stall[loc] / max(tots, 1):5.1f}%  {loc[0].split('/')[-1]}:{loc[1]:4d}  {src[:110]}")
    return per


if __name__ == "__main__":
    main()
