"""Where do the warp-stall samples of one profiled launch go?  SASS-level view of an ncu report (robust: every instruction
is listed once), grouped by instruction class, plus the stall-reason totals and the hottest instructions with context.
usage: python tools/ncu_hot_lines.py report.ncu-rep <launch index in the report> [top N instructions]
(The combined sass+cuda source page lists an instruction once per inlining level, so per-source-line sums over it
over-count; this tool does not use it.)"""
import collections
import csv
import io
import subprocess
import sys


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


def classify(op, prev_src):
    if op.startswith("BAR"):
        return "named / CTA barrier (BAR)"
    if op.startswith("SYNCS") or (op == "BRA" and "SYNCS" in prev_src):
        return "mbarrier try_wait loop"
    if op.startswith(("UTCHMMA", "UTCBAR", "R2UR", "ELECT")):
        return "MMA issue / commit"
    if op.startswith(("LDTM", "STTM")):
        return "tcgen05.ld / tcgen05.st"
    if op.startswith("SHFL"):
        return "shuffles"
    if op.startswith(("LDG", "STG", "ATOMG", "RED")):
        return "global memory"
    if op.startswith(("LDS", "STS", "ATOMS", "LDL", "STL")):
        return "shared / local memory"
    if op.startswith(("UBLKCP",)):
        return "bulk copies"
    return "arithmetic / control"


def main():
    rep, skip = sys.argv[1], int(sys.argv[2])
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(skip),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = rows[0][1]
    hdr = rows[1]
    idx = {}
    for i, h in enumerate(hdr):
        idx.setdefault(h, i)
    seen, data = set(), []
    for r in rows[2:]:
        if r and r[0].startswith("0x") and r[0] not in seen:       # the page repeats the listing: keep each address once
            seen.add(r[0])
            data.append(r)
    S, SRC = idx["# Samples"], idx["Source"]
    tot = sum(num(r[S]) for r in data) or 1
    cat = collections.Counter()
    for i, r in enumerate(data):
        src = r[SRC].strip()
        toks = src.split()
        op = toks[1] if src.startswith("@") and len(toks) > 1 else toks[0]
        cat[classify(op, data[i - 1][SRC] if i else "")] += num(r[S])
    reasons = {h[6:]: sum(num(r[j]) for r in data) for h, j in idx.items() if h.startswith("stall_") and "(" not in h}
    print(f"# {name}: {tot} warp-stall samples over {len(data)} SASS instructions\n")
    print("| instruction class (the sample is filed on the instruction the warp could not issue) | share |\n|---|---|")
    for k, v in cat.most_common():
        print(f"| {k} | {100 * v / tot:.1f} % |")
    print("\n| stall reason | share |\n|---|---|")
    for k, v in sorted(reasons.items(), key=lambda kv: -kv[1])[:8]:
        print(f"| {k} | {100 * v / tot:.1f} % |")
    print(f"\nhottest {topn} instructions:\n")
    for r in sorted(data, key=lambda r: -num(r[S]))[:topn]:
        print(f"  {100 * num(r[S]) / tot:5.1f} %  exec={r[idx['Instructions Executed']]:>9}  {r[SRC].strip()[:90]}")


if __name__ == "__main__":
    main()
