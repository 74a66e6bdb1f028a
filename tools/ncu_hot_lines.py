"""Aggregate the warp-stall samples of one profiled launch per CUDA source line.
usage: python tools/ncu_hot_lines.py report.ncu-rep <launch index in the report> [top N]
Needs a report captured with `--set full --import-source on` from a `-lineinfo` build (the default build flags)."""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, skip = sys.argv[1], int(sys.argv[2])
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--launch-skip", str(skip),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    cur = hdr = fn = None
    agg = collections.OrderedDict()

    def num(x):
        try:
            return int(float(x))
        except ValueError:
            return 0

    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            fn = r[1]
        elif r[0] == "Line No":
            hdr = r
            idx = {}
            for i, h in enumerate(hdr):
                idx.setdefault(h, i)
        elif hdr is not None and len(r) >= len(hdr) - 3 and r[0].isdigit():
            a = agg.setdefault((cur, int(r[0])), [0, 0, r[1].strip(), collections.Counter()])
            a[0] += num(r[idx["# Samples"]])
            a[1] += num(r[idx["Instructions Executed"]])
            for h, i in idx.items():
                if h.startswith("stall_") and "(" not in h:
                    a[3][h] += num(r[i])
    tot = sum(a[0] for a in agg.values()) or 1
    print(f"# {fn}: {tot} stall samples\n")
    print("| samples | source line | warp instr. | dominant stalls | source |\n|---|---|---|---|---|")
    for (f, line), (s, e, src, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
        dom = ", ".join(f"{k[6:]}:{v}" for k, v in st.most_common(2))
        print(f"| {100 * s / tot:5.1f} % | {f}:{line} | {e} | {dom} | `{src[:100]}` |")


if __name__ == "__main__":
    main()
