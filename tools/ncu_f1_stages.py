"""Where do the warp instructions of sa_conv1_stream_kernel go?  Buckets every SASS instruction of one profiled launch (ncu --set full
--import-source on) by pipeline stage, using the source lines ncu attaches to it (innermost inlining level first, call sites after).
usage: python tools/ncu_f1_stages.py report.ncu-rep
The line ranges below are those of scanobjectnn_b200/csrc/sa_train.cu at commit 22a74e3 (the build that was profiled)."""
import collections
import csv
import io
import subprocess
import sys

SEARCH, EXTRACT, CONV, EPILOGUE = (315, 491), (492, 535), (536, 617), (618, 690)
LAMBDA = (294, 313)            # extract_rows (called from the extract stage; from the search stage when HAS_U)
F2_HELPERS = (226, 232)        # f2_add / f2_sub / f2_mul / f2_fma / f2_unpack_bits: only the search uses them


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def stage(lines, op):
    body = [l for f, l in lines if f == "sa_train.cu" and l >= 236]
    if body:
        l = body[-1]                                   # outermost call site
        for name, (a, b) in (("search", SEARCH), ("extract + rows", EXTRACT), ("conv + store", CONV), ("statistics epilogue", EPILOGUE), ("extract + rows", LAMBDA)):
            if a <= l <= b:
                return name
        return "prologue (ranges, barriers, setmaxnreg)"
    if any(f == "sa_train.cu" and F2_HELPERS[0] <= l <= F2_HELPERS[1] for f, l in lines) or (op.startswith("SHF") and any(f == "sm_32_intrinsics.hpp" for f, _ in lines)):
        return "search"
    if any(f == "ball_query.cuh" for f, _ in lines):
        return "extract + rows"
    if any(f == "sm_100_rt.hpp" for f, _ in lines):
        return "conv + store"                          # __ffma2_rn / __fadd2_rn
    return "unattributed (inlined intrinsics)"


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    cur, hdr, line, cur_fn = None, None, None, None
    occ = collections.OrderedDict()
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            if cur_fn != r[1]:
                cur_fn = r[1]
                print("#", r[1])
        elif r[0] == "Line No":
            hdr = r
            ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        elif r[0] not in ("", "-"):
            line = (cur, num(r[0]))
        elif r[ia].startswith("0x"):
            occ.setdefault(r[ia], []).append((line, num(r[ie]), num(r[isamp]), r[3].strip()))
    agg, smp, ops = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    for v in occ.values():
        text = v[0][3]
        op = (text.split()[1] if text.startswith("@") else text.split()[0]).split(".")[0]
        s = stage([x[0] for x in v], op)
        agg[s] += v[0][1]
        smp[s] += v[0][2]
        ops[s][op] += v[0][1]
    tot, ts = sum(agg.values()), sum(smp.values())
    print(f"\n{tot / 1e6:.2f} M warp instructions, {ts} stall samples\n")
    print("| stage | warp instructions | share | stall samples | largest opcodes (M warp instructions) |\n|---|---|---|---|---|")
    for s, c in agg.most_common():
        print(f"| {s} | {c / 1e6:.2f} M | {100 * c / tot:.1f} % | {100 * smp[s] / ts:.1f} % | " + ", ".join(f"{o} {n / 1e6:.2f}" for o, n in ops[s].most_common(6)) + " |")


main()
