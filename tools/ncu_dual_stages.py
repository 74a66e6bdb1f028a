"""Where do the warp instructions of tc_sa_dual_kernel<3,2> (SA1 level, transposed last layer) go?  Buckets every SASS instruction of
one profiled launch (ncu --set full --import-source on) by phase of the tile loop, using the source lines ncu attaches to it
(innermost inlining level first, call sites after).
usage: python tools/ncu_dual_stages.py report.ncu-rep [launch index]
The line ranges below are those of scanobjectnn_b200/csrc/tc_mlp.cu at commit 22a74e3 (the build that was profiled)."""
import collections
import csv
import io
import subprocess
import sys

PHASES = [("prologue (TMEM alloc, weight images, constants)", 391, 490), ("tile claim + geometry (idx -> point -> centred xyz)", 491, 521),
          ("layer 1 on the FFMA2 pipe + fp16x2 split + tcgen05.st (A operand)", 522, 580),
          ("inner layer: MMA issue, wait, tcgen05.ld, affine + ReLU, split, H -> shared memory", 581, 616),
          ("last layer (transposed): MMA issue", 617, 637), ("next tile's geometry prefetch (under the MMAs)", 638, 656),
          ("last layer: wait, tcgen05.ld, in-thread max-pool, affine, store", 657, 691), ("range-guard flag, TMEM dealloc", 776, 802)]


# inlined helpers (file, first line, last line) -- the report lists them without the call site
HELPERS = [("packed FFMA2 (__ffma2_rn): layer-1 chain, BN affine, fp16 residual", "sm_100_rt.hpp", 1, 10 ** 6),
           ("fp16x2 operand split (float2 -> half2, residual, range tracking)", "cuda_fp16.hpp", 1, 10 ** 6),
           ("operand stores: tcgen05.st (A operand) / swizzled STS.128 (H operand)", "tc_mlp.cu", 203, 262),
           ("affine_chunk: affine + ReLU of the inner epilogue", "tc_mlp.cu", 341, 361),
           ("tensor-pipe token (acquire / release)", "tc_mlp.cu", 363, 389),
           ("mbarrier wait loop", "tc_common.cuh", 40, 56),
           ("MMA descriptors / issue / commit", "tc_common.cuh", 95, 180),
           ("MMA issue helpers (issue_tile)", "tc_mlp.cu", 263, 300)]


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    cmd = ["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass,cuda"]
    if len(sys.argv) > 2:
        cmd += ["--launch-skip", sys.argv[2], "--launch-count", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    cur, hdr, line, fn, nfn = None, None, None, None, 0
    occ = collections.OrderedDict()
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            if fn != r[1]:
                fn = r[1]
                nfn += 1
                print("#", fn)
        elif r[0] == "Line No":
            hdr = r
            ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        elif r[0] not in ("", "-"):
            line = (cur, num(r[0]))
        elif r[ia].startswith("0x") and nfn == 1:
            occ.setdefault(r[ia], []).append((line, num(r[ie]), num(r[isamp]), r[3].strip()))
    agg, smp, ops = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    for v in occ.values():
        locs = [x[0] for x in v]
        body = [l for (f, l) in locs if f == "tc_mlp.cu" and 391 <= l <= 802]
        name = None
        if body:
            for n, a, b in PHASES:
                if a <= body[-1] <= b:
                    name = "phase: " + n
                    break
        if name is None:
            # inlined helpers carry no call-site record in the report: bucket them by the helper itself
            f, l = locs[-1]
            name = "helper: other"
            for hn, hf, a, b in HELPERS:
                if f == hf and a <= l <= b:
                    name = "helper: " + hn
                    break
        text = v[0][3]
        op = (text.split()[1] if text.startswith("@") else text.split()[0]).split(".")[0]
        agg[name] += v[0][1]
        smp[name] += v[0][2]
        ops[name][op] += v[0][1]
    tot, ts = sum(agg.values()), sum(smp.values())
    print(f"\n{tot / 1e6:.2f} M warp instructions, {ts} stall samples\n")
    print("| phase | warp instructions | share | stall samples | largest opcodes (M warp instructions) |\n|---|---|---|---|---|")
    for s, c in agg.most_common():
        print(f"| {s} | {c / 1e6:.2f} M | {100 * c / tot:.1f} % | {100 * smp[s] / max(ts, 1):.1f} % | " + ", ".join(f"{o} {n / 1e6:.2f}" for o, n in ops[s].most_common(6)) + " |")


main()
