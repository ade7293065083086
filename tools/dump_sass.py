#!/usr/bin/env python
"""Write SASS evidence of the hot-path kernels under profiles/sass/ (runs anywhere: cuobjdump reads the built libpof.so).

    python tools/dump_sass.py [--lib planar_optical_flow_b200/libpof.so] [--out profiles/sass]

For every kernel of the library: one line in `summary.txt` (instruction count, register use is in build.log) with the
counts of the Blackwell-specific opcodes that prove what it runs on - UTCHMMA[.2CTA] (tcgen05.mma), LDTM (tcgen05.ld),
UTMALDG (TMA tensor loads), UBLKCP (TMA bulk copies), UTCBAR / SYNCS (mbarrier traffic), UTCATOMSWS / UTCPMULTI etc.
For the kernels the bench line names, a file each with the opcode histogram and the full listing (encodings dropped),
the hot loops marked: a backward branch closes a loop; loops that contain a tensor-core / TMA / bulk-copy instruction or
belong to the innermost sample loop are the ones that matter.
"""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY_OPS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UTCBAR", "SYNCS", "UTCCP", "HMMA", "LDGSTS",
           "LDS", "STS", "LDG", "STG", "DFMA", "FFMA", "SHFL", "BAR")
FULL = {                       # file stem -> regex on the demangled name
    "conv_tc_kernel_256_2_f16": r"conv_tc_kernel<256, 2, true>",
    "conv_tc_kernel_128_2_f16": r"conv_tc_kernel<128, 2, true>",
    "conv_tc_kernel_64_2_f16": r"conv_tc_kernel<64, 2, true>",
    "gate_stream_kernel_11_fwd": r"gate_stream_kernel<11, 0>",
    "gate_stream_kernel_11_bwd": r"gate_stream_kernel<11, 1>",
    "cutout_scan_kernel_f32": r"cutout_scan_kernel<float, false>",
    "cutout_scan_kernel_f32_multi_scan": r"cutout_scan_kernel<float, true>",
    "gate_stream_kernel_11_bwd_scores": r"gate_stream_kernel<11, 2>",
    "bn_act_bwd_kernel_pool2": r"bn_act_bwd_kernel<2>",
    "cutout_scan_exact_kernel_f32": r"cutout_scan_exact_kernel<float, false>",
    "cutout_kernel_exact_f32_staged": r"cutout_kernel<float, false, true>",
    "nms_sweep_kernel": r"nms_sweep_kernel",
}
INSTR = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return [re.sub(r"pof::\(anonymous namespace\)::", "", o) for o in out[:len(names)]]


def parse(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    funcs, cur = [], None
    for line in txt.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = (m.group(1), [])
            funcs.append(cur)
            continue
        m = INSTR.match(line)
        if m and cur is not None:
            cur[1].append((int(m.group(1), 16), m.group(2).strip()))
    names = demangle([f[0] for f in funcs])
    return [(n, ins) for n, (_, ins) in zip(names, funcs)]


def opcode(text):
    t = text.split()
    if t and t[0].startswith("@"):
        t = t[1:]
    return t[0] if t else "?"


def loops(ins):
    """(start_addr, end_addr) of every backward branch."""
    out = []
    for addr, text in ins:
        if opcode(text).startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", text)
            if m and int(m.group(1), 16) <= addr:
                out.append((int(m.group(1), 16), addr))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "planar_optical_flow_b200", "libpof.so"))
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "sass"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    funcs = parse(a.lib)
    with open(os.path.join(a.out, "summary.txt"), "w") as f:
        f.write("cuobjdump -sass planar_optical_flow_b200/libpof.so (sm_100a): instructions and Blackwell-specific opcode counts per kernel\n")
        f.write("UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = TMA bulk copy, "
                "UTMAPF = TMA L2 prefetch, SYNCS = mbarrier ops\n\n")
        for name, ins in funcs:
            hist = collections.Counter(opcode(t) for _, t in ins)
            keys = collections.Counter()
            for op, n in hist.items():
                for k in ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTMAPF", "SYNCS", "UTCBAR"):
                    if op.startswith(k):
                        keys[op if k in ("UTCHMMA", "UTMALDG", "UBLKCP") else k] += n
            f.write("%-90s %6d instr  %s\n" % (name[:90], len(ins), "  ".join("%s x%d" % kv for kv in sorted(keys.items()))))
    for stem, pat in FULL.items():
        hit = [(n, ins) for n, ins in funcs if re.search(pat, n)]
        if not hit:
            sys.stderr.write("no kernel matches %s\n" % pat)
            continue
        name, ins = hit[0]
        hist = collections.Counter(opcode(t) for _, t in ins)
        lp = loops(ins)
        hot = []
        for s, e in lp:
            body = [t for ad, t in ins if s <= ad <= e]
            ops = collections.Counter(opcode(t).split(".")[0] for t in body)
            tag = [k for k in ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTMAPF", "LDS", "STS", "STG", "LDG") if ops.get(k)]
            hot.append((s, e, len(body), tag, ops))
        with open(os.path.join(a.out, stem + ".txt"), "w") as f:
            f.write("%s\n%d instructions (cuobjdump -sass, sm_100a, encodings dropped)\n\n" % (name, len(ins)))
            f.write("opcode histogram (whole kernel):\n")
            for op, n in hist.most_common():
                f.write("  %-34s %5d\n" % (op, n))
            f.write("\nloops (backward branches), innermost first by size:\n")
            for s, e, n, tag, ops in sorted(hot, key=lambda h: h[2]):
                f.write("  0x%05x..0x%05x  %4d instr  %s\n" % (s, e, n, " ".join("%s=%d" % (k, ops[k]) for k in tag)))
            f.write("\nlisting:\n")
            starts = {s for s, _, _, _, _ in hot}
            ends = {e for _, e, _, _, _ in hot}
            for addr, text in ins:
                mark = ("L>" if addr in starts else "  ") + ("<L" if addr in ends else "  ")
                f.write("%s /*%05x*/ %s\n" % (mark, addr, text))
    print("wrote", a.out)


if __name__ == "__main__":
    main()
