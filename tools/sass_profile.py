"""Static SASS profile of the scan4 kernel: instructions per phase / source line / pipe, from `nvdisasm -gi`.

The scan part of a block is straight-line code (everything is unrolled), so static counts there ARE the dynamic
warp-instructions per 4 KiB block; loops (emit, look-back) show their body sizes.  No GPU needed:

  python tools/sass_profile.py [--lib simdjson_b200/libsjb200.so] [--lines]
"""
import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "simdjson_b200", "csrc", "sjb200_scan4.cuh")

ALU = {"LOP3", "SHF", "PRMT", "IADD3", "VIADD", "ISETP", "SEL", "LEA", "PLOP3", "MOV", "IABS", "VIADDMNMX", "IMNMX", "P2R", "R2P", "BMSK", "SGXT", "VABSDIFF"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2"}
XU = {"POPC", "FLO", "BREV", "MUFU"}
LSU = {"LDS", "STS", "LDG", "STG", "LDL", "STL", "LD", "ST", "ATOMG", "ATOMS", "RED", "SHFL", "LDSM", "STSM"}
UNI = {"UMOV", "ULOP3", "UIADD3", "USHF", "ULEA", "R2UR", "S2UR", "UFLO", "UPOPC", "VOTEU", "REDUX", "CREDUX", "UISETP", "USEL", "UPRMT", "UIMAD", "ULDC", "LDCU", "UTMALDG", "ELECT"}
CTL = {"BRA", "BSSY", "BSYNC", "BRX", "BREAK", "EXIT", "WARPSYNC", "NOP", "YIELD", "BAR", "CALL", "RET"}


def pipe_of(op):
    for name, s in (("alu", ALU), ("fma", FMA), ("xu", XU), ("lsu", LSU), ("uni", UNI), ("ctl", CTL)):
        if op in s:
            return name
    return "other"


def function_ranges():
    """[(name, first_line, last_line)] of the SJ_DEV functions of sjb200_scan4.cuh"""
    starts = []
    lines = open(SRC).read().split("\n")
    for i, l in enumerate(lines, 1):
        m = re.match(r"(?:template <[^>]*>\s*)?SJ_DEV\s+[\w:<> \*&]+?\s+(\w+)\(", l)
        if m:
            starts.append((m.group(1), i))
    out = []
    for k, (n, s) in enumerate(starts):
        e = starts[k + 1][1] - 1 if k + 1 < len(starts) else len(lines)
        out.append((n, s, e))
    return out


ROOTS = {"scan_role", "chain_role", "scan4_body"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "simdjson_b200", "libsjb200.so"))
    ap.add_argument("--lines", action="store_true", help="per source line inside scan_block / emit")
    a = ap.parse_args()
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(a.lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.startswith("sjb200_kernels.") and f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], cwd=tmp, stdout=subprocess.PIPE, text=True).stdout
    fr = function_ranges()

    def func_of(line):
        for n, s, e in fr:
            if s <= line <= e:
                return n
        return None

    phase_cnt = collections.defaultdict(collections.Counter)
    line_cnt = collections.defaultdict(collections.Counter)
    infunc, chain = False, []
    pending = []  # instructions waiting for their annotation block? (annotations precede the instructions)
    for l in dis.split("\n"):
        if l.startswith(".text."):
            infunc = "scan4_kernel" in l
            chain = []
            continue
        if not infunc:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            if not pending_is_annot[0]:
                chain = []
            pending_is_annot[0] = True
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m:
            pending_is_annot[0] = False
            op = m.group(1)
            # outermost frame that lies in a non-root scan4 function decides the phase
            phase, pline = None, None
            for f, ln in reversed(chain):
                if f == "sjb200_scan4.cuh":
                    fn = func_of(ln)
                    if fn and fn not in ROOTS:
                        phase, pline = fn, ln
                        break
            if phase is None:
                for f, ln in reversed(chain):
                    if f == "sjb200_scan4.cuh" and func_of(ln):
                        phase, pline = func_of(ln) + " (glue)", ln
                        break
            if phase is None:
                phase, pline = "?", 0
            phase_cnt[phase][pipe_of(op)] += 1
            line_cnt[(phase, pline)][pipe_of(op)] += 1
    print(f"{'phase':28s} {'total':>6s} {'alu':>6s} {'fma':>6s} {'xu':>5s} {'lsu':>5s} {'uni':>5s} {'ctl':>5s} {'other':>6s}")
    for ph, c in sorted(phase_cnt.items(), key=lambda kv: -sum(kv[1].values())):
        print(f"{ph:28s} {sum(c.values()):6d} {c['alu']:6d} {c['fma']:6d} {c['xu']:5d} {c['lsu']:5d} {c['uni']:5d} {c['ctl']:5d} {c['other']:6d}")
    if a.lines:
        src = open(SRC).read().split("\n")
        for (ph, ln), c in sorted(line_cnt.items()):
            if ph in ("scan_block", "emit_columns", "emit_block") and sum(c.values()) >= 3:
                print(f"{ph:14s} L{ln:4d} {sum(c.values()):5d} alu={c['alu']:4d} fma={c['fma']:4d} xu={c['xu']:3d} lsu={c['lsu']:3d}  | {src[ln-1].strip()[:90]}")


pending_is_annot = [False]
if __name__ == "__main__":
    sys.exit(main())
