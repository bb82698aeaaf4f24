"""Dev (CPU only): static instruction mix of the traversal kernel's march loop, per issue pipe, from cuobjdump SASS.

    python tools/sass_pipe_mix.py [object] [kernel-substring]

The march is issue-bound.  Heuristic pipe model (B300_MICROARCH.md "Pipe rates" + the fmaheavy / fmalite split of
earlier architectures): one issue slot per cycle per SM sub-partition; the `alu` pipe (IADD3/LOP3/SHF/FMNMX/ISETP/LEA/
MOV ...) takes one warp instruction per 2 cycles; FP32 FFMA/FMUL/FADD one per cycle over two fma pipes, of which only
one executes IMAD / half-precision ops.  A region needs at least max(issue, 2 x alu, fp32 + 2 x fmah) cycles per warp;
a lower bound to compare variants with, not a timing model.  Regions reported:
the innermost backward branch inside the march loop = one descent round; the march loop minus that round and minus the
shading block (the span holding the payload LDG.E.128s) = the per-step base."""
import re
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mega-nerf-viewer_b200", "csrc", "mnv_render.o")
KERNEL = sys.argv[2] if len(sys.argv) > 2 else "render_voxels_kernelILi9ELb0ELb0ELb0E"

FP32 = ("FFMA", "FMUL", "FADD", "FSWZADD")          # fmaheavy or fmalite
FMAH = ("IMAD", "HFMA2", "HADD2", "HMUL2")         # fmaheavy only (assumed, as on earlier architectures)
ALU = ("IADD3", "IADD", "LOP3", "SHF", "PRMT", "FMNMX", "FMNMX3", "ISETP", "FSETP", "LEA", "VIADD", "VIADDMNMX", "MOV", "SEL",
       "FSEL", "IABS", "IMNMX", "VIMNMX", "PLOP3", "P2R", "R2P", "BMSK", "SGXT", "FCHK", "CS2R", "S2R", "UMOV", "UISETP", "UIADD3",
       "ULOP3", "USHF", "ULEA", "UPLOP3", "UIMAD", "UFLO", "UPRMT", "USEL")
XU = ("MUFU", "FLO", "POPC", "I2F", "F2I", "F2F", "I2I", "BREV", "I2FP", "F2FP")
LSU = ("LDG", "STG", "LDS", "STS", "LDL", "STL", "LDC", "LDCU", "ATOM", "ATOMG", "RED", "SULD", "SUST", "LD", "ST")
CTL = ("BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "BREAK", "CALL", "RET", "NOP", "BAR", "YIELD", "NANOSLEEP")
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")


def pipe(op):
    base = op.split(".")[0]
    for name, group in (("fp32", FP32), ("fmah", FMAH), ("alu", ALU), ("xu", XU), ("lsu", LSU), ("ctl", CTL), ("fp64", FP64)):
        if base in group:
            return name
    return "other:" + base


def main():
    sass = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True, check=True).stdout
    ins, on = [], False
    for line in sass.splitlines():
        if "Function :" in line:
            on = KERNEL in line
            continue
        if not on:
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            text = m.group(2).strip()
            pred = ""
            if text.startswith("@"):
                pred, text = text.split(None, 1)
            ins.append((int(m.group(1), 16), text.split()[0], text))
    if not ins:
        sys.exit(f"kernel {KERNEL} not found in {OBJ}")
    addr = {a: i for i, (a, _, _) in enumerate(ins)}
    back = []  # (target index, branch index)
    for i, (a, op, text) in enumerate(ins):
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", text)
            if m and int(m.group(1), 16) in addr and int(m.group(1), 16) <= a:
                back.append((addr[int(m.group(1), 16)], i))
    # the march loop = the backward branch that spans the payload loads; unrolled builds have one body per copy
    payload = [i for i, (_, op, _) in enumerate(ins) if op.startswith("LDG.E.128")]
    loops = [b for b in back if any(b[0] <= p <= b[1] for p in payload)]
    if not loops:
        sys.exit("no march loop found")
    lo, hi = max(loops, key=lambda b: b[1] - b[0])
    rounds = [b for b in back if lo <= b[0] and b[1] < hi and not any(b[0] <= p <= b[1] for p in payload)]
    copies = max(1, len(rounds))

    def mix(idx):
        out = {}
        for i in idx:
            k = pipe(ins[i][1])
            out[k] = out.get(k, 0) + 1
        return out

    def show(name, idx, per=1):
        m = mix(idx)
        n = len(idx)
        alu, fmah, fp32 = m.get("alu", 0), m.get("fmah", 0), m.get("fp32", 0)
        print(f"{name:34s} {n / per:6.1f} instr   " + "  ".join(f"{k} {v / per:.1f}" for k, v in sorted(m.items())) +
              f"   | cycles >= issue {n / per:.0f}, 2*alu {2 * alu / per:.0f}, fp32+2*fmah {(fp32 + 2 * fmah) / per:.0f}")

    in_round = set()
    for t, b in rounds:
        in_round.update(range(t, b + 1))
    # shading block: from the first payload load of a copy to the instruction before the loop tail's t update; take the
    # span between the first 128-bit load and the last MUFU/FFMA.SAT-free join — approximated by [first payload, last STS]
    shade = set()
    body = list(range(lo, hi + 1))
    pl = [p for p in payload if lo <= p <= hi]
    sts = [i for i in body if ins[i][1].startswith("STS")]
    step = len(pl) // copies if copies else len(pl)
    for c in range(copies):
        ps = pl[c * step:(c + 1) * step] if step else pl
        if not ps:
            continue
        last_sts = max([s for s in sts if s > ps[0] and (c == copies - 1 or s < pl[(c + 1) * step])], default=ps[-1])
        shade.update(range(ps[0], last_sts + 1))
    print(f"{KERNEL}: march loop {ins[lo][0]:#x}..{ins[hi][0]:#x}, {copies} step(s) per trip")
    show("descent round", sorted(in_round), copies)
    show("step without rounds and shading", [i for i in body if i not in in_round and i not in shade], copies)
    show("shading block (payload .. stores)", sorted(shade), copies)


if __name__ == "__main__":
    main()
