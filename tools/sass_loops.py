"""Per-loop opcode histogram of one kernel's SASS (read here, on the CPU box; cuobjdump needs no GPU).

  python tools/sass_loops.py <lib.so> <kernel-name regex> <out.sass> [--per N]

Finds the kernel whose (mangled or demangled) name matches, finds its loops (a backward branch BRA <target> with
target <= address; nested loops are listed innermost first by size), and writes for every loop the opcode histogram
grouped by issue pipe plus, for the HOT loop (--key: highest density of that opcode; default: most instructions), its
listing. `--per N` divides the counts of
the hottest loop by N (units of work per trip: e.g. comparisons per trip of the K1 inner loop = unroll x queries per
thread) so the numbers can be read as instructions per unit. The output is what DESIGN.md / bench.py mean by
"counted in the SASS".
"""
import argparse
import collections
import re
import subprocess

PIPES = [
    ("xu", r"^(POPC|MUFU|FLO|BREV|F2I|I2F|F2F|FRND|I2I)"),
    ("fp64", r"^(DFMA|DADD|DMUL|DSETP|DMNMX)"),
    ("fma", r"^(IMAD|FFMA|FMUL|FADD|IDP|IMUL)"),
    ("alu", r"^(LOP3|IADD3|VIADD|LEA|SHF|SEL|ISETP|VIMNMX|IMNMX|IABS|PLOP3|MOV|PRMT|FSETP|FMNMX|FSEL|P2R|R2P|SGXT|BMSK|UIADD|ULOP|VABSDIFF)"),
    ("lsu", r"^(LDS|STS|LDG|STG|LD|ST|LDL|STL|ATOM|RED|ATOMS|ATOMG|LDSM|STSM|LDC)"),
    ("uniform", r"^(U[A-Z0-9]+|R2UR|S2UR|REDUX|VOTEU)"),
    ("control", r"^(BRA|BSSY|BSYNC|EXIT|RET|CALL|WARPSYNC|BAR|NANOSLEEP|YIELD|BREAK|BMOV|SYNCS|DEPBAR|ERRBAR|MEMBAR|FENCE|NOP|ELECT)"),
    ("warp", r"^(SHFL|VOTE|MATCH|S2R|CS2R)"),
]


def pipe_of(op):
    for name, pat in PIPES:
        if re.match(pat, op):
            return name
    return "other"


def kernel_sass(lib, pattern):
    text = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", text)[1:]
    rx = re.compile(pattern)
    hits = []
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        if rx.search(name) or rx.search(dem):
            hits.append((name, dem, f))
    if len(hits) != 1:
        raise SystemExit(f"{len(hits)} kernels match {pattern!r}: " + "; ".join(h[1][:100] for h in hits[:8]))
    return hits[0]


def parse(body):
    ins = []
    for line in body.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\*", line)
        if not m:
            continue
        addr, text = int(m.group(1), 16), m.group(2).strip()
        parts = text.split()
        pred = parts[0] if parts[0].startswith("@") else ""
        op = parts[1] if pred else parts[0]
        ins.append((addr, op, text))
    return ins


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("lib")
    ap.add_argument("pattern")
    ap.add_argument("out")
    ap.add_argument("--per", type=float, default=0.0)
    ap.add_argument("--key", default="", help="opcode that marks the hot loop (POPC, DFMA ...): the loop with the highest "
                                              "density of it is the one listed; default: the largest loop")
    args = ap.parse_args()
    name, dem, body = kernel_sass(args.lib, args.pattern)
    ins = parse(body)
    loops = []
    for addr, op, text in ins:
        if op.startswith("BRA"):
            m = re.search(r"(0x[0-9a-f]+)\s*$", text)
            if m and int(m.group(1), 16) <= addr:
                loops.append((int(m.group(1), 16), addr))
    loops.sort(key=lambda l: l[1] - l[0])
    with open(args.out, "w") as f:
        f.write(f"# kernel   {dem}\n# mangled  {name}\n# library  {args.lib}\n# total    {len(ins)} instructions, {len(loops)} loops\n")
        hottest = max(loops, key=lambda l: l[1] - l[0]) if loops else None
        if args.key and loops:
            def density(l):
                inside = [i for i in ins if l[0] <= i[0] <= l[1]]
                k = sum(1 for i in inside if i[1].split(".")[0] == args.key)
                return (k / len(inside)) if k >= 8 else 0.0
            hottest = max(loops, key=density)
        for lo, hi in loops:
            inside = [i for i in ins if lo <= i[0] <= hi]
            hist = collections.Counter(i[1].split(".")[0] for i in inside)
            pipes = collections.Counter()
            for op, n in hist.items():
                pipes[pipe_of(op)] += n
            tag = "  <== hot loop" if (lo, hi) == hottest else ""
            f.write(f"\n## loop 0x{lo:04x}..0x{hi:04x}: {len(inside)} instructions{tag}\n")
            f.write("   by pipe: " + ", ".join(f"{p} {n}" for p, n in pipes.most_common()) + "\n")
            f.write("   opcodes: " + ", ".join(f"{op} {n}" for op, n in hist.most_common()) + "\n")
            if (lo, hi) == hottest and args.per > 0:
                f.write(f"   per unit of work ({args.per:g} units per trip): " +
                        ", ".join(f"{op} {n / args.per:.2f}" for op, n in hist.most_common(12)) + "\n")
                f.write("   per unit by pipe: " + ", ".join(f"{p} {n / args.per:.2f}" for p, n in pipes.most_common()) + "\n")
        if hottest:
            f.write(f"\n## listing of the hot loop 0x{hottest[0]:04x}..0x{hottest[1]:04x}\n")
            for addr, op, text in ins:
                if hottest[0] <= addr <= hottest[1]:
                    f.write(f"  /*{addr:04x}*/ {text}\n")
    print(f"{args.out}: {dem[:80]}: {len(ins)} instructions, {len(loops)} loops")


if __name__ == "__main__":
    main()
