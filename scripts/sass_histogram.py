"""Opcode histogram per kernel from `cuobjdump -sass` of a built object / library (evidence for profiles/: the TMA,
mbarrier and wide-access mnemonics, and the fp64 instruction mix).

    python scripts/sass_histogram.py exahype_b200/build/inst_euler3d.o [regex on the demangled kernel name] [--top N]
"""
import collections
import re
import subprocess
import sys


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def histogram(path):
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    per_fn, fn = collections.OrderedDict(), None
    ins = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)")
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            per_fn[fn] = collections.Counter()
            continue
        m = ins.match(line)
        if m and fn:
            per_fn[fn][m.group(1)] += 1
    return per_fn


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    top = 45
    if "--top" in sys.argv:
        top = int(sys.argv[sys.argv.index("--top") + 1])
        args = [a for a in args if a != str(top)]
    path, pattern = args[0], (args[1] if len(args) > 1 else ".")
    per_fn = histogram(path)
    names = demangle(list(per_fn))
    for fn, counts in per_fn.items():
        nice = names.get(fn, fn)
        if not re.search(pattern, nice):
            continue
        total = sum(counts.values())
        base = collections.Counter()
        for op, n in counts.items():
            base[op.split(".")[0]] += n
        print(f"== {nice}\n   {total} instructions")
        print("   by mnemonic: " + ", ".join(f"{op} {n}" for op, n in base.most_common(top)))
        keys = [op for op in counts if re.match(r"(UBLKCP|UBLKPF|SYNCS|LDG|STG|LDS|STS|ATOM|RED|MEMBAR|ERRBAR|BAR|CCTL)", op)]
        print("   memory / sync forms: " + ", ".join(f"{op} {counts[op]}" for op in sorted(keys, key=lambda o: -counts[o])))


if __name__ == "__main__":
    main()
