"""List the loops (backward branches) of one kernel in a cuobjdump -sass dump with instruction counts and an opcode
histogram: the offline proxy used to judge issue-slot changes before spending GPU time.
usage: python tools/sass_loops.py lib.so <kernel-name-substring> [min_len] [max_len] [nth match]"""
import collections
import re
import subprocess
import sys


def kernel_instrs(lib, pat, nth=0):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    hits = [f for f in funcs[1:] if pat in f.split("\n", 1)[0]]
    f = hits[nth]
    name = f.split("\n", 1)[0]
    ins = []
    for m in re.finditer(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", f, re.M):
        ins.append((int(m.group(1), 16), m.group(2).strip()))
    return name, ins


def opcode(t):
    t = re.sub(r"^@!?U?P\w+\s+", "", t)
    return t.split()[0].split(".")[0]


if __name__ == "__main__":
    lib, pat = sys.argv[1], sys.argv[2]
    minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    maxlen = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
    nth = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    name, ins = kernel_instrs(lib, pat, nth)
    print(name, len(ins), "instructions")
    idx = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\w+,\s*)?0x([0-9a-f]+)", t)
        if m:
            ta = int(m.group(1), 16)
            if ta < a and ta in idx and minlen <= i - idx[ta] + 1 <= maxlen:
                body = ins[idx[ta]:i + 1]
                h = collections.Counter(opcode(x) for _, x in body)
                fp = sum(h[k] for k in ("FFMA", "FMUL", "FADD", "FFMA2", "FMUL2", "FADD2", "FMNMX", "FSEL"))
                print("loop %05x..%05x: %d instrs, FP %d | %s" % (ta, a, len(body), fp, " ".join("%s:%d" % kv for kv in h.most_common(14))))
