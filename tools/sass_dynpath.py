"""Dynamic instruction count of the common path through one SASS loop (offline proxy for issue slots per wavefront step).
Walks the loop from its head; BRA.DIV and branches leaving the loop are not taken; a forward conditional branch is TAKEN
(its region skipped) when the skipped region contains one of the given marker opcodes (= a rare block, e.g. the
once-per-8-steps boundary-ring refill "UMOV UR7, 0x10" or the exponent switch "MUFU.EX2"); unconditional branches are followed.
usage: python tools/sass_dynpath.py lib.so <kernel-substr> <loop_start_hex> <loop_end_hex> [marker ...]"""
import collections
import re
import sys

from sass_loops import kernel_instrs, opcode


def walk(ins, lo, hi, markers):
    idx = {a: i for i, (a, _) in enumerate(ins)}
    i, n, h = idx[lo], 0, collections.Counter()
    seen = set()
    while True:
        a, t = ins[i]
        if a in seen:
            break
        seen.add(a)
        n += 1
        h[opcode(t)] += 1
        if a == hi:
            break
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\w+,\s*)?0x([0-9a-f]+)", t)
        if m and "BRA.DIV" not in t:
            ta = int(m.group(1), 16)
            cond = t.startswith("@")
            if lo <= ta <= hi and ta > a:
                region = " ; ".join(x for _, x in ins[i + 1:idx[ta]])
                rare = any(mk in region for mk in markers)
                if rare and cond:   # skip only the INNERMOST conditional region that holds a marker
                    for j in range(i + 1, idx[ta]):
                        a2, t2 = ins[j]
                        m2 = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\w+,\s*)?0x([0-9a-f]+)", t2)
                        if m2 and t2.startswith("@") and "BRA.DIV" not in t2:
                            tb = int(m2.group(1), 16)
                            if a2 < tb <= ta and any(mk in " ; ".join(x for _, x in ins[j + 1:idx[tb]]) for mk in markers):
                                rare = False
                                break
                if not cond or rare:
                    i = idx[ta]
                    continue
        i += 1
    return n, h


if __name__ == "__main__":
    lib, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
    markers = sys.argv[5:] or ["UMOV UR7, 0x10 ", "MUFU.EX2"]
    name, ins = kernel_instrs(lib, pat)
    n, h = walk(ins, lo, hi, markers)
    fp = sum(h[k] for k in ("FFMA", "FMUL", "FADD", "FFMA2", "FMUL2", "FADD2"))
    print("%s: common path %d instrs (FP %d) | %s" % (lib, n, fp, " ".join("%s:%d" % kv for kv in h.most_common(16))))
