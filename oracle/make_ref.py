"""Stage the reference's own CPU implementation of the hot path under oracle/_ref/ -- TEST INFRASTRUCTURE ONLY.

The arithmetic of WITCH's hot path lives in two prebuilt third-party binaries the reference ships
(/root/reference/witch_msa/tools/magus/tools/hmmer/{hmmsearch,hmmalign}, HMMER 3.1b2; hmmbuild is needed to
make eHMM inputs). There is no source to compile, so "building" oracle/_ref means copying those executables,
byte for byte, into oracle/_ref/hmmer/ (git-ignored, NOT gpurun-ignored: they travel to the GPU box like our own
.so files). No reference source file is copied into the repository.

Used by: tests/golden/make_golden.py (pins the oracle), bench.py --impl reference (CPU arm), tests that
cross-check against the live binaries when present.
"""
import os
import shutil
import stat
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/witch_msa/tools/magus/tools/hmmer"
DST = os.path.join(HERE, "_ref", "hmmer")
TOOLS = ("hmmsearch", "hmmalign", "hmmbuild")


def ref_tool(name):
    """Absolute path of a staged reference binary, or None."""
    p = os.path.join(DST, name)
    return p if os.path.isfile(p) and os.access(p, os.X_OK) else None


def have_ref():
    return all(ref_tool(t) for t in TOOLS)


def build():
    """Copy the binaries when the reference tree is mounted (this container); no-op on the GPU box."""
    if not os.path.isdir(SRC):
        return have_ref()
    os.makedirs(DST, exist_ok=True)
    for t in TOOLS:
        d = os.path.join(DST, t)
        if not os.path.isfile(d) or os.path.getsize(d) != os.path.getsize(os.path.join(SRC, t)):
            shutil.copyfile(os.path.join(SRC, t), d)
        os.chmod(d, os.stat(d).st_mode | stat.S_IXUSR | stat.S_IXGRP | stat.S_IXOTH)
    return have_ref()


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref staged:" if ok else "oracle/_ref NOT available:", DST)
    sys.exit(0 if ok else 1)
