"""Counts, per kernel of libd3d_b200.so, the SASS mnemonics that prove bulk copies (UBLKCP), tensor-map copies (UTMALDG /
UTMASTG), tcgen05 (UTC*MMA, UTCBAR, LDTM), mbarrier waits (SYNCS) and the vector float reduction, and writes
profiles/<round>_sass_summary.md.   usage: python tools/sass_summary.py [profiles/r02_sass_summary.md]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = {"UBLKCP": r"\bUBLKCP", "UTMALDG": r"\bUTMALDG", "UTMASTG": r"\bUTMASTG", "UTC*MMA (tcgen05.mma)": r"\bUTC[A-Z]*MMA",
       "UTCBAR (tcgen05.commit)": r"\bUTCBAR", "LDTM (tcgen05.ld)": r"\bLDTM", "SYNCS (mbarrier)": r"\bSYNCS",
       "REDG.E.ADD.F32x4": r"REDG\.E\.ADD\.F32x4", "LDGSTS": r"\bLDGSTS"}


def main(dst):
    lib = os.path.join(ROOT, "deep3dpointclouddenoising_b200", "libd3d_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    rows = []
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        cnt = {k: len(re.findall(p, f)) for k, p in PAT.items()}
        if any(v for k, v in cnt.items() if not k.startswith("SYNCS")):
            rows.append((f.split("\n", 1)[0], cnt))
    names = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
    out = ["# SASS evidence (cuobjdump -sass deep3dpointclouddenoising_b200/libd3d_b200.so, sm_100a)", "",
           "Instruction counts per kernel: UBLKCP = cp.async.bulk, UTMALDG / UTMASTG = cp.async.bulk.tensor (TMA tensor maps), "
           "UTC*MMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, SYNCS = mbarrier operations, REDG.E.ADD.F32x4 = "
           "red.global.add.v4.f32.  Regenerate: `python tools/sass_summary.py`.", "",
           "| kernel | " + " | ".join(PAT) + " |", "|---|" + "---:|" * len(PAT)]
    for (_, c), d in zip(rows, names):
        d = re.sub(r"\(anonymous namespace\)::", "", d)
        d = re.sub(r"\(.*", "", d)[:60]
        out.append(f"| `{d}` | " + " | ".join(str(c[k]) for k in PAT) + " |")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_summary.md"))
