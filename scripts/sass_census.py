"""Per-kernel SASS census of libstp.so: counts of the mnemonics that prove the tcgen05 / TMA path (B200_PROFILING.md):
UTC*MMA (tcgen05.mma), UTMALDG / UTMASTG (TMA load / store), LDTM / STTM (tcgen05.ld / st), HMMA (legacy mma.sync),
LDGSTS (cp.async).   python scripts/sass_census.py > profiles/r2_sass_census.txt   (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "segmentation_training_pipeline_b200", "libstp.so")
PAT = [("UTCHMMA", re.compile(r"\bUTC\w*MMA\b")), ("UTMALDG", re.compile(r"\bUTMALDG\b")), ("UTMASTG", re.compile(r"\bUTMASTG\b")),
       ("LDTM", re.compile(r"\bLDTM\b")), ("STTM", re.compile(r"\bSTTM\b")), ("HMMA", re.compile(r"\bHMMA\b")),
       ("LDGSTS", re.compile(r"\bLDGSTS\b")), ("UTCBAR", re.compile(r"\bUTCBAR\b"))]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = {}
    try:
        import shutil
        if shutil.which("cu++filt"):
            pass
    except Exception:
        pass
    cur, counts, order, ninst = None, collections.defaultdict(collections.Counter), [], collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        if cur is None or "/*" not in line:
            continue
        if re.search(r"/\*[0-9a-f]{4}\*/", line):
            ninst[cur] += 1
            for k, p in PAT:
                if p.search(line):
                    counts[cur][k] += 1
    dem = subprocess.run(["cu++filt"] + order, capture_output=True, text=True).stdout.splitlines() if order else []
    if len(dem) != len(order):
        dem = order
    print("# SASS census of %s (cuobjdump -sass, sm_100a)" % os.path.relpath(LIB, ROOT))
    print("%-110s %8s " % ("kernel", "instrs") + " ".join("%8s" % k for k, _ in PAT))
    tot = collections.Counter()
    for f, d in zip(order, dem):
        d = re.sub(r"stp::\(anonymous namespace\)::|\(anonymous namespace\)::", "", d)
        d = re.sub(r"\((int|bool|unsigned int)\)", "", d)
        d = re.sub(r"\(.*", "", d)[:108]
        c = counts[f]
        for k, _ in PAT:
            tot[k] += c[k]
        print("%-110s %8d " % (d, ninst[f]) + " ".join("%8d" % c[k] for k, _ in PAT))
    print("%-110s %8d " % ("TOTAL (%d kernels)" % len(order), sum(ninst.values())) + " ".join("%8d" % tot[k] for k, _ in PAT))


if __name__ == "__main__":
    main()
