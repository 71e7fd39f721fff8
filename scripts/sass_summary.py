"""Per-kernel counts of the SASS mnemonics that prove a Blackwell tensor-core / TMA kernel (B200_PROFILING.md): UTC*MMA
(tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tile load / store), UTCBAR (tcgen05.commit), plus FFMA
for the CUDA-core kernels. Runs on the CPU box (cuobjdump only reads the cubin):

    python scripts/sass_summary.py [pdequinox_b200/libpdeqx_b200.so] > profiles/sass_summary.txt
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "pdequinox_b200/libpdeqx_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
PAT = [("UTCHMMA", r"\bUTCHMMA"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCQMMA/UTCOMMA", r"\bUTC[QO]MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
       ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("FFMA", r"\bFFMA\b"),
       ("FFMA2/FMUL2/FADD2", r"\bF(FMA|MUL|ADD)2\b"),
       ("HMMA/IMMA (legacy mma.sync)", r"\b[HI]MMA\b")]
rows, cur, idx = [], None, -1
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        idx += 1
        cur = collections.Counter()
        rows.append((names[idx] if idx < len(names) else m.group(1), cur))
        continue
    if cur is not None:
        for key, pat in PAT:
            if re.search(pat, line):
                cur[key] += 1
print(f"# SASS summary of {lib} (sm_100a cubin; `cuobjdump -sass`), one line per kernel with at least one tensor-core / TMA instruction")
print("# columns: " + " | ".join(k for k, _ in PAT))
tot = collections.Counter()
for name, c in sorted(rows, key=lambda r: -r[1]["UTCHMMA"]):
    tot.update(c)
    if c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"]:
        short = name
        if short.endswith(")"):  # drop the parameter list, keep template arguments
            depth = 0
            for i in range(len(short) - 1, -1, -1):
                depth += short[i] == ")"
                depth -= short[i] == "("
                if depth == 0:
                    short = short[:i]
                    break
        short = short.replace("pdeqx::", "").replace("void ", "")
        print(f"{short:90s} " + " ".join(f"{c[k]:6d}" for k, _ in PAT))
print(f"{'TOTAL (all ' + str(len(rows)) + ' kernels)':90s} " + " ".join(f"{tot[k]:6d}" for k, _ in PAT))
