"""SASS opcode census of the built library (cuobjdump -sass): which kernels carry tcgen05 / TMA / TMEM / mma.sync / FFMA2 opcodes."""
import re
import subprocess
import sys
from collections import Counter, OrderedDict

KEYS = ["UTCHMMA", "HMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOM", "SYNCS", "LDGSTS", "UBLKCP", "FFMA2", "MUFU", "REDUX", "ATOMS", "RED", "UTMACCTL",
        "ELECT", "ACQBULK"]


def main(so, out):
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    per = OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            per[cur]["_n"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    per[cur][k] += 1
    tot = Counter()
    for c in per.values():
        tot.update(c)
    lines = [f"SASS opcode census of {so} (cuobjdump -sass, sm_100a)",
             "tcgen05.mma -> UTC*MMA; tcgen05.ld/st -> LDTM/STTM; cp.async.bulk.tensor -> UTMALDG/UTMASTG; tcgen05.commit -> UTCBAR; mbarrier -> SYNCS; "
             "cp.async -> LDGSTS; cp.async.bulk (1-D) -> UBLKCP; fma.rn.f32x2 -> FFMA2; warp-level mma.sync -> HMMA", "",
             f"totals over {len(per)} kernels:"]
    for k in KEYS:
        if tot[k]:
            lines.append(f"  {k:<10} {tot[k]}")
    lines += ["", "per kernel (instructions | tensor / TMA / TMEM / async-copy opcodes):"]
    for name, c in sorted(per.items(), key=lambda kv: -kv[1]["_n"]):
        ops = " ".join(f"{k}={c[k]}" for k in KEYS if c[k])
        lines.append(f"  {c['_n']:6d}  {name}  {ops}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:24]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "whisper-diarize-rs_b200/csrc/libwdr_b200.so", sys.argv[2] if len(sys.argv) > 2 else "profiles/r02/sass_census.txt")
