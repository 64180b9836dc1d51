"""SASS evidence per kernel of the in-tree library (no GPU needed):  python tools/sass_summary.py > profiles/r02_sass_summary.md
Counts the mnemonics that matter for the B200 discussion (B200_PROFILING.md 'What proves a Blackwell-native kernel')."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gnn-model-explainer_b200", "gnnx", "lib", "libgnnx.so")
KEYS = ["HMMA", "UBLKCP", "UTMALDG", "UTC", "LDTM", "SYNCS", "LDGSTS", "UCGABAR", "CCTL", "MEMBAR", "ATOM", "RED", "BAR", "LDG", "STG", "LDS", "STS", "FFMA", "FMUL", "FADD", "MUFU", "SHFL", "I2F", "F2F"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fn, counts, total = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", fn)
            fn = re.sub(r"\(.*\)$", "", fn)[:110]
            counts[fn] = collections.Counter(); total[fn] = 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            op = m.group(1)
            total[fn] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[fn][k] += 1
                    break
    print("# r02 — SASS summary of `libgnnx.so` (sm_100a), `cuobjdump -sass` mnemonic counts per kernel\n")
    print("`HMMA.1688.F32.TF32` = `mma.sync.m16n8k8` TF32 (the 3xTF32 feature GEMMs), `UBLKCP` = `cp.async.bulk` (TMA engine), `SYNCS.*` = mbarrier arrive / try_wait,")
    print("`UCGABAR` = hardware cluster barrier, `LDGSTS` = `cp.async` (first-generation streaming kernel), `CCTL.IVALL` = L1 invalidation emitted by gpu-scope acquire fences (gang barrier).\n")
    print("| kernel | instructions | " + " | ".join(KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    for fn, c in counts.items():
        if total[fn] < 200:
            continue
        print("| `%s` | %d | " % (fn, total[fn]) + " | ".join(str(c.get(k, 0)) if c.get(k, 0) else "" for k in KEYS) + " |")


if __name__ == "__main__":
    main()
