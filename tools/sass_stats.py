#!/usr/bin/env python
"""Static SASS statistics of one kernel: opcode histogram and instruction count (no GPU needed).

    python tools/sass_stats.py <cubin-or-so> <substring of the mangled or demangled kernel name> [--top N]

Used while slimming the sampling kernels: `cuobjdump -sass` of a scratch instantiation (seconds to compile) shows
whether ptxas predicated or branched, whether registers get zero-filled (CS2R), packed FFMA2/FMUL2 were emitted, ...
"""
import collections
import re
import subprocess
import sys


def functions(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], stdout=subprocess.PIPE, text=True, check=True).stdout
    cur, out = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur is not None:
            out[cur].append(m.group(2).strip())
    return out


def main():
    path, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 20
    for name, ins in functions(path).items():
        dem = subprocess.run(["cu++filt", name], stdout=subprocess.PIPE, text=True).stdout.strip()
        if pat not in name and pat not in dem:
            continue
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i).split()[0].split(".")[0] for i in ins)
        print(f"{dem[:110]}\n  {len(ins)} instructions")
        print("  " + "  ".join(f"{o}:{n}" for o, n in ops.most_common(top)))


if __name__ == "__main__":
    main()
