#!/usr/bin/env python
"""SASS opcode histogram and ptxas resource table of gpt_image_edit_b200/lib/libb2f.so (no GPU needed).

    python scripts/sass_histogram.py > profiles/rNN_sass_opcode_histogram.json

Per kernel (demangled name): instruction count, the opcodes that prove the tcgen05 / TMEM / TMA path (UTCHMMA = tcgen05.mma,
UTMALDG / UTMASTG = cp.async.bulk.tensor, LDTM / STTM = tcgen05.ld / .st, UTCBAR = tcgen05.commit, SYNCS = mbarrier,
UCGABAR = cluster barrier, MUFU.EX2, the packed FFMA2 / FADD2 / FMUL2 arithmetic, F2FP packing) and any legacy tensor-core
opcode (HMMA / IMMA: there must be none).  Registers / spills / shared memory come from the `ptxas -v` logs the Makefile
keeps under build/ (`make` first)."""
import collections
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "gpt_image_edit_b200" / "lib" / "libb2f.so"
WATCH = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "UCGABAR", "MUFU",
         "FFMA2", "FADD2", "FMUL2", "F2FP", "HMMA", "IMMA", "LDGSTS", "STL", "LDL", "ATOMG", "RED", "SETMAXREG", "USETMAXREG"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def sass():
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[T\d]+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    return kernels


def ptxas():
    res = {}
    for log in sorted((ROOT / "build").glob("*.ptxas.log")):
        name = None
        for line in log.read_text().splitlines():
            m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
            if m:
                name = m.group(1)
            m = re.search(r"Used (\d+) registers.*?(?:, (\d+) bytes smem)?", line)
            if m and name:
                res.setdefault(name, {})["registers"] = int(m.group(1))
                sm = re.search(r"(\d+) bytes smem", line)
                if sm:
                    res[name]["static_smem_bytes"] = int(sm.group(1))
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and name:
                res.setdefault(name, {}).update(stack_bytes=int(m.group(1)), spill_store_bytes=int(m.group(2)),
                                                spill_load_bytes=int(m.group(3)))
    return res


def main():
    ks = sass()
    px = ptxas()
    dm = demangle(sorted(set(ks) | set(px)))
    total = collections.Counter()
    rows = []
    for k, c in sorted(ks.items(), key=lambda kv: -sum(kv[1].values())):
        fam = collections.Counter()
        for op, n in c.items():
            base = op.split(".")[0]
            total[base] += n
            if base in WATCH:
                fam[op if base == "MUFU" else base] += n
        short = re.sub(r"\(.*", "", dm.get(k, k).replace("(anonymous namespace)::", "").replace("void ", ""))
        rows.append({"kernel": short, "instructions": sum(c.values()), **{"ptxas": px.get(k)}, "opcodes": dict(sorted(fam.items()))})
    legacy = {k: total[k] for k in ("HMMA", "IMMA") if total[k]}
    out = {"library": str(LIB.relative_to(ROOT)), "kernels": len(rows), "instructions": sum(total.values()),
           "totals": {k: total[k] for k in WATCH if total[k]}, "legacy_tensor_core_opcodes": legacy,
           "kernels_with_spills": sorted({r["kernel"] for r in rows if (r["ptxas"] or {}).get("spill_store_bytes")}),
           "per_kernel": rows}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
