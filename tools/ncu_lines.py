#!/usr/bin/env python
"""Per-source-line stall profile of one kernel from an ncu report (CPU-side analysis, no GPU needed).

`ncu --page source --csv` lists SASS instructions with their stall samples but not the CUDA-C line they came from;
`nvdisasm -gi` on the cubin lists the same instructions with their (inlined) line chains.  This joins the two by
instruction order and sums the samples per OUTERMOST line (the line of the kernel body, so a butterfly helper inlined
at line N counts for line N).

    python tools/ncu_lines.py <report.ncu-rep> <library.so> <kernel regex> [top_n]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, lib, pat = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    kname = rows[0][1]
    hdr = rows[1]
    i_src, i_all, i_exec = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    insts, seen = [], set()
    for r in rows[2:]:  # (a report with several matching launches lists the kernel once per launch: keep the first)
        if len(r) > i_all and r[0].startswith("0x") and r[0] not in seen:
            seen.add(r[0])
            insts.append(r)
    # mangled name of the kernel: take it from the cubins
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    short = re.sub(r"^void ", "", kname).split("(")[0]
    base = short.split("<")[0].split("::")[-1]
    targs = re.findall(r"\(int\)(-?\d+)|\(bool\)(\d)", short)
    lines = None
    for cub in sorted(os.listdir(tmp)):
        if not cub.endswith(".cubin"):
            continue
        dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
        # candidate functions: mangled names containing the base name; pick the one whose instruction count matches
        starts = [i for i, ln in enumerate(dis) if ln.startswith(".text.") and base in ln]
        for s in starts:
            chain, cur, out = [], [], []
            for ln in dis[s + 1:]:
                if ln.startswith("\t.section") or ln.startswith(".text."):
                    break
                m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
                if m:
                    cur.append(int(m.group(2)))
                    continue
                if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
                    if cur:
                        chain = cur
                    cur = []
                    out.append((chain[-1] if chain else 0, chain[0] if chain else 0, ln.split("*/", 1)[1].strip()))
            if len(out) == len(insts):
                lines = out
                break
        if lines:
            break
    if not lines:
        sys.exit(f"no function with {len(insts)} instructions matching {base} found in {lib}")
    total = sum(int(r[i_all] or 0) for r in insts)
    by_line = {}
    for r, (outer, inner, sass) in zip(insts, lines):
        d = by_line.setdefault(outer, {"n": 0, "exec": 0, "st": {}, "ops": {}})
        v = int(r[i_all] or 0)
        d["n"] += v
        d["exec"] += int(r[i_exec] or 0)
        for h, i in stall_cols:
            if r[i] not in ("", "0"):
                d["st"][h] = d["st"].get(h, 0) + int(r[i])
        op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
        d["ops"][op] = d["ops"].get(op, 0) + v
    srcfile = None
    for ln in dis:
        m = re.match(r'\s*//## File "([^"]+)"', ln)
        if m:
            srcfile = m.group(1)
            break
    text = open(srcfile).read().splitlines() if srcfile and os.path.exists(srcfile) else []
    print(f"{kname}\n{len(insts)} SASS instructions, {total} stall samples; per kernel-body line (outermost of the inline chain):")
    for outer, d in sorted(by_line.items(), key=lambda kv: -kv[1]["n"])[:top]:
        st = sorted(d["st"].items(), key=lambda kv: -kv[1])[:3]
        ops = sorted(d["ops"].items(), key=lambda kv: -kv[1])[:3]
        code = text[outer - 1].strip()[:70] if 0 < outer <= len(text) else ""
        print(f"{100 * d['n'] / max(total, 1):5.1f}%  L{outer:<5d} {code:70s} {', '.join(f'{k[6:]}={v}' for k, v in st)} | {', '.join(f'{k}={v}' for k, v in ops)}")
    tot_st = {}
    for d in by_line.values():
        for k, v in d["st"].items():
            tot_st[k] = tot_st.get(k, 0) + v
    print("stall reasons overall:", ", ".join(f"{k[6:]}={100 * v / max(total, 1):.1f}%" for k, v in sorted(tot_st.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    main()
