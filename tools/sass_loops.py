#!/usr/bin/env python
"""Loop bodies of one kernel (backward branches) with their instruction mix per MUFU.SIN (= per evaluation):
python tools/sass_loops.py <file.so> <mangled-name-substring>"""
import collections, re, subprocess, sys
sys.argv[1:3]
txt = subprocess.run([sys.executable, __file__.replace("sass_loops", "sass_fn"), sys.argv[1], sys.argv[2]],
                     capture_output=True, text=True).stdout
ins = []
for line in txt.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
    if m and "BRA" in t:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_index:
            body = ins[addr_index[tgt]:i + 1]
            ops = collections.Counter()
            for _, b in body:
                parts = b.split()
                op = parts[1] if parts[0].startswith("@") else parts[0]
                ops[op.split(".")[0]] += 1
            n_eval = ops.get("MUFU", 0) / 2
            if n_eval >= 1:
                print(f"loop {tgt:#x}..{a:#x}: {len(body)} instr, {n_eval:g} evals -> {len(body)/n_eval:.1f} per eval")
                print("   " + " ".join(f"{o}:{c/n_eval:.2f}" for o, c in ops.most_common(24)))
