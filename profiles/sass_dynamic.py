"""Dynamic SASS picture of one kernel from an `ncu --set full --import-source on` report: executed warp
instructions per straight-line block (consecutive SASS instructions with the same execution count), the
out-of-line subroutines that actually ran (CALL targets with a non-zero count: the slow paths of the f64
division / square root), and the local-memory traffic of the hot blocks.

    python profiles/sass_dynamic.py <report.ncu-rep> <kernel-regex> [top_n] > profiles/<tag>_sass_dynamic_<kernel>.txt

The static view (instructions per source function) is sass_sections.py; this one says how often each block RAN,
which is what found the two things r02w removed: the integer detour of the packet bounds' min / max reductions in
k_visibility_fused (125 instructions per packet) and the division slow path that k_sort_write entered for every
observation of the synthetic cities (zero numerator).
"""
import csv
import io
import re
import subprocess
import sys


def load(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                          "regex:" + kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name, hdr, ins = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "Kernel Name":
            if name is not None:
                break  # first matching launch only
            name = r[1]
            continue
        if r[0] == "Address":
            hdr = r
            continue
        if hdr is None or len(r) <= hdr.index("Instructions Executed"):
            continue
        try:
            e = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        ins.append((r[0], r[1].strip(), e, int(r[hdr.index("# Samples")] or 0)))
    return name, ins


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 24
    name, ins = load(rep, kernel)
    tot = sum(x[2] for x in ins)
    samples = sum(x[3] for x in ins) or 1
    print(f"{name}\n{len(ins)} SASS instructions, {tot} warp instructions executed, {samples} stall samples\n")
    blocks, cur = [], None
    for i, (_, src, e, sm) in enumerate(ins):
        if cur and abs(cur["e"] - e) <= 0.02 * max(e, 1) and not ins[i - 1][1].startswith(("RET", "EXIT", "BRA ", "CALL")):
            cur["n"] += 1
            cur["tot"] += e
            cur["samp"] += sm
            cur["end"] = i
        else:
            cur = {"start": i, "end": i, "e": e, "n": 1, "tot": e, "samp": sm}
            blocks.append(cur)
    print(f"hottest straight-line blocks (index range in the kernel's SASS, instructions, executions, share of the warp "
          f"instructions, share of the stall samples, first .. last instruction):")
    for b in sorted(blocks, key=lambda b: -b["tot"])[:top]:
        first, last = ins[b["start"]][1], ins[b["end"]][1]
        local = sum(1 for i in range(b["start"], b["end"] + 1) if re.match(r"(@!?U?P\d+\s+)?(LDL|STL)", ins[i][1]))
        print(f"  [{b['start']:5d}-{b['end']:5d}] {b['n']:4d} instr x {b['e']:9d} = {100 * b['tot'] / tot:5.2f} %  "
              f"samples {100 * b['samp'] / samples:5.2f} %  local ld/st {local:2d}   {first[:34]:34s} .. {last[:34]}")
    # subroutines that ran: executed instructions from each CALL target up to its RET
    addr_index = {a[-5:]: i for i, (a, _, _, _) in enumerate(ins)}
    print("\nout-of-line calls that were executed (call site index, executions, target, instructions the target ran):")
    seen_targets = {}
    any_call = False
    for i, (_, src, e, _) in enumerate(ins):
        m = re.match(r"(@!?U?P\d+\s+)?CALL\S*\s+(0x[0-9a-f]+)", src)
        if not m or e == 0:
            continue
        any_call = True
        t = m.group(2)[-5:]
        if t not in seen_targets and t in addr_index:
            j, s = addr_index[t], 0
            while j < len(ins):
                s += ins[j][2]
                if ins[j][1].startswith("RET"):
                    break
                j += 1
            seen_targets[t] = s
        print(f"  [{i:5d}] x {e:9d}  -> ..{t}   ({100 * seen_targets.get(t, 0) / tot:5.2f} % of the kernel's warp instructions in that routine)")
    if not any_call:
        print("  none")
    ops = {}
    for _, src, e, _ in ins:
        m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", src)
        if m:
            ops[m.group(1)] = ops.get(m.group(1), 0) + e
    print("\nexecuted opcode mix (top 16):")
    for k, v in sorted(ops.items(), key=lambda x: -x[1])[:16]:
        print(f"  {k:10s} {100 * v / tot:5.2f} %")


if __name__ == "__main__":
    main()
