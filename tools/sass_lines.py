"""Join an ncu SASS-level source page (``ncu -i X.ncu-rep --page source --csv --kernel-name regex:K``) with the line table of
the cubin (``nvdisasm -g``): executed warp instructions and stall samples per CUDA source line.

    python tools/sass_lines.py <ncu_sass.csv> <cubin> <mangled kernel name substring> [top]

The cubin comes from ``cuobjdump -xelf all liblm_b200.so``.  Instructions are matched by their order inside the function
(the CSV lists absolute addresses, nvdisasm offsets), and the opcode of every pair is compared as a guard."""
import collections
import csv
import re
import subprocess
import sys


def disasm_lines(cubin, kernel):
    out = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True, check=True).stdout.splitlines()
    start = None
    for i, l in enumerate(out):
        if l.lstrip().startswith(".section") and ".text." in l and kernel in l:
            start = i
            break
    assert start is not None, "kernel not found in the cubin"
    cur, ins = ("?", 0), []
    for l in out[start + 1:]:
        if l.lstrip().startswith(".section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return ins


def main():
    path, cubin, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    cols = {n: k for k, n in enumerate(rows[hdr])}
    body = []
    for r in rows[hdr + 1:]:
        if r and r[0] == "Kernel Name":                           # the next captured launch: the first one is enough
            break
        if len(r) > cols["Instructions Executed"]:
            body.append(r)
    ins = disasm_lines(cubin, kernel)
    assert len(ins) == len(body), "instruction counts differ: %d in the cubin, %d in the profile (rebuilt since?)" % (len(ins), len(body))
    per = collections.defaultdict(lambda: [0, 0, 0, 0])          # warp instr, thread instr, stall samples, sass count
    mismatch = 0
    for (off, text, line), r in zip(ins, body):
        op_a = text.split()[1] if text.startswith("@") else text.split()[0]
        src = r[cols["Source"]].strip()
        op_b = src.split()[1] if src.startswith("@") else src.split()[0]
        mismatch += op_a.split(".")[0] != op_b.split(".")[0]
        a = per[line]
        a[0] += int(r[cols["Instructions Executed"]] or 0)
        a[1] += int(r[cols["Thread Instructions Executed"]] or 0)
        a[2] += int(r[cols["Warp Stall Sampling (All Samples)"]] or 0)
        a[3] += 1
    assert mismatch < 0.01 * len(ins), "%d opcodes differ: not the same build" % mismatch
    tot_i = sum(a[0] for a in per.values())
    tot_s = sum(a[2] for a in per.values())
    print("total warp instructions %d, stall samples %d, %d SASS instructions, %d source lines" % (tot_i, tot_s, len(ins), len(per)))
    print("%-28s %8s %7s %7s %7s %5s" % ("file:line", "Minstr", "instr%", "lanes", "stall%", "sass"))
    for line, a in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-28s %8.2f %7.2f %7.1f %7.2f %5d" % ("%s:%d" % line, a[0] / 1e6, 100.0 * a[0] / max(tot_i, 1), a[1] / max(a[0], 1),
                                                    100.0 * a[2] / max(tot_s, 1), a[3]))


if __name__ == "__main__":
    main()
