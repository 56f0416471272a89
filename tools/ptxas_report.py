"""Developer tool (CPU only): registers, spill bytes and static shared memory of every kernel of the product, from
`nvcc -Xptxas -v` with the build's flags.  Output kept under profiles/ (static evidence, not a measurement).

    python tools/ptxas_report.py > profiles/r1z_ptxas_resources.txt
"""
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "lagrangian_microbes_b200", "csrc")


def main():
    print("# regs  spill_st  spill_ld  smem_B  kernel      (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xptxas -v)")
    with tempfile.TemporaryDirectory() as tmp:
        for f in sorted(os.listdir(CSRC)):
            if not f.endswith(".cu"):
                continue
            r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                                "-Xptxas", "-v", "-c", os.path.join(CSRC, f), "-o", os.path.join(tmp, f + ".o")],
                               capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            print(f"## {f}")
            name = None
            spill = ("0", "0")
            for line in r.stderr.splitlines():
                m = re.search(r"Compiling entry function '(\S+)'", line)
                if m:
                    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                    name = re.sub(r"\(.*", "", name)
                    continue
                m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
                if m:
                    spill = m.groups()
                    continue
                m = re.search(r"Used (\d+) registers", line)
                if m and name:
                    sm = re.search(r"(\d+) bytes smem", line)
                    print(f"{m.group(1):>5} {spill[0]:>8} {spill[1]:>9} {sm.group(1) if sm else 0:>7}  {name}")
                    name = None


if __name__ == "__main__":
    main()
