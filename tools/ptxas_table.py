"""profiles/<round>_ptxas_v.txt: registers, stack and spill bytes of every kernel of libldcbf_b200.so, from
`nvcc -Xptxas -v` with the flags of csrc/Makefile (compile only, nothing is linked).  Usage: python tools/ptxas_table.py OUT"""
import os, re, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "humanoid-navigation-using-mpc-ldcbf_b200", "csrc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I" + os.path.join(ROOT, "include"), "-I" + SRC, "-Xptxas", "-v"]
EXTRA = {"lidar.cu": ["-fmad=false"]}


def one(name):
    r = subprocess.run(["nvcc"] + FLAGS + EXTRA.get(name, []) + ["-c", os.path.join(SRC, name), "-o", os.devnull],
                       capture_output=True, text=True)
    return r.stderr


def main(out):
    files = sorted(f for f in os.listdir(SRC) if f.endswith(".cu"))
    with ThreadPoolExecutor(8) as ex:
        logs = list(ex.map(one, files))
    rows = {}
    for log in logs:
        cur = None
        for line in log.splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                cur = re.sub(r"\(.*$", "", cur).replace("void ", "").replace("ldcbf::", "")
                rows[cur] = [0, 0, 0, 0]
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and cur:
                rows[cur][1:] = [int(m.group(1)), int(m.group(2)), int(m.group(3))]
            m = re.search(r"Used (\d+) registers", line)
            if m and cur:
                rows[cur][0] = int(m.group(1))
    with open(out, "w") as f:
        f.write("# nvcc 12.9 -O3 -gencode arch=compute_100a,code=sm_100a -Xptxas -v, every kernel of libldcbf_b200.so (tools/ptxas_table.py)\n")
        f.write("%-90s %5s %6s %9s %9s\n" % ("kernel", "regs", "stack", "spill_st", "spill_ld"))
        for k in sorted(rows):
            f.write("%-90s %5d %6d %9d %9d\n" % ((k,) + tuple(rows[k])))


if __name__ == "__main__":
    main(sys.argv[1])
