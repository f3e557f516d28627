"""Where are the instructions of a kernel?  (development tool)

Maps the SASS of one kernel of libmpp_b200.so back to the source functions it was inlined from and lists the instruction
kinds that turned out to matter on the latency-bound path of the window sampler (profiles/r02_summary.md): local-memory
traffic (`LDL` / `STL`: references and out-pointers across out-of-line calls, run-time indexed locals, spills), generic
loads (`LD`: pointers whose address space the compiler does not know), range checks of IEEE divisions (`FCHK`), calls,
fences (`MEMBAR`) and L1 invalidations (`CCTL`).  Needs the library built with -lineinfo (build.py does).

    python tools/sass_map.py [kernel-name-substring]        (default: the production instantiation of k_windows_dataflow)
"""
import bisect
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mpp_cnn_rs_object_detection_b200", "libmpp_b200.so")
SRC = os.path.join(ROOT, "mpp_cnn_rs_object_detection_b200", "csrc")


def outline(path):
    """(first line, name) of every function definition of a source file (good enough for attribution)."""
    out = []
    for i, line in enumerate(open(path).read().split("\n"), 1):
        if re.match(r"^\s*(template\s*<[^>]*>\s*)?(__device__|__global__|static|inline)", line) and "(" in line and not line.rstrip().endswith(";"):
            m = re.search(r"([A-Za-z_0-9]+)\s*\(", line.split("__launch_bounds__")[-1] if "__launch_bounds__" not in line else line.split(")", 1)[-1])
            if m:
                out.append((i, m.group(1)))
    return out


def main():
    want = sys.argv[1] if len(sys.argv) > 1 else "k_windows_dataflowIfLi8ELb0ELb0"
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
        text = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True, check=True).stdout.split("\n")
    start = next(i for i, l in enumerate(text) if l.startswith(".text.") and want in l and l.rstrip().endswith(":"))
    end = next(i for i in range(start + 5, len(text)) if text[i].lstrip().startswith(".section"))
    print(text[start].strip())
    outlines = {n: outline(os.path.join(SRC, n)) for n in os.listdir(SRC) if n.endswith((".cuh", ".cu"))}

    def fn_of(file, line):
        st = outlines.get(file)
        if not st:
            return file
        i = bisect.bisect_right([s for s, _ in st], line) - 1
        return st[i][1] if i >= 0 else "?"

    cur, per_fn, special = ("?", 0), collections.Counter(), collections.Counter()
    kinds = r"LDL|STL|FCHK|CALL|MEMBAR|CCTL|LD|ST|ATOM|RED"
    for l in text[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_]+)", l)
        if not m:
            continue
        fn = fn_of(*cur)
        per_fn[fn] += 1
        op = m.group(2)
        if re.fullmatch(kinds, op):
            special[(op, fn, cur[0], cur[1])] += 1
    print(f"{sum(per_fn.values())} instructions")
    for fn, n in per_fn.most_common(30):
        print(f"  {n:6d}  {fn}")
    for op in ("FCHK", "CCTL", "MEMBAR", "CALL", "LDL", "STL", "LD", "ST"):
        rows = [(k, v) for k, v in special.items() if k[0] == op]
        if rows:
            print(f"{op}: {sum(v for _, v in rows)} sites")
            for (o, fn, f, ln), v in sorted(rows, key=lambda x: -x[1])[:12]:
                print(f"  {v:4d}  {fn}  ({f}:{ln})")


if __name__ == "__main__":
    main()
