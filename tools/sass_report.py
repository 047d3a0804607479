"""Static SASS report for one kernel of an in-tree object (no GPU needed): instruction count, opcode classes and the
source lines that own the most instructions (objects are built with -lineinfo).  The pointwise kernels of this engine
are issue-bound, so instructions per byte moved is the figure to drive down (profiles/r01_summary.md).

    python tools/sass_report.py stylegan.pytorch_b200/lib/sg_pointwise.o blur3_kernelI13 [--lines 12]"""
import argparse, collections, os, re, subprocess, tempfile

ap = argparse.ArgumentParser()
ap.add_argument("obj")
ap.add_argument("kernel", help="substring of the mangled kernel name")
ap.add_argument("--lines", type=int, default=10)
a = ap.parse_args()

with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(a.obj)], cwd=td, check=True, stdout=subprocess.DEVNULL)
    cubins = [os.path.join(td, f) for f in os.listdir(td) if f.endswith(".cubin")]
    if not cubins:                                   # a plain cubin / already an ELF
        cubins = [os.path.abspath(a.obj)]
    txt = "".join(subprocess.run(["nvdisasm", "--print-line-info", c], capture_output=True, text=True).stdout
                  for c in cubins)

found = False
for sec in re.split(r"//--------------------- \.text\.", txt)[1:]:
    name = sec.split()[0]
    if a.kernel not in name:
        continue
    found = True
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    ops, per_line, cur = collections.Counter(), collections.Counter(), None
    for line in sec.splitlines():
        m = re.search(r'//## File "(.*?)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            ops[m.group(1)] += 1
            per_line[cur] += 1
    total = sum(ops.values())
    print("%s\n  %d instructions" % (dem[:140], total))
    print("  opcodes: " + "  ".join("%s %d" % kv for kv in ops.most_common(14)))
    for (f, l), c in per_line.most_common(a.lines):
        print("  %6d  %s:%d" % (c, f, l))
if not found:
    raise SystemExit("no kernel matching %r in %s" % (a.kernel, a.obj))
